"""How many Jacobi iterations does NATRIX_OPT_WARM_START save?  Config 2 (1024^2, obstacles, vorticity), 60 frames:
the residual is the RMS divergence of the projected velocity of the last frame (oracle.divergence on the host).

    python scripts/warm_start_probe.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from natrix_b200 import workloads as W  # noqa: E402
from natrix_b200.core.fluid_simulator import FluidSimulator  # noqa: E402
from oracle import natrix_oracle as O  # noqa: E402

w = W.cfg2_workload()
w.splats_per_step, w.orbit_seed = 4, 3          # keep stirring: a decaying flow would make any solver look good
print(f"{'mode':>6s} {'N':>4s} {'rms div after projection':>26s} {'ms/step':>9s}")
for warm, n in [(False, 50), (False, 100), (False, 200), (True, 10), (True, 20), (True, 30), (True, 50)]:
    w.iterations = n
    sim, _ = W.build(w, FluidSimulator, None)
    sim.warm_start = warm
    sim.set_option(2, 1)
    ms = []
    for k in range(60):
        W.run_step(w, sim, None, k)
        if k >= 50:
            ms.append(sum(sim.timings().values()))
    for (px, py, r) in w.circles:                 # one more frame, read back BEFORE its impulses are added
        sim.add_circle_obstacle((px, py), r)
    sim.update(W.DT)
    for (px, py, r) in w.circles:                 # the obstacle map of that frame (update() cleared it)
        sim.add_circle_obstacle((px, py), r)
    v, obs = sim.download("velocity"), sim.download("obstacles")
    div = O.divergence(v, obs)
    fluid = obs[..., 0] + obs[..., 1] == 0
    print(f"{'warm' if warm else 'cold':>6s} {n:4d} {float(np.sqrt(np.mean(div[fluid].astype(np.float64) ** 2))):26.6e} {np.mean(ms):9.3f}")
    sim.destroy()
