"""SURVEY 8(f)-4 on the GPU: residual of the pressure system after one projection against the device time of the solve,
for the reference's Jacobi loop (cold and warm-started) and the opt-in red-black SOR / multigrid solvers.

    python scripts/solver_probe.py [size]        (default 4096; config 3's workload without dye)

Residual = RMS over fluid cells of x1 + x2 + y1 + y2 - 4 p - div (the system shader.Poisson.comp iterates on),
evaluated on the host from the downloaded pressure / divergence / obstacles of the LAST of 6 steps."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from natrix_b200 import _lib as L  # noqa: E402
from natrix_b200 import workloads as W  # noqa: E402
from natrix_b200.core.fluid_simulator import FluidSimulator  # noqa: E402
from oracle import natrix_oracle as O  # noqa: E402   (the checker computes the residual; nothing here is product path)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = W.cfg4_workload(size)
CASES = [("jacobi", n, False) for n in (50, 100, 200, 400, 800)] + [("jacobi", n, True) for n in (25, 50, 100)] + \
        [("sor", n, False) for n in (25, 50, 100)] + [("multigrid", n, False) for n in (1, 2, 3, 4)] + [("multigrid", 1, True)]
rows = []
for solver, iters, warm in CASES:
    sim, _ = W.build(w, FluidSimulator, None)
    sim.solver, sim.iterations, sim.warm_start = solver, iters, warm
    sim.set_option(L.OPT_TIMING, 1)
    ms = []
    for k in range(6):
        for (px, py, r) in w.circles:
            sim.add_circle_obstacle((px, py), r)
        obstacles = sim.download("obstacles") if k == 5 else None
        sim.update(W.DT)
        ms.append(sim.timings()["jacobi"])
        for (px, py, vx, vy) in W.orbit_positions(w, k):
            sim.add_velocity((px, py), (vx, vy), w.splat_radius)
    p, div = sim.download("pressure"), sim.download("divergence")
    solid = O.solid(obstacles)
    nb = O.neighbours(solid)
    r = O.poisson_sweep(p, div, None, nb) * np.float32(4.0) - np.float32(4.0) * p
    res = float(np.sqrt(np.mean(r[~solid].astype(np.float64) ** 2)))
    row = {"solver": solver, "iterations": iters, "warm_start": warm, "rms_residual": res, "solve_ms": float(np.median(ms[2:])),
           "grid": [w.width, w.height]}
    rows.append(row)
    print(json.dumps(row), flush=True)
    sim.destroy()
