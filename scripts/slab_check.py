"""Multi-GPU parity: the row-slab run over NCCL must be bit-identical to the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/slab_check.py [width height steps iterations]

Every rank computes the single-GPU reference of the whole grid on its own device, then compares
its slab's rows of every field - and of a dye field cut into the same slabs (2x the grid's resolution
under pipeline 1, 1.5x under pipeline 0) - after each step."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from natrix_b200 import _lib as L, workloads as W  # noqa: E402
from natrix_b200.core.fluid_simulator import FluidSimulator  # noqa: E402
from natrix_b200.slabs import SlabSimulator, SlabSmoothParticlesArea  # noqa: E402
from natrix_b200.smooth_particles_area import SmoothParticlesArea  # noqa: E402

width, height, steps, iters = (int(a) for a in (sys.argv[1:5] + ["1024", "2048", "3", "37"][len(sys.argv) - 1:]))
warm = len(sys.argv) > 5 and sys.argv[5] == "warm"        # NATRIX_OPT_WARM_START on both sides
rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

rng = np.random.default_rng(11)
v0 = (0.8 * rng.uniform(-1, 1, (height, width, 2))).astype(np.float32)
circles = [(0.3, 0.25, 40.0), (0.7, 0.5, 70.0), (0.5, 0.98, 30.0)]     # one straddles a slab boundary for N = 2, 4
splats = [((0.5, 0.5), (0.9, -0.6), 48.0), ((0.2, 0.74), (-0.5, 0.8), 25.0)]
ok = True
for pipeline in (1, 0):
    ref = FluidSimulator(width, height, None, device=local)
    ref.set_option(L.OPT_PIPELINE, pipeline)
    slab = SlabSimulator(width, height, device=local, depth=8)
    slab.sim.set_option(L.OPT_PIPELINE, pipeline)
    for s in (ref, slab.sim):
        s.vorticity, s.viscosity, s.iterations = 1.0, (0.3 if pipeline else 0.0), iters
        s.warm_start = warm
    slab.iterations = iters
    ref.upload("velocity", v0)
    slab.sim.upload("velocity", v0[slab.row0:slab.row0 + slab.rows])
    pw, ph = (2 * width, 2 * height) if pipeline else (3 * width // 2, 3 * height // 2)
    ref_dye, slab_dye = SmoothParticlesArea(pw, ph, ref), SlabSmoothParticlesArea(pw, ph, slab)
    for d in (ref_dye, slab_dye):
        d.dissipation = 0.98
    for k in range(steps):
        for s in (ref, slab):
            for (px, py, r) in circles:
                s.add_circle_obstacle((px, py), r)
            s.add_triangle_obstacle((0.55, 0.1), (0.9, 0.2), (0.6, 0.45))
            s.update(W.DT)
            for pos, vel, r in splats:
                s.add_velocity(pos, vel, r)
        for d in (ref_dye, slab_dye):
            d.add_particles((0.5, 0.5), 0.2 * ph, 0.6)            # straddles the slab boundaries
            d.add_particles((0.2, 0.74), 0.05 * ph, 0.9)
            d.update(W.DT)
        for name in ("velocity", "pressure", "divergence", "vorticity"):
            a = ref.download(name)[slab.row0:slab.row0 + slab.rows]
            b = slab.sim.download(name)
            same = bool(np.array_equal(a, b))
            ok &= same
            if not same or k == steps - 1:
                print(f"[rank {rank}/{world}] pipeline {pipeline} step {k} {name}: bit-identical={same} "
                      f"max|diff|={float(np.abs(a - b).max()):.3e} max|ref|={float(np.abs(a).max()):.3e}", flush=True)
        a = ref_dye.download()[slab_dye.row0:slab_dye.row0 + slab_dye.rows]
        b = slab_dye.engine.area.download()
        same = bool(np.array_equal(a, b)) and bool(np.array_equal(
            ref_dye.export_rgba8()[slab_dye.row0:slab_dye.row0 + slab_dye.rows], slab_dye.engine.area.export_rgba8()))
        ok &= same
        if not same or k == steps - 1:
            print(f"[rank {rank}/{world}] pipeline {pipeline} step {k} dye {pw}x{ph}: bit-identical={same} "
                  f"max|diff|={float(np.abs(a - b).max()):.3e} max|ref|={float(np.abs(a).max()):.3e}", flush=True)
    ref_dye.destroy()
    slab_dye.engine.area.destroy()
    ref.destroy()
    slab.sim.destroy()
flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
dist.all_reduce(flag)
if rank == 0:
    print("SLAB_CHECK", "PASS" if flag.item() == 0 else "FAIL", f"world={world} grid={width}x{height}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
