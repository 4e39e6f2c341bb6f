"""Multi-GPU parity check: the row-slab run over NCCL must be bit-identical to the single-GPU run.

Used by ``bench.py --gpus N`` (every N > 1 run executes it before timing and prints the outcome in its JSON line
as ``slab_parity``), by ``scripts/slab_check.py`` and by ``tests/test_gpu_multi.py``.  Every rank computes the
single-GPU result of the whole (small) grid on its own device, then compares its slab's rows of every field - and
of a dye field cut into the same slabs - after each step, for both pipelines and both exchange schedules
(overlapped / in-line).  Product against product (single-GPU parity with the CPU checker is what ``pytest -m gpu`` establishes)."""
from __future__ import annotations

import os

import numpy as np


def check(width: int, height: int, steps: int = 2, iterations: int = 37, local: int | None = None, warm: bool = False,
          pipelines=(1, 0), schedules=(True, False), v0_scale: float = 0.8, vorticity: float = 1.0, speed: float | None = None,
          verbose: bool = False, drivers=("native", "python")) -> dict:
    import torch
    import torch.distributed as dist

    from natrix_b200 import _lib as L, workloads as W
    from natrix_b200.core.fluid_simulator import FluidSimulator
    from natrix_b200.slabs import SlabSimulator, SlabSmoothParticlesArea
    from natrix_b200.smooth_particles_area import SmoothParticlesArea

    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0")) if local is None else local
    rng = np.random.default_rng(11)
    v0 = (v0_scale * rng.uniform(-1, 1, (height, width, 2))).astype(np.float32)
    circles = [(0.3, 0.25, 40.0), (0.7, 0.5, 70.0), (0.5, 0.98, 30.0)]     # one straddles a slab boundary for N = 2, 4
    splats = [((0.5, 0.5), (0.9, -0.6), 48.0), ((0.2, 0.74), (-0.5, 0.8), 25.0)]
    mismatches, compared = [], 0
    combos = [(p, o, d) for p in pipelines for o in schedules for d in drivers]
    driver_env = os.environ.get("NATRIX_SLAB_DRIVER")
    for pipeline, overlap, driver in combos:
        for _ in (0,):
            # native: the exchange inside libnatrix_b200.so (natrix_comm_init + natrix_step); python: the model of
            # the same schedule in natrix_b200/slabs.py over torch.distributed
            os.environ["NATRIX_SLAB_DRIVER"] = driver
            ref = FluidSimulator(width, height, None, device=local)
            ref.set_option(L.OPT_PIPELINE, pipeline)
            slab = SlabSimulator(width, height, device=local, depth=8, overlap=overlap)
            assert slab.native == (driver == "native" and world > 1)
            slab.sim.set_option(L.OPT_PIPELINE, pipeline)
            for s in (ref, slab.sim):
                s.vorticity, s.viscosity, s.iterations = vorticity, (0.3 if pipeline else 0.0), iterations
                s.warm_start = warm
                if speed is not None:
                    s.speed = speed
            slab.iterations = iterations
            ref.upload("velocity", v0)
            slab.sim.upload("velocity", v0[slab.row0:slab.row0 + slab.rows])
            pw, ph = (2 * width, 2 * height) if pipeline else (3 * width // 2, 3 * height // 2)
            ref_dye, slab_dye = SmoothParticlesArea(pw, ph, ref), SlabSmoothParticlesArea(pw, ph, slab)
            for d in (ref_dye, slab_dye):
                d.dissipation = 0.98
                if speed is not None:
                    d.speed = speed
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
                pairs = [(name, ref.download(name)[slab.row0:slab.row0 + slab.rows], slab.sim.download(name))
                         for name in ("velocity", "pressure", "divergence", "vorticity")]
                pairs.append((f"dye{pw}x{ph}", ref_dye.download()[slab_dye.row0:slab_dye.row0 + slab_dye.rows],
                              slab_dye.engine.area.download()))
                pairs.append(("dye_rgba8", ref_dye.export_rgba8()[slab_dye.row0:slab_dye.row0 + slab_dye.rows],
                              slab_dye.engine.area.export_rgba8()))
                for name, a, b in pairs:
                    compared += 1
                    same = bool(np.array_equal(a, b))
                    if not same:
                        mismatches.append(f"rank {rank} {driver} pipeline {pipeline} overlap {int(overlap)} step {k} {name}: "
                                          f"max|diff| {float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()):.3e}")
                    if verbose and (not same or k == steps - 1):
                        print(f"[rank {rank}/{world}] {driver} pipeline {pipeline} overlap {int(overlap)} step {k} {name}: "
                              f"bit-identical={same}", flush=True)
            ref_dye.destroy()
            slab_dye.engine.area.destroy()
            ref.destroy()
            slab.sim.destroy()
    os.environ.pop("NATRIX_SLAB_DRIVER", None)
    if driver_env is not None:
        os.environ["NATRIX_SLAB_DRIVER"] = driver_env
    flag = torch.tensor([len(mismatches)], device=f"cuda:{local}")
    dist.all_reduce(flag)
    return {"bit_identical": int(flag.item()) == 0, "grid": [width, height], "world": world, "steps": steps,
            "jacobi_iterations": iterations, "pipelines": list(pipelines), "exchange_schedules": ["overlapped" if o else "in-line" for o in schedules],
            "drivers": list(drivers), "v0_scale": v0_scale, "vorticity": vorticity, "speed": speed,
            "fields": ["velocity", "pressure", "divergence", "vorticity", "dye", "dye_rgba8"],
            "arrays_compared_per_rank": compared, "mismatches_all_ranks": int(flag.item()), "first_mismatches_rank0": mismatches[:4],
            "what": "row-slab run over NCCL vs the single-GPU run of the same grid, every rank compares its own rows"}
