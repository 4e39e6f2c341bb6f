"""Row-slab multi-GPU driver: one process per GPU, halo exchange with torch.distributed.

The reference is single-device; this is new capability.  The global grid
``width x height`` is cut into contiguous row slabs, one per rank.  Every stage of the step is a
local stencil or a bounded gather, so the only data-path communication is a point-to-point
exchange of halo rows with the two neighbouring ranks (NCCL send/recv over NVLink):

    before advect                 velocity     ceil(1.25 dt speed) + 5 rows
    after divergence              divergence + blocked-neighbour mask (one batch)   k T rows   (T = Jacobi depth)
    before every k Jacobi blocks  pressure     k T rows, k = halo // T   (not the first: p starts at zero)
    before gradient subtraction   pressure     1 row

The exchanges inside the Jacobi phase are overlapped with compute: the interior rows of a group of k
blocks need no halo and start at once on the simulator's stream (natrix_step_phase 4) while the exchange
and then the few edge rows (phase 5) run on a second, high-priority stream.

Obstacles and impulses are functions of global cell coordinates: every rank rasterises its own
rows (halo rows included), no exchange.  Halo rows are recomputed redundantly from the same
inputs in the same order, so the result is bit-identical to the single-GPU run.

Two drivers of the same schedule:

* **native** (default on CUDA): the exchange lives inside libnatrix_b200.so - ``natrix_comm_init`` gives the slab
  handle an NCCL communicator and ``natrix_step`` / ``natrix_dye_step`` run the slab's share of the step, halo
  exchanges included, in one C call per step.  This module then only does the rank plumbing (partition, handing
  rank 0's NCCL unique id to the other ranks over torch.distributed).
* **python** (``NATRIX_SLAB_DRIVER=python``, and every non-CUDA engine): ``SlabSimulator.update`` below issues
  the phases and the ``torch.distributed`` send/recv pairs itself.  It is the executable model of the schedule:
  the gloo tests drive it with a CPU engine, and hosts with their own transport follow it.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Tuple

import numpy as np

DEFAULT_HALO = 48          # 6 Jacobi launches of depth 8 per pressure exchange (measured at N = 4, 8: 24..96 within 2 %)


def partition_rows(height: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous (row0, rows) per rank; the first ``height % world`` ranks get one extra row."""
    base, extra = divmod(height, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        out.append((r, n))
        r += n
    return out


class CudaSlabEngine:
    """One slab handle of libnatrix_b200.so plus torch views of its halo regions."""

    def __init__(self, width, height, row0, rows, halo, device):
        import torch

        from natrix_b200 import _lib as L
        from natrix_b200.core.fluid_simulator import FluidSimulator

        self.torch, self.L = torch, L
        self.sim = FluidSimulator(width, height, None, device=device, slab=(row0, rows, halo))
        self.device = device
        self.stream = torch.cuda.ExternalStream(self.sim.cuda_stream, device=device)
        comm = C.c_void_p()
        L.check(self.sim._lib.natrix_comm_stream(self.sim._handle(), C.byref(comm)))
        self.comm_stream = torch.cuda.ExternalStream(comm.value, device=device)
        self._views = {}

    supports_overlap = True          # natrix_step_phase 4 / 5 (interior / edges of a Jacobi group)
    native = False                   # True once init_comm has given the handle its own communicator

    def init_comm(self, dist, group, rank: int, world: int, overlap: bool):
        """Hand rank 0's NCCL unique id to every rank (plumbing) and let the library build its communicator."""
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self.L.check(self.sim._lib.natrix_comm_unique_id(buf))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        os.environ["NATRIX_SLAB_OVERLAP"] = "1" if overlap else "0"       # read by natrix_comm_init
        self.L.check(self.sim._lib.natrix_comm_init(self.sim._handle(), ident, rank, world))
        self.native = True

    def comm_stats(self):
        n, b = C.c_ulonglong(), C.c_ulonglong()
        self.L.check(self.sim._lib.natrix_comm_stats(self.sim._handle(), C.byref(n), C.byref(b)))
        return n.value, b.value

    FIELD_IDS = {"velocity": 0, "pressure": 1, "divergence": 2, "nbmask": 5, "div4": 6}

    def push_params(self):
        self.sim._push_params()

    def phase(self, phase: int, dt: float, sweeps: int = 0):
        if phase == 0:
            self.sim._push_params()
        self.L.check(self.sim._lib.natrix_step_phase(self.sim._handle(), phase, dt, sweeps))

    def rows_needed(self, phase: int, dt: float) -> int:
        return self.L.check(self.sim._lib.natrix_halo_rows_needed(self.sim._handle(), phase, dt))

    def halo_region(self, field: str, side: int, rows: int):
        send, recv, nbytes = C.c_void_p(), C.c_void_p(), C.c_size_t()
        if field == "divergence" and self.sim.get_option(self.L.OPT_PIPELINE) != 0:
            field = "div4"       # the fused pipeline's sweeps read the scaled copy (NATRIX_DIV4), pipeline 0 the divergence
        self.L.check(self.sim._lib.natrix_halo_region(self.sim._handle(), self.FIELD_IDS[field], side, rows,
                                                      C.byref(send), C.byref(recv), C.byref(nbytes)))
        return self._view(send.value, nbytes.value), self._view(recv.value, nbytes.value)

    def _view(self, ptr: int, nbytes: int):
        # the fields ping-pong between two buffers, so the same few (pointer, size) pairs come back every
        # step: building a tensor view costs more host time than the exchange itself
        key = (ptr, nbytes)
        t = self._views.get(key)
        if t is None:
            class _Raw:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                            "strides": None}
            t = self._views[key] = self.torch.as_tensor(_Raw(), device=f"cuda:{self.device}")
        return t

    def stream_context(self, comm: bool = False):
        return self.torch.cuda.stream(self.comm_stream if comm else self.stream)


class SlabSimulator:
    """The reference's FluidSimulator surface for one rank's slab of a global grid."""

    def __init__(self, width: int, height: int, engine=None, halo: Optional[int] = None, device: Optional[int] = None,
                 group=None, depth: int = 8, overlap: Optional[bool] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.width, self.height = int(width), int(height)
        self.row0, self.rows = partition_rows(self.height, self.world)[self.rank]
        self.halo = int(os.environ.get("NATRIX_SLAB_HALO", DEFAULT_HALO)) if halo is None else int(halo)
        if self.rows < self.halo:
            raise ValueError(f"slab of {self.rows} rows is shorter than its halo ({self.halo})")
        if engine is None:
            dev = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
            engine = CudaSlabEngine(self.width, self.height, self.row0, self.rows, self.halo, dev)
        self.engine = engine
        self.depth = int(depth)
        # overlap the pressure exchanges with the interior Jacobi launches when the engine can split a group
        can = bool(getattr(engine, "supports_overlap", False))
        if overlap is None:
            overlap = os.environ.get("NATRIX_SLAB_OVERLAP", "1") != "0"      # A/B switch for measurements
        self.overlap = bool(overlap) and can
        self.iterations = 50
        self.simulate = True
        self._exchanges = 0
        self._exchanged_bytes = 0
        # the library's own exchange (one natrix_step per update) unless the Python model of it is asked for
        self.native = False
        if self.world > 1 and hasattr(engine, "init_comm") and os.environ.get("NATRIX_SLAB_DRIVER", "native") != "python":
            engine.init_comm(dist, group, self.rank, self.world, self.overlap)
            self.native = True

    @property
    def exchanges(self) -> int:
        return self.engine.comm_stats()[0] if self.native else self._exchanges

    @property
    def exchanged_bytes(self) -> int:
        """Bytes sent to neighbours so far (per exchange: all fields, both neighbours)."""
        return self.engine.comm_stats()[1] if self.native else self._exchanged_bytes

    # -- the reference's mutators, forwarded to the slab (global normalised coordinates)
    @property
    def sim(self):
        return self.engine.sim

    def add_velocity(self, position, velocity, radius):
        if self.simulate:
            self.sim.add_velocity(position, velocity, radius)

    def add_circle_obstacle(self, position, radius, static=False):
        if self.simulate:
            self.sim.add_circle_obstacle(position, radius, static)

    def add_triangle_obstacle(self, p1, p2, p3, static=False):
        if self.simulate:
            self.sim.add_triangle_obstacle(p1, p2, p3, static)

    # -- halo exchange with the two neighbouring ranks
    def exchange(self, fields, rows: int, comm: bool = False, region=None, limit: Optional[int] = None):
        """Swap `rows` halo rows of one field (or of several, in one batched NCCL group) with both neighbours,
        on the simulator's stream or (comm=True) on its second stream.  `region(field, side, rows)` returns
        the (send, recv) buffers; the default asks the simulator's engine."""
        if isinstance(fields, str):
            fields = (fields,)
        if rows <= 0 or self.world == 1:
            return
        limit = self.halo if limit is None else limit
        if rows > limit:
            raise ValueError(f"{'+'.join(fields)}: step needs {rows} halo rows but the slab was created with {limit}")
        region = region or self.engine.halo_region
        dist = self.dist
        ops, keep = [], []
        with (self.engine.stream_context(True) if comm else self.engine.stream_context()):
            for peer, side in ((self.rank - 1, 0), (self.rank + 1, 1)):
                if peer < 0 or peer >= self.world:
                    continue
                for field in fields:
                    send, recv = region(field, side, rows)
                    ops += [dist.P2POp(dist.isend, send, peer, self.group), dist.P2POp(dist.irecv, recv, peer, self.group)]
                    keep += [send, recv]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self._exchanges += 1
        self._exchanged_bytes += sum(t.numel() * t.element_size() for t in keep) // 2

    # -- one step (ref: FluidSimulator.update, fluid_simulator.py:174-280) in four phases
    def update(self, time_delta: float):
        if not self.simulate:
            return
        e = self.engine
        if self.native:
            self.sim.iterations = int(self.iterations)
            self.sim.update(time_delta)                       # natrix_step: the whole step, exchanges included
            return
        if hasattr(e, "push_params"):
            e.push_params()                                   # rows_needed reads the simulator's CURRENT speed
        self.exchange("velocity", e.rows_needed(0, time_delta))
        e.phase(0, time_delta)
        e.phase(1, time_delta)
        # Several Jacobi launches per exchange: with k * depth halo rows the slab recomputes the rows its
        # neighbour owns for the first k - 1 launches (a few rows) instead of exchanging after every one.
        # A partial group goes first so that launch depths never decrease (natrix_step_phase 4 / 5).
        n = int(self.iterations)
        span = max(1, self.halo // self.depth) * self.depth
        groups = ([n % span] if n % span else []) + [span] * (n // span)
        overlap = self.overlap and self.world > 1 and self.rows >= 2 * span
        for i, t in enumerate(groups):
            if overlap:
                e.phase(4, time_delta, t)                     # interior rows: no halo needed
            if i == 0:
                # p starts at zero, halos included - unless the simulator warm-starts from the last step's pressure
                warm = bool(getattr(self.sim, "warm_start", False))
                self.exchange(("divergence", "nbmask") + (("pressure",) if warm else ()), min(span, n), comm=overlap)
            else:
                self.exchange("pressure", t, comm=overlap)
            e.phase(5 if overlap else 2, time_delta, t)       # edge rows (or all rows) after the exchange
        self.exchange("pressure", 1)
        e.phase(3, time_delta)


def velocity_rows_for_dye(dye_height: int, grid_height: int, world: int) -> int:
    """Rows of post-projection velocity a dye slab samples beyond its simulator slab, maximum over ranks
    (every rank must exchange the same count).  Mirrors natrix_dye_halo_rows_needed(dye, 0, ..): the
    shader's float32 ``(y / dye_height) * grid_height`` at the first and last dye row of each slab
    (ref: demo/shaders/shader.AdvectParticle.comp:46)."""
    f32 = np.float32
    need = 0
    for (p0, pn), (v0, vn) in zip(partition_rows(dye_height, world), partition_rows(grid_height, world)):
        lo = (f32(p0) / f32(dye_height)) * f32(grid_height)
        hi = (f32(p0 + pn - 1) / f32(dye_height)) * f32(grid_height)
        first, last = max(0, int(math.floor(lo))), min(grid_height - 1, int(math.ceil(hi)))
        need = max(need, v0 - first, last - (v0 + vn - 1))
    return need


class CudaDyeSlabEngine:
    """One dye slab handle of libnatrix_b200.so attached to a CudaSlabEngine's simulator slab."""

    def __init__(self, width, height, row0, rows, halo, sim_engine: CudaSlabEngine):
        from natrix_b200.smooth_particles_area import SmoothParticlesArea

        self.e = sim_engine
        self.area = SmoothParticlesArea(width, height, sim_engine.sim, None, slab=(row0, rows, halo))

    def add(self, position, radius, strength):
        self.area.add_particles(position, radius, strength)

    def rows_needed(self, which: int, dt: float, speed: float) -> int:
        L = self.e.L
        return L.check(self.area._lib.natrix_dye_halo_rows_needed(self.area._handle(), which, dt, speed))

    def halo_region(self, field, side: int, rows: int):
        send, recv, nbytes = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self.e.L.check(self.area._lib.natrix_dye_halo_region(self.area._handle(), side, rows, C.byref(send),
                                                             C.byref(recv), C.byref(nbytes)))
        return self.e._view(send.value, nbytes.value), self.e._view(recv.value, nbytes.value)

    def step(self, dt: float, speed: float, dissipation: float):
        self.e.L.check(self.area._lib.natrix_dye_step(self.area._handle(), dt, speed, dissipation))


class SlabSmoothParticlesArea:
    """The reference's SmoothParticlesArea surface (demo/smooth_particles_area.py:15-211) for one rank's slab
    of a global dye grid, cut into the same normalised-y slabs as the simulator.

    Per update: the post-projection velocity rows the dye samples beyond the simulator slab and the dye rows
    within back-trace reach are exchanged with the two neighbours, then the slab's own rows are advected.
    add_particles is a function of global coordinates: every rank applies it to the rows it holds."""

    def __init__(self, width: int, height: int, fluid_simulation: SlabSimulator, vertex_layout=None,
                 halo: Optional[int] = None, engine=None):
        self.slab = fluid_simulation
        self.width, self.height = int(width), int(height)
        self.row0, self.rows = partition_rows(self.height, self.slab.world)[self.slab.rank]
        ratio = self.height / self.slab.height
        self.halo = int(math.ceil(self.slab.halo * max(1.0, ratio))) if halo is None else int(halo)
        if self.rows < self.halo:
            raise ValueError(f"dye slab of {self.rows} rows is shorter than its halo ({self.halo})")
        self.velocity_rows = velocity_rows_for_dye(self.height, self.slab.height, self.slab.world)
        if engine is None:
            engine = CudaDyeSlabEngine(self.width, self.height, self.row0, self.rows, self.halo, self.slab.engine)
        self.engine = engine
        self._speed, self._dissipation = 500.0, 1.0
        self.simulate = True

    @property
    def speed(self):
        return self._speed

    @speed.setter
    def speed(self, value):
        if value > 0:
            self._speed = value
        else:
            raise ValueError("'Speed' should be greater than zero")

    @property
    def dissipation(self):
        return self._dissipation

    @dissipation.setter
    def dissipation(self, value):
        if value > 0:
            self._dissipation = value
        else:
            raise ValueError("'Dissipation' should be grater than zero")

    def add_particles(self, position, radius, strength):
        if self.simulate:
            self.engine.add(position, radius, strength)

    def update(self, time_delta: float):
        if not self.simulate:
            return
        if self.slab.native and hasattr(self.engine, "area"):
            self.engine.step(time_delta, self._speed, self._dissipation)      # natrix_dye_step exchanges by itself
            return
        self.slab.exchange("velocity", self.velocity_rows)
        rows = self.engine.rows_needed(1, time_delta, self._speed)
        self.slab.exchange("dye", rows, region=self.engine.halo_region, limit=self.halo)
        self.engine.step(time_delta, self._speed, self._dissipation)


# ----------------------------------------------------------------------------------------- bench
def strong_scaling_point(args, w, local: int, depth: int, metric: str) -> dict:
    """Workload `w` (one fixed grid) on all ranks as row slabs, and on one GPU alone; device-timed, max over ranks."""
    import torch
    import torch.distributed as dist

    from natrix_b200 import _lib as L
    from natrix_b200 import workloads as W
    from natrix_b200.core.fluid_simulator import FluidSimulator

    import time

    world = dist.get_world_size()
    steps = max(3, min(args.steps, 10))

    def timed(sim, stream, one_step, synchronize):
        for k in range(3):
            one_step(k)
        synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        t0 = time.perf_counter()
        for k in range(steps):
            one_step(3 + k)
        host = 1e3 * (time.perf_counter() - t0) / steps
        synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=f"cuda:{local}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), host

    slab = SlabSimulator(w.width, w.height, device=local, depth=depth)
    W.configure(slab.sim, w)
    slab.iterations = w.iterations
    slab.sim.set_option(L.OPT_JACOBI_DEPTH, depth)
    slab.sim.upload("velocity", W.smooth_velocity(w.width, w.height, slab.row0, slab.rows))
    dye = None
    if w.dye_size:
        dye = SlabSmoothParticlesArea(w.dye_size[0], w.dye_size[1], slab)
        dye.dissipation = w.dye_dissipation

    def slab_step(k):
        for (px, py, r) in W.circles_at(w, k):
            slab.add_circle_obstacle((px, py), r)
        slab.update(W.DT)
        if dye is not None:
            dye.update(W.DT)
        for (px, py, vx, vy) in W.orbit_positions(w, k):
            slab.add_velocity((px, py), (vx, vy), w.splat_radius)
            if dye is not None:
                dye.add_particles((px, py), w.dye_radius, w.dye_strength)

    ms_n, host_n = timed(slab.sim, slab.engine.stream, slab_step, slab.sim.synchronize)
    native = slab.native
    slab.sim.destroy()
    from natrix_b200.smooth_particles_area import SmoothParticlesArea
    solo, solo_dye = W.build(w, FluidSimulator, SmoothParticlesArea if w.dye_size else None, device=local)
    solo.set_option(L.OPT_JACOBI_DEPTH, depth)
    ms_1, _ = timed(solo, torch.cuda.ExternalStream(solo.cuda_stream, device=local), lambda k: W.run_step(w, solo, solo_dye, k),
                    solo.synchronize)
    solo.destroy()
    v_n, v_1 = w.cells / (ms_n * 1e-3) / 1e6, w.cells / (ms_1 * 1e-3) / 1e6
    return {"workload": w.name, "grid": [w.width, w.height], "dye_grid": list(w.dye_size) if w.dye_size else None,
            "jacobi_iterations": w.iterations, "scaling": "strong",
            "n_gpus": world, "value": v_n, "unit": metric, "ms_per_step": ms_n, "host_enqueue_ms_per_step": host_n,
            "one_gpu": {"value": v_1, "ms_per_step": ms_1, "note": "the same grid on one GPU of this box (every rank, max)"},
            "efficiency": v_n / (world * v_1), "steps": steps, "driver": "native" if native else "python"}


def _slab_roofline(w, world, depth, jacobi_ms, jacobi_bytes, peak):
    """Per-GPU roofline of the Jacobi phase of one step.  achieved / frac are PHYSICAL: the 13 B per cell a launch must
    move (p 4 + scaled divergence 4 + mask 1 read, p 4 written) x launches over the phase's duration (max over ranks,
    halo exchanges and recomputed halo rows included), against the measured HBM peak; algorithmic_* is SURVEY 8(d)'s
    20 B x cells x sweeps over the same time (it exceeds the peak by design: temporal blocking)."""
    launches = -(-w.iterations // depth)
    cells = w.cells / world
    phys = 13 * cells * launches / (jacobi_ms * 1e-3) / 1e9
    algo = jacobi_bytes * cells * w.iterations / (jacobi_ms * 1e-3) / 1e9
    return {"kernel": "k_jacobi_tb (+ pressure halo exchanges)", "bound": "hbm", "achieved": phys, "peak": peak,
            "unit": "GB/s per GPU", "frac": phys / peak, "traffic": None, "min_bytes_per_launch": 13 * cells,
            "launches_per_step": launches, "jacobi_ms_per_step": jacobi_ms, "algorithmic_achieved": algo,
            "algorithmic_frac": algo / peak,
            "note": "physical: 13 B per cell per launch (the bytes a launch must move; ncu DRAM traffic is in the N = 1 line) "
                    "over the Jacobi phase of one step, max over ranks, halo exchanges included; algorithmic_*: 20 B x "
                    "cells x sweeps (SURVEY 8(d)), above the peak by design"}


def run_bench(args, w, metric, jacobi_bytes, measured_peak_gbs, ClockSampler):
    """bench.py --gpus N under torchrun: weak scaling, one 32768 x 4096 slab per rank."""
    import json
    import time

    import torch
    import torch.distributed as dist

    from natrix_b200 import _lib as L
    from natrix_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    depth = args.depth or 8
    # driver-visible multi-GPU parity: a small slab-vs-single-GPU run, both pipelines and both exchange schedules,
    # before anything is timed; the outcome rides in the JSON line
    from natrix_b200 import slab_parity
    parity = slab_parity.check(2048, 512 * world, steps=2, iterations=37, local=local)
    slab = SlabSimulator(w.width, w.height, device=local, depth=depth)
    sim = slab.sim
    sim.vorticity, sim.viscosity, sim.iterations = w.vorticity, w.viscosity, w.iterations
    slab.iterations = w.iterations
    if args.pipeline is not None:
        sim.set_option(L.OPT_PIPELINE, args.pipeline)
    sim.set_option(L.OPT_JACOBI_DEPTH, depth)
    sim.set_option(L.OPT_TIMING, 1)
    sim.upload("velocity", W.smooth_velocity(w.width, w.height, slab.row0, slab.rows))

    dye = None
    if w.dye_size:
        dye = SlabSmoothParticlesArea(w.dye_size[0], w.dye_size[1], slab)
        dye.dissipation = w.dye_dissipation

    def one_step(k):                                  # the demo's call order, as workloads.run_step
        for (px, py, r) in W.circles_at(w, k):
            slab.add_circle_obstacle((px, py), r)
        slab.update(W.DT)
        if dye is not None:
            dye.update(W.DT)
        for (px, py, vx, vy) in W.orbit_positions(w, k):
            slab.add_velocity((px, py), (vx, vy), w.splat_radius)
            if dye is not None:
                dye.add_particles((px, py), w.dye_radius, w.dye_strength)

    stream = slab.engine.stream
    step = 0
    for _ in range(max(args.warmup, 3)):
        one_step(step); step += 1
    sim.stats("velocity")        # warm-up covers the e2e leg's readback kernel too (CUDA loads a kernel on its first launch)
    sim.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    h0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(step); step += 1
    host_ms = 1e3 * (time.perf_counter() - h0) / args.steps      # host time to enqueue one step (not a result)
    sim.synchronize()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = sim.launch_count - launches0
    jac = torch.tensor([sim.timings()["jacobi"]], device=f"cuda:{local}")      # last timed step, incl. halo exchanges
    dist.all_reduce(jac, op=dist.ReduceOp.MAX)
    jacobi_ms = float(jac.item())
    ms_per_step = total_ms / args.steps
    value = w.cells / (ms_per_step * 1e-3) / 1e6

    # e2e: per-step host sync and readback of a velocity statistic on every rank
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(step); step += 1
        sim.stats("velocity")
    dist.barrier()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0) / args.steps], device=f"cuda:{local}")
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None

    # the 1-GPU point of the same weak-scaling series, measured on this box: every rank runs the
    # 32768 x 4096 grid standalone (no exchange); max over ranks
    from natrix_b200.core.fluid_simulator import FluidSimulator
    w1 = W.cfg5_workload(1, width=w.width, rows_per_gpu=slab.rows)
    solo, _ = W.build(w1, FluidSimulator, None, device=local)
    solo.set_option(L.OPT_JACOBI_DEPTH, depth)
    sstream = torch.cuda.ExternalStream(solo.cuda_stream, device=local)
    for k in range(2):
        W.run_step(w1, solo, None, k)
    solo.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = max(3, min(args.steps, 5))
    b0.record(sstream)
    for k in range(nb):
        W.run_step(w1, solo, None, 2 + k)
    solo.synchronize()
    b1.record(sstream)
    torch.cuda.synchronize()
    bms = torch.tensor([b0.elapsed_time(b1) / nb], device=f"cuda:{local}")
    dist.all_reduce(bms, op=dist.ReduceOp.MAX)
    base_value = w1.cells / (float(bms.item()) * 1e-3) / 1e6
    solo.destroy()

    # config 4 (BASELINE.json): 16384^2 STRONG scaling - the same grid cut into `world` slabs, against the same grid
    # on one GPU of this box (every rank runs it standalone, max over ranks), measured in this run
    strong = dye_point = None
    n_exchanges, n_bytes = slab.exchanges, slab.exchanged_bytes
    if getattr(args, "workload", "auto") == "auto":
        slab.sim.destroy()
        strong = strong_scaling_point(args, W.cfg4_workload(16384), local, depth, metric)
        # ... and config 3 (4096^2 velocity + 4096^2 dye, splats) on the same slabs: the dye field's own exchange
        dye_point = strong_scaling_point(args, W.cfg3_workload(4096), local, depth, metric)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        algo = jacobi_bytes * w.cells * w.iterations + (116 if w.viscosity == 0 else 132) * w.cells
        line = {
            "metric": metric, "value": value, "unit": metric, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if "weak" in w.name else "strong",       # only config 5 grows with N
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "grid": [w.width, w.height], "per_gpu_grid": [w.width, slab.rows],
                       "jacobi_iterations": w.iterations, "obstacles_per_step": len(w.circles),
                       "dye_grid": list(w.dye_size) if w.dye_size else None,
                       "parallelism": f"row-slabs x{world}, halo {slab.halo} rows, NCCL send/recv"
                                      + (", pressure exchanges overlapped with interior Jacobi" if slab.overlap else ""),
                       "jacobi_depth": depth, "l2": f"per-GPU state {38 * w.width * slab.rows / 1e9:.1f} GB exceeds the 126 MB L2; no flush needed",
                       "algorithmic_GBps_per_gpu": algo / world / (ms_per_step * 1e-3) / 1e9},
            "weak_base": {"workload": w1.name, "n_gpus": 1, "value": base_value, "ms_per_step": float(bms.item()),
                          "note": "same per-GPU slab run standalone on every rank of this box (max over ranks)"},
            "slab_parity": parity, "config4_strong": strong, "config3_dye_slabs": dye_point,
            "halo": {"driver": "libnatrix_b200.so (natrix_step: NCCL send/recv from C)" if slab.native else "python (torch.distributed)",
                     "exchanges_per_step": n_exchanges / (2 * args.steps + max(args.warmup, 3)),
                     "bytes_per_exchange": n_bytes / max(n_exchanges, 1),
                     "host_enqueue_ms_per_step": host_ms},
            "e2e": {"value": w.cells / (float(e2e_ms.item()) * 1e-3) / 1e6, "unit": metric,
                    "ms_per_step": float(e2e_ms.item()), "h2d_bytes_per_step": 16 * len(w.circles) + 32,
                    "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None,
            "roofline": _slab_roofline(w, world, depth, jacobi_ms, jacobi_bytes, peak),
            "peak_hbm_gbs": peak, "peak_source": peak_src,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0
