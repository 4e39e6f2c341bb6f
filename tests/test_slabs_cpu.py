"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic in natrix_b200/slabs.py: row
partitioning, halo-exchange plumbing and the per-step exchange schedule of SlabSimulator.update.

The engine here is a NumPy stand-in built from the oracle's translation-invariant stages
(divergence, Jacobi sweep, gradient subtraction), so a slab run must reproduce the single-process
oracle bit for bit; the CUDA engine is checked the same way on real GPUs by
tests/test_gpu_multi.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from natrix_b200.slabs import SlabSimulator, partition_rows  # noqa: E402
from oracle import natrix_oracle as O  # noqa: E402

W, H, ITER, DEPTH, HALO = 48, 61, 19, 4, 9        # halo 9: two Jacobi blocks (8 sweeps) per pressure exchange


def test_partition_rows_covers_the_grid():
    for h, n in [(4096, 8), (61, 3), (7, 7), (10, 4)]:
        parts = partition_rows(h, n)
        assert parts[0][0] == 0 and sum(r for _, r in parts) == h
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(n - 1))
        assert max(r for _, r in parts) - min(r for _, r in parts) <= 1


class NumpySlabEngine:
    """Slab of the projection half of the step (divergence -> N Jacobi sweeps -> gradient) on arrays
    with `halo` extra rows where a neighbour exists."""

    def __init__(self, vel, obs, row0, rows, halo, height, overlap=False):
        self.supports_overlap = overlap
        self._interior = None
        self.row0, self.rows, self.height = row0, rows, height
        self.top = halo if row0 > 0 else 0                    # halo rows actually present above / below
        self.bot = halo if row0 + rows < height else 0
        n = self.top + rows + self.bot
        self.vel = np.zeros((n, W, 2), np.float32)
        self.vel[self.top:self.top + rows] = vel[row0:row0 + rows]
        lo, hi = row0 - self.top, row0 + rows + self.bot
        self.obs = obs[lo:hi].copy()                          # obstacles: rasterised locally, no exchange
        self.p = np.zeros((n, W), np.float32)
        self.div = np.zeros((n, W), np.float32)
        self.sim = self                                       # SlabSimulator forwards mutators to .sim

    # -- interface used by SlabSimulator
    def rows_needed(self, phase, dt):
        return {0: 1, 1: 0, 2: DEPTH, 3: 1}[phase]

    def stream_context(self, comm=False):
        import contextlib
        return contextlib.nullcontext()

    def _arr(self, field):
        return {"velocity": self.vel, "pressure": self.p, "divergence": self.div, "nbmask": None}[field]

    def halo_region(self, field, side, rows):
        a = self._arr(field)
        if a is None:                                        # the CPU engine keeps no mask: exchange a dummy
            t = torch.zeros(rows * W, dtype=torch.uint8)
            return t, t.clone()
        if side == 0:
            send, recv = a[self.top:self.top + rows], a[self.top - rows:self.top]
        else:
            end = self.top + self.rows
            send, recv = a[end - rows:end], a[end:end + rows]
        return torch.from_numpy(send), torch.from_numpy(recv)

    def phase(self, phase, dt, sweeps=0):
        if phase == 1:
            self.div = O.divergence(self.vel, self.obs)      # valid on the slab's rows (needs velocity +-1)
            self.p[...] = 0
        elif phase == 2:
            for _ in range(sweeps):
                self.p = O.poisson_sweep(self.p, self.div, self.obs)
        elif phase == 4:
            # interior of a group, BEFORE the exchange: rows at least `sweeps` away from a neighbour's rows
            assert self._interior is None
            q = self.p.copy()
            for _ in range(sweeps):
                q = O.poisson_sweep(q, self.div, self.obs)
            lo = self.top + (sweeps if self.top else 0)
            hi = self.top + self.rows - (sweeps if self.bot else 0)
            self._interior = (sweeps, lo, hi, q[lo:hi].copy())
        elif phase == 5:
            t, lo, hi, rows = self._interior
            assert t == sweeps
            self._interior = None
            self.phase(2, dt, sweeps)                        # after the exchange: every row
            assert np.array_equal(self.p[lo:hi], rows), "interior rows depend on the halo"
        elif phase == 3:
            self.vel = O.subtract_gradient(self.vel, self.p, self.obs)

    def own(self, a):
        return a[self.top:self.top + self.rows]


def _reference():
    rng = np.random.default_rng(5)
    vel = rng.uniform(-0.7, 0.7, (H, W, 2)).astype(np.float32)
    obs = np.zeros((H, W, 2), np.float32)
    O.add_circle_obstacle(obs, (0.4, 0.5), 9.0)
    O.add_triangle_obstacle(obs, (0.6, 0.1), (0.9, 0.3), (0.7, 0.9))
    div = O.divergence(vel, obs)
    p = np.zeros((H, W), np.float32)
    for _ in range(ITER):
        p = O.poisson_sweep(p, div, obs)
    out = O.subtract_gradient(vel, p, obs)
    return vel, obs, div, p, out


def _worker(rank, world, port, overlap, errors):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        vel, obs, div, p, out = _reference()
        row0, rows = partition_rows(H, world)[rank]
        eng = NumpySlabEngine(vel, obs, row0, rows, HALO, H, overlap)
        slab = SlabSimulator(W, H, engine=eng, halo=HALO, depth=DEPTH)
        slab.iterations = ITER
        assert slab.overlap == overlap
        assert (slab.row0, slab.rows) == (row0, rows)
        slab.update(1.0 / 60.0)
        assert np.array_equal(eng.own(eng.div), div[row0:row0 + rows]), "divergence"
        assert np.array_equal(eng.own(eng.p), p[row0:row0 + rows]), "pressure"
        assert np.array_equal(eng.own(eng.vel), out[row0:row0 + rows]), "velocity"
        # schedule: velocity, divergence + mask (one batch), ceil(N / span) - 1 pressure exchanges, final pressure row
        span = (HALO // DEPTH) * DEPTH
        assert slab.exchanges == 2 + (-(-ITER // span) - 1) + 1, slab.exchanges
        with pytest.raises(ValueError):
            slab.exchange("pressure", HALO + 1)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001 - reported to the parent
        import traceback
        errors.put(f"rank {rank}: {e}\n{traceback.format_exc()}")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,overlap", [(2, False), (3, False), (2, True), (3, True)])
def test_slab_driver_matches_single_process_oracle_over_gloo(world, overlap):
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, overlap, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("a rank hung")
        elif p.exitcode != 0:
            msgs.append(f"exit code {p.exitcode}")
    assert not msgs, "\n".join(msgs)


# ------------------------------------------------------------------------------------------------ dye slabs
PW, PH, VW, VH, DHALO, DSTEPS = 40, 90, 24, 45, 15, 3
DYE_DT, DYE_SPEED = 1.0 / 60.0, 300.0


class NumpyDyeRig:
    """Simulator + dye stand-ins for SlabSmoothParticlesArea.  Every array has the GLOBAL shape, but only the
    rows a rank holds carry data - the rest is poison (NaN dye, 1e6 velocity: a NaN velocity has no back-trace
    cell), so any read outside the exchanged halos ruins the result.  Stages are the oracle's own functions
    evaluated on those arrays."""

    def __init__(self, rank, world, vel):
        self.v0, self.vn = partition_rows(VH, world)[rank]
        self.p0, self.pn = partition_rows(PH, world)[rank]
        self.vel = np.full((VH, VW, 2), 1e6, np.float32)
        self.vel[self.v0:self.v0 + self.vn] = vel[self.v0:self.v0 + self.vn]
        self.obs = np.zeros((VH, VW, 2), np.float32)
        self.dye = np.full((PH, PW), np.nan, np.float32)
        self.dye[max(0, self.p0 - DHALO):self.p0 + self.pn + DHALO] = 0.0
        self.sim = self

    # -- simulator engine surface used by SlabSimulator.exchange
    def stream_context(self, comm=False):
        import contextlib
        return contextlib.nullcontext()

    def halo_region(self, field, side, rows):
        a, r0, rn = (self.vel, self.v0, self.vn) if field == "velocity" else (self.dye, self.p0, self.pn)
        if side == 0:
            send, recv = a[r0:r0 + rows], a[r0 - rows:r0]
        else:
            send, recv = a[r0 + rn - rows:r0 + rn], a[r0 + rn:r0 + rn + rows]
        return torch.from_numpy(send), torch.from_numpy(recv)

    # -- dye engine surface
    def add(self, position, radius, strength):
        with np.errstate(invalid="ignore"):
            self.dye = O.add_particles(self.dye, position, radius, strength)

    def rows_needed(self, which, dt, speed):
        assert which == 1
        return int(np.ceil(1.25 * dt * speed * PH / VH)) + 2

    def step(self, dt, speed, dissipation):
        with np.errstate(invalid="ignore"):
            out = O.advect_particles(self.dye, self.vel, self.obs, dt, speed, dissipation)
        own = out[self.p0:self.p0 + self.pn]
        assert not np.isnan(own).any(), "the dye step read rows that were never exchanged"
        self.dye = np.full_like(out, np.nan)
        self.dye[self.p0:self.p0 + self.pn] = own
        lo, hi = max(0, self.p0 - DHALO), min(PH, self.p0 + self.pn + DHALO)
        self.dye[lo:self.p0] = 0.0                      # stale halo rows: finite garbage, refreshed by the next exchange
        self.dye[self.p0 + self.pn:hi] = 0.0
        self.vel[:self.v0] = 1e6                        # the velocity halo is only valid for this step
        self.vel[self.v0 + self.vn:] = 1e6


def _dye_reference():
    rng = np.random.default_rng(9)
    vel = rng.uniform(-1.0, 1.0, (VH, VW, 2)).astype(np.float32)
    obs = np.zeros((VH, VW, 2), np.float32)
    dye = np.zeros((PH, PW), np.float32)
    frames = []
    for k in range(DSTEPS):
        dye = O.add_particles(dye, (0.5, 0.2 + 0.3 * k), 9.0, 0.7)
        dye = O.add_particles(dye, (0.3, 0.5), 14.0, 0.4)
        dye = O.advect_particles(dye, vel, obs, DYE_DT, DYE_SPEED, 0.97)
        frames.append(dye)
    return vel, frames


def _dye_worker(rank, world, port, errors):
    try:
        from natrix_b200.slabs import SlabSmoothParticlesArea
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        vel, frames = _dye_reference()
        rig = NumpyDyeRig(rank, world, vel)
        slab = SlabSimulator(VW, VH, engine=rig, halo=6, depth=2)
        area = SlabSmoothParticlesArea(PW, PH, slab, halo=DHALO, engine=rig)
        area.speed, area.dissipation = DYE_SPEED, 0.97
        assert (area.row0, area.rows) == (rig.p0, rig.pn) and area.velocity_rows >= 1
        for k in range(DSTEPS):
            area.add_particles((0.5, 0.2 + 0.3 * k), 9.0, 0.7)
            area.add_particles((0.3, 0.5), 14.0, 0.4)
            area.update(DYE_DT)
            assert np.array_equal(rig.dye[rig.p0:rig.p0 + rig.pn], frames[k][rig.p0:rig.p0 + rig.pn]), f"dye, step {k}"
        with pytest.raises(ValueError):
            slab.exchange("dye", DHALO + 1, region=rig.halo_region, limit=DHALO)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001 - reported to the parent
        import traceback
        errors.put(f"rank {rank}: {e}\n{traceback.format_exc()}")


def _run_world(target, world, extra=()):
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, *extra, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("a rank hung")
        elif p.exitcode != 0:
            msgs.append(f"exit code {p.exitcode}")
    assert not msgs, "\n".join(msgs)


@pytest.mark.parametrize("world", [2, 3])
def test_dye_slab_driver_matches_single_process_oracle_over_gloo(world):
    _run_world(_dye_worker, world)


def test_velocity_rows_for_dye():
    from natrix_b200.slabs import velocity_rows_for_dye
    assert velocity_rows_for_dye(4096, 4096, 8) == 0          # same grid: a dye row samples exactly its own velocity row
    assert velocity_rows_for_dye(720, 360, 2) == 1            # the demo's 2x dye grid: one row below (ceil of x.5)
    assert 1 <= velocity_rows_for_dye(1000, 360, 3) <= 2


# ------------------------------------------------------------------------------------------ the whole step
FW, FH, FHALO, FDEPTH, FITER, FSTEPS, FSPEED = 40, 66, 10, 2, 13, 3, 120.0


class NumpyFullStepRig:
    """Every phase of SlabSimulator.update on an OracleFluidSimulator over the GLOBAL grid whose arrays only carry
    data in the rows this rank holds (own rows + exchanged halos); everything else is poison (1e6 velocity,
    NaN pressure / divergence), so a missing or too-short exchange ruins the rows that are compared."""

    supports_overlap = False

    def __init__(self, rank, world, v0):
        self.r0, self.rn = partition_rows(FH, world)[rank]
        s = self.sim = O.OracleFluidSimulator(FW, FH, None)
        s.speed, s.vorticity, s.viscosity, s.iterations = FSPEED, 1.5, 0.4, FITER
        v = np.full((FH, FW, 2), 1e6, np.float32)
        v[self.r0:self.r0 + self.rn] = v0[self.r0:self.r0 + self.rn]
        s.velocity = v

    def _poison(self, a, value):
        a[:self.r0] = value
        a[self.r0 + self.rn:] = value

    def rows_needed(self, phase, dt):
        return {0: int(np.ceil(1.25 * dt * FSPEED)) + 5, 1: 0, 2: FDEPTH, 3: 1}[phase]

    def stream_context(self, comm=False):
        import contextlib
        return contextlib.nullcontext()

    def halo_region(self, field, side, rows):
        s = self.sim
        a = {"velocity": s.velocity, "pressure": s.pressure, "divergence": s.divergence, "nbmask": None}[field]
        if a is None:
            t = torch.zeros(rows * FW, dtype=torch.uint8)
            return t, t.clone()
        if side == 0:
            send, recv = a[self.r0:self.r0 + rows], a[self.r0 - rows:self.r0]
        else:
            end = self.r0 + self.rn
            send, recv = a[end - rows:end], a[end:end + rows]
        return torch.from_numpy(send), torch.from_numpy(recv)

    def phase(self, phase, dt, sweeps=0):
        s = self.sim
        with np.errstate(invalid="ignore", over="ignore"):
            if phase == 0:                                     # fluid_simulator.py:181-228
                if s.has_borders:
                    O.init_boundaries(s._vel[s.VELOCITY_READ])
                s._vel[s.VELOCITY_WRITE] = O.advect_velocity(s.velocity, s.obstacles, np.float32(dt), s.speed, s.dissipation)
                s._flip_v()
                s.vorticity_field = O.calc_vorticity(s.velocity)
                s._vel[s.VELOCITY_WRITE] = O.apply_vorticity(s.velocity, s.vorticity_field, np.float32(dt), s.vorticity)
                s._flip_v()
                alpha, rbeta = O.viscosity_alpha_rbeta(s.viscosity)
                s._vel[s.VELOCITY_WRITE] = O.viscosity_sweep(s.velocity, alpha, rbeta)
                s._flip_v()
            elif phase == 1:                                   # :231-248
                s.divergence = O.divergence(s.velocity, s.obstacles)
                self._poison(s.divergence, np.nan)
                s._p[s.PRESSURE_READ] = np.zeros_like(s._p[0])
            elif phase == 2:                                   # :251-255
                for _ in range(sweeps):
                    s._p[s.PRESSURE_WRITE] = O.poisson_sweep(s.pressure, s.divergence, s.obstacles)
                    s._flip_p()
            elif phase == 3:                                   # :258-280
                s._vel[s.VELOCITY_WRITE] = O.subtract_gradient(s.velocity, s.pressure, s.obstacles)
                s._flip_v()
                s.obstacles = np.zeros_like(s.obstacles)
                self._poison(s._vel[s.VELOCITY_READ], 1e6)
                self._poison(s._p[s.PRESSURE_READ], np.nan)

    def own(self, a):
        return a[self.r0:self.r0 + self.rn]


def _full_step_script(sim, k):
    sim.add_circle_obstacle((0.45, 0.5), 6.0)                   # straddles the slab boundaries
    sim.add_triangle_obstacle((0.1, 0.1), (0.5, 0.2), (0.2, 0.45))
    sim.update(1.0 / 60.0)
    sim.add_velocity((0.5, 0.34 + 0.15 * k), (0.8, -0.6), 7.0)


def _full_step_worker(rank, world, port, errors):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        v0 = np.random.default_rng(21).uniform(-0.9, 0.9, (FH, FW, 2)).astype(np.float32)
        ref = O.OracleFluidSimulator(FW, FH, None)
        ref.speed, ref.vorticity, ref.viscosity, ref.iterations = FSPEED, 1.5, 0.4, FITER
        ref.velocity = v0
        rig = NumpyFullStepRig(rank, world, v0)
        slab = SlabSimulator(FW, FH, engine=rig, halo=FHALO, depth=FDEPTH)
        slab.iterations = FITER
        for k in range(FSTEPS):
            _full_step_script(ref, k)
            _full_step_script(slab, k)
            for name in ("velocity", "pressure", "divergence"):
                a, b = rig.own(getattr(rig.sim, name)), rig.own(getattr(ref, name))
                assert a.tobytes() == b.tobytes(), f"step {k}: {name} differs in {int((a != b).sum())} values"
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001 - reported to the parent
        import traceback
        errors.put(f"rank {rank}: {e}\n{traceback.format_exc()}")


@pytest.mark.parametrize("world", [2, 3])
def test_whole_step_over_slabs_matches_single_process_oracle_over_gloo(world):
    """advect -> vorticity -> confinement -> viscosity -> divergence -> Jacobi -> gradient with the exchange
    schedule of SlabSimulator.update, three frames with obstacles and impulses across the slab boundaries; a
    velocity halo of 2 rows instead of ceil(1.25 dt speed) + 5 makes it fail."""
    _run_world(_full_step_worker, world)
