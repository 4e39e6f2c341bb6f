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
