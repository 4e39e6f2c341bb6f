"""A model of the overlapped Jacobi group schedule (natrix_b200/csrc/api.cu: phase_jacobi_interior /
phase_jacobi_edges, driven by natrix_b200/slabs.py) and a hazard check over it.

Every queued operation is listed with its stream, the rows of the two pressure buffers it reads and
writes, the events it waits for and the event recorded behind it.  Two operations that touch the same
rows of the same buffer, one of them writing, must be ordered by stream order or by an event - otherwise
the slab result is no longer bit-identical to the single-GPU run.  The model also shows WHY the schedule
has its two waits (negative cases): the exchange waits for everything queued before the group, and the rest of
the group waits for the first launch's edge zones (and through them for the exchange).
"""
import itertools

import pytest


def group_depths(sweeps, depth):
    full = [depth] * (sweeps // depth)
    part = [sweeps % depth] if sweeps % depth else []
    return part + full


class Schedule:
    def __init__(self):
        self.ops = []            # dicts: stream, name, reads, writes, waits (events), record (event or None)

    def op(self, stream, name, reads=(), writes=(), waits=(), record=None):
        self.ops.append(dict(stream=stream, name=name, reads=[r for r in reads if r[2] > r[1]],
                             writes=[w for w in writes if w[2] > w[1]], waits=set(waits), record=record))

    def happens_before(self):
        n = len(self.ops)
        hb = [[False] * n for _ in range(n)]
        recorded = {}                                   # event -> index of the op it was recorded behind (latest)
        last_in_stream = {}
        pending_waits = {"main": set(), "comm": set()}
        for i, o in enumerate(self.ops):
            s = o["stream"]
            preds = set()
            if s in last_in_stream:
                preds.add(last_in_stream[s])
            for e in o["waits"]:
                assert e in recorded, f"{o['name']} waits for {e}, which nobody has recorded yet (a no-op in CUDA)"
                preds.add(recorded[e])
            for p in preds:
                hb[p][i] = True
                for q in range(n):
                    if hb[q][p]:
                        hb[q][i] = True
            last_in_stream[s] = i
            if o["record"]:
                recorded[o["record"]] = i
        return hb

    def hazards(self):
        hb = self.happens_before()
        out = []
        for i, j in itertools.combinations(range(len(self.ops)), 2):
            if hb[i][j] or hb[j][i]:
                continue
            a, b = self.ops[i], self.ops[j]
            for (x, y) in ((a["writes"], b["writes"] + b["reads"]), (a["reads"], b["writes"])):
                for (bu, lo, hi) in x:
                    for (bv, l2, h2) in y:
                        if bu == bv and lo < h2 and l2 < hi:
                            out.append((a["name"], b["name"], bu, max(lo, l2), min(hi, h2)))
        return out


def build(groups, depth, hl, up=True, down=True, wait_edges=True, wait_group=True):
    """The operations of consecutive groups of `groups` sweeps on a slab of hl rows (api.cu: phase_jacobi_interior,
    phase_jacobi_edges; the exchange between them is queued by step_slab or by slabs.py)."""
    s = Schedule()
    src = 0
    s.op("main", "divergence / previous step", writes=[(0, -10**6, 10**6), (1, -10**6, 10**6)])
    for gi, t in enumerate(groups):
        depths = group_depths(t, depth)
        g = f"g{gi}"
        d1 = depths[0]
        # phase 4: everything queued so far precedes the exchange and the edge zones
        s.op("main", f"{g} mark", record=f"{g}.group")
        s.op("main", f"{g} I1", reads=[(src, 0 if up else -d1, hl if down else hl + d1)],
             writes=[(1 - src, d1 if up else 0, hl - d1 if down else hl)])
        s.op("comm", f"{g} exchange", waits=[f"{g}.group"] if wait_group else [],
             reads=([(src, 0, t)] if up else []) + ([(src, hl - t, hl)] if down else []),
             writes=([(src, -t, 0)] if up else []) + ([(src, hl, hl + t)] if down else []))
        # phase 5: the first launch's edge zones behind the exchange, the rest of the group in one piece on main
        rem = t - d1
        if up:
            s.op("comm", f"{g} B1 top", reads=[(src, -rem - d1, d1 + d1)], writes=[(1 - src, -rem, d1)])
        if down:
            s.op("comm", f"{g} B1 bottom", reads=[(src, hl - d1 - d1, hl + rem + d1)], writes=[(1 - src, hl - d1, hl + rem)])
        s.op("comm", f"{g} edges done", record=f"{g}.edges")
        s.op("main", f"{g} join", waits=[f"{g}.edges"] if wait_edges else [])
        cur, done = 1 - src, d1
        for j, d in enumerate(depths[1:], start=2):
            done += d
            rem = t - done
            lo, hi = (-rem if up else 0), (hl + rem if down else hl)
            s.op("main", f"{g} F{j}", reads=[(cur, lo - d, hi + d)], writes=[(1 - cur, lo, hi)])
            cur = 1 - cur
        src = cur
    s.op("main", "gradient", reads=[(src, -1, hl + 1)])
    return s


@pytest.mark.parametrize("groups,depth,hl", [([8, 48, 48], 8, 4096), ([4, 24, 24, 24, 24], 8, 2048), ([13, 24], 8, 512),
                                             ([5, 8], 4, 64), ([3, 7, 7], 1, 40), ([48], 8, 96)])
@pytest.mark.parametrize("up,down", [(True, True), (True, False), (False, True)])
def test_overlapped_group_schedule_has_no_hazard(groups, depth, hl, up, down):
    assert build(groups, depth, hl, up, down).hazards() == []


def test_the_first_interior_launch_really_runs_beside_the_exchange():
    s = build([48], 8, 4096)
    hb = s.happens_before()
    names = [o["name"] for o in s.ops]
    i1, x, b1 = names.index("g0 I1"), names.index("g0 exchange"), names.index("g0 B1 top")
    assert not hb[i1][x] and not hb[x][i1]                       # the interior of the first launch overlaps the exchange
    assert not hb[i1][b1] and not hb[b1][i1]                     # ... and the edge zones
    assert hb[x][names.index("g0 F2")] and hb[b1][names.index("g0 F2")]


def test_the_rules_of_the_schedule_are_all_needed():
    # the rest of the group writes the buffer the exchange sends from and the edge zones read: without the wait
    # for the edges it races with both
    bad = build([24, 24], 8, 2048, wait_edges=False).hazards()
    assert any({a, b} == {"g0 exchange", "g0 F2"} for a, b, *_ in bad) and any("B1" in a + b and "F2" in a + b for a, b, *_ in bad)
    # the exchange must wait for everything queued before the group (it sends rows the previous group wrote)
    bad = build([24, 24], 8, 2048, wait_group=False).hazards()
    assert any("g1 exchange" in (a, b) and ("g0 F3" in (a, b) or "g0 F2" in (a, b)) for a, b, *_ in bad)
