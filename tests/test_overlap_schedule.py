"""A model of the overlapped Jacobi group schedule (natrix_b200/csrc/api.cu: phase_jacobi_interior /
phase_jacobi_edges, driven by natrix_b200/slabs.py) and a hazard check over it.

Every queued operation is listed with its stream, the rows of the two pressure buffers it reads and
writes, the events it waits for and the event recorded behind it.  Two operations that touch the same
rows of the same buffer, one of them writing, must be ordered by stream order or by an event - otherwise
the slab result is no longer bit-identical to the single-GPU run.  The model also shows WHY the schedule
has its three rules (negative cases): the second interior launch waits for the exchange, launch depths
never decrease inside a group, and the next group waits for this group's edges.
"""
import itertools

import pytest


def group_depths(sweeps, depth, partial_first=True):
    full = [depth] * (sweeps // depth)
    part = [sweeps % depth] if sweeps % depth else []
    return part + full if partial_first else full + part


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


def build(groups, depth, hl, up=True, down=True, wait_exchange=True, partial_first=True, wait_edges=True):
    """The operations of consecutive groups of `groups` sweeps on a slab of hl rows (api.cu, slabs.py)."""
    s = Schedule()
    src = 0
    s.op("main", "divergence / previous step", writes=[(0, -10**6, 10**6), (1, -10**6, 10**6)])
    for gi, t in enumerate(groups):
        depths = group_depths(t, depth, partial_first)
        g = f"g{gi}"
        # phase 4: everything queued so far precedes the exchange and the edge zones
        s.op("main", f"{g} mark", record=f"{g}.group")
        s.op("comm", f"{g} exchange", waits=[f"{g}.group"],
             reads=([(src, 0, t)] if up else []) + ([(src, hl - t, hl)] if down else []),
             writes=([(src, -t, 0)] if up else []) + ([(src, hl, hl + t)] if down else []), record=f"{g}.xchg")
        done, cur = 0, src

        def interior(j, cur, done):
            d = depths[j]
            lo, hi = (done if up else 0), (hl - done if down else hl)
            waits = [f"{g}.xchg"] if (j == 1 and wait_exchange) else []
            s.op("main", f"{g} I{j + 1}", waits=waits, reads=[(cur, lo - d, hi + d)], writes=[(1 - cur, lo, hi)],
                 record=f"{g}.int{j}")

        interior(0, cur, depths[0])
        # phase 5
        for j, d in enumerate(depths):
            done += d
            rem = t - done
            waits = [f"{g}.int{j - 1}"] if j > 0 else []
            if up:
                s.op("comm", f"{g} B{j + 1} top", waits=waits, reads=[(cur, -rem - d, done + d)], writes=[(1 - cur, -rem, done)])
            if down:
                s.op("comm", f"{g} B{j + 1} bottom", waits=waits if not up else [],
                     reads=[(cur, hl - done - d, hl + rem + d)], writes=[(1 - cur, hl - done, hl + rem)])
            cur = 1 - cur
            if j + 1 < len(depths):
                interior(j + 1, cur, done + depths[j + 1])
        s.op("comm", f"{g} edges done", record=f"{g}.edges")
        s.op("main", f"{g} join", waits=[f"{g}.edges"] if wait_edges else [])
        src = cur
    s.op("main", "gradient", reads=[(src, -1, hl + 1)])
    return s


@pytest.mark.parametrize("groups,depth,hl", [([8, 48, 48], 8, 4096), ([4, 24, 24, 24, 24], 8, 2048), ([13, 24], 8, 512),
                                             ([5, 8], 4, 64), ([3, 7, 7], 1, 40), ([48], 8, 96)])
@pytest.mark.parametrize("up,down", [(True, True), (True, False), (False, True)])
def test_overlapped_group_schedule_has_no_hazard(groups, depth, hl, up, down):
    assert build(groups, depth, hl, up, down).hazards() == []


def test_interior_launches_really_run_beside_the_exchange():
    s = build([48], 8, 4096)
    hb = s.happens_before()
    names = [o["name"] for o in s.ops]
    i1, x = names.index("g0 I1"), names.index("g0 exchange")
    assert not hb[i1][x] and not hb[x][i1]                       # the first interior launch overlaps the exchange
    i3, b2 = names.index("g0 I3"), names.index("g0 B2 top")
    assert not hb[i3][b2] and not hb[b2][i3]                     # later edge launches overlap interior launches


def test_the_rules_of_the_schedule_are_all_needed():
    # without the wait, I2 overwrites rows of the buffer the exchange is still sending from (the bug that
    # showed up as 1-ulp differences in the neighbour's pressure)
    bad = build([24, 24], 8, 2048, wait_exchange=False).hazards()
    assert any({"g1 exchange", "g1 I2"} == {a, b} for a, b, *_ in bad)
    # a partial launch at the END of a group lets an edge launch read rows the next interior launch writes
    bad = build([20], 8, 2048, partial_first=False).hazards()
    assert any("B2" in a + b and "I3" in a + b for a, b, *_ in bad)
    # the next group must wait for this group's edges
    assert build([24, 24], 8, 2048, wait_edges=False).hazards()
