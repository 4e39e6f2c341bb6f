"""Summarise a NATRIX_TB_TRACE dump (per-tile start / end / SM of one k_jacobi_tb launch).
usage: python scripts/tb_trace_summary.py trace.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.DictReader(open(sys.argv[1])))
t0 = min(int(r["start_ns"]) for r in rows)
t1 = max(int(r["end_ns"]) for r in rows)
dur = sorted(int(r["end_ns"]) - int(r["start_ns"]) for r in rows)
starts = sorted(int(r["start_ns"]) - t0 for r in rows)
ends = sorted(int(r["end_ns"]) - t0 for r in rows)
n = len(rows)
q = lambda a, f: a[min(len(a) - 1, int(f * len(a)))]
print(f"{n} tiles, depth {rows[0]['depth']}, grid width {rows[0]['w']}, rows [{rows[0]['r0']}, {rows[0]['r1']}); span {(t1 - t0) / 1e3:.1f} us")
print(f"tile duration us: min {dur[0] / 1e3:.1f} p10 {q(dur, .1) / 1e3:.1f} p50 {q(dur, .5) / 1e3:.1f} p90 {q(dur, .9) / 1e3:.1f} max {dur[-1] / 1e3:.1f}")
print(f"tile start   us: p0 {starts[0] / 1e3:.1f} p50 {q(starts, .5) / 1e3:.1f} p90 {q(starts, .9) / 1e3:.1f} max {starts[-1] / 1e3:.1f}")
print(f"tile end     us: min {ends[0] / 1e3:.1f} p10 {q(ends, .1) / 1e3:.1f} p50 {q(ends, .5) / 1e3:.1f} p90 {q(ends, .9) / 1e3:.1f} max {ends[-1] / 1e3:.1f}")
sm = defaultdict(list)
for r in rows:
    sm[int(r["smid"])].append((int(r["start_ns"]) - t0, int(r["end_ns"]) - t0, int(r["row1"]) - int(r["row0"])))
busy = sorted((max(e for _, e, _ in v) - min(s for s, _, _ in v)) for v in sm.values())
last = sorted(max(e for _, e, _ in v) for v in sm.values())
print(f"{len(sm)} SMs used; SM busy span us: min {busy[0] / 1e3:.1f} p50 {q(busy, .5) / 1e3:.1f} max {busy[-1] / 1e3:.1f}; "
      f"SM finish us: min {last[0] / 1e3:.1f} p50 {q(last, .5) / 1e3:.1f} max {last[-1] / 1e3:.1f}")
print(f"mean SM busy / span = {sum(busy) / len(busy) / (t1 - t0):.3f}; mean tile busy / span = {sum(dur) / n / (t1 - t0):.3f}")
hs = defaultdict(list)
for r in rows:
    hs[int(r["row1"]) - int(r["row0"])].append(int(r["end_ns"]) - int(r["start_ns"]))
for h in sorted(hs):
    v = hs[h]
    print(f"  tiles of {h:5d} rows: {len(v):5d}, mean {sum(v) / len(v) / 1e3:.1f} us, max {max(v) / 1e3:.1f} us")
