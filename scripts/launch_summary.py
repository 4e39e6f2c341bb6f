"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/launch_summary.py gpurun_out/<tag>/launches.csv [steps]
(steps defaults to the number of pre-projection launches in the list: one per step)"""
import collections
import csv
import io
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
text = "".join(line for line in open(path) if line.startswith('"'))
rows = list(csv.DictReader(io.StringIO(text)))
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"])
if steps <= 0:
    # one pre-projection (or, under the reference-order pipeline, one advect) launch per step
    steps = max([c for name, (c, _) in agg.items() if name.startswith(("k_preproject", "k_advect"))] + [1])
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':45s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:45s} {c:8d} {t / 1e3:10.1f} {t / c / 1e3:9.1f} {100 * t / tot:5.1f}%")
print(f"all kernels: {tot / 1e3:.1f} us over {len(rows)} launches; {tot / 1e3 / steps:.1f} us per step ({steps} steps)")
