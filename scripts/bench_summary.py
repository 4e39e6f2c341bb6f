"""One line per bench JSON in a directory: value, ms/step, stages, e2e, physical roofline fraction.
usage: python scripts/bench_summary.py gpurun_out/<tag>"""
import glob
import json
import sys

for f in sorted(glob.glob(sys.argv[1] + "/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    if "stage_ms" not in d:
        continue
    print(f.split("/")[-1], d["config"]["workload"], round(d["value"]), "ms", round(d["ms_per_step"], 4),
          {k: round(v, 4) for k, v in d["stage_ms"].items() if v > 0.004}, "e2e", round(d["e2e"]["ms_per_step"], 4),
          "median", round(d["e2e"].get("ms_per_step_median", 0), 4), "frac", round(d["roofline"]["frac"], 3))
    for k in ("config3_4096", "config5_moving", "config4_16384"):
        if k in d:
            s = d[k]
            print("   ", k, round(s["value"]), round(s["ms_per_step"], 3), {a: round(v, 4) for a, v in s["stage_ms"].items() if v > 0.004},
                  "e2e", round(s["e2e"]["ms_per_step"], 3), "frac", s.get("roofline", {}).get("frac"))
