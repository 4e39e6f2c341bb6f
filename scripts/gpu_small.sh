#!/usr/bin/env bash
# Small-grid session: parity tests + the demo / config 2 bench lines with the shared-memory Jacobi kernel in its
# launch shapes.  usage: bash scripts/gpu_small.sh <tag>
set -u
TAG=${1:-small}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest_gpu.log"
B="python bench.py --no-cpu --steps 100 --warmup 5"
for WL in demo cfg2; do
  timeout 200 $B --workload $WL > "$OUT/bench_${WL}_auto.json" 2>> "$OUT/bench.err"
  for D in 8 16; do
    timeout 200 $B --workload $WL --smem-depth $D > "$OUT/bench_${WL}_depth$D.json" 2>> "$OUT/bench.err"
  done
done
python - "$OUT" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try: d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("/")[-1], d["config"].get("jacobi_kernel"), "ms", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stage_ms"].items() if v > 0.004}, "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "launches/step", d["gpu_launches"] / d["steps"])
PY
for WL in demo cfg2; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_$WL.csv" $NB > "$OUT/ncu_list_$WL.log" 2>&1; echo "list $WL rc=$?"
  python scripts/launch_summary.py "$OUT/launches_$WL.csv" 5 | head -24
done
