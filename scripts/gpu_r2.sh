#!/usr/bin/env bash
# Round-2 GPU session: parity tests, bench lines (default + small grids with either Jacobi kernel), ncu launch
# lists and --set full captures of the top kernels.  usage: bash scripts/gpu_r2.sh <tag> [quick]
set -u
TAG=${1:-r2}; MODE=${2:-full}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=10 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest_gpu.log"
timeout 600 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; tail -3 "$OUT/bench.err"
B="python bench.py --no-cpu --steps 100 --warmup 5"
for WL in demo cfg2; do
  timeout 200 $B --workload $WL > "$OUT/bench_${WL}.json" 2>> "$OUT/bench.err"
  timeout 200 $B --workload $WL --jacobi-kernel 1 > "$OUT/bench_${WL}_tb.json" 2>> "$OUT/bench.err"
done
timeout 200 python bench.py --no-cpu --steps 20 --warmup 3 --workload cfg3 --pipeline 0 > "$OUT/bench_cfg3_pipeline0.json" 2>> "$OUT/bench.err"
python - "$OUT" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if "stage_ms" not in d: continue
    print(f.split("/")[-1], d["config"]["workload"], d["config"].get("jacobi_kernel"), d["config"].get("jacobi_depth"), "value", round(d["value"]), "ms", round(d["ms_per_step"], 4),
          {k: round(v, 4) for k, v in d["stage_ms"].items() if v > 0.004}, "e2e_ms", round(d["e2e"]["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3))
    for k in ("config3_4096", "config5_moving", "config4_16384"):
        if k in d:
            s = d[k]; print("   ", k, round(s["value"]), round(s["ms_per_step"], 3), s["stage_ms"], "e2e", round(s["e2e"]["value"]))
PY
[ "$MODE" = quick ] && exit 0
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref rc=$?"
# ---- ncu: launch lists, then --set full captures of the top kernels at the bench sizes.  gpurun brings back at most
# 64 MiB: every report is turned into its raw-metric CSV (+ the source page of the hot kernels) on the box and removed.
keep() {   # keep <name> [source]
  ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.raw.csv" 2>/dev/null
  ncu -i "$OUT/$1.ncu-rep" --page details > "$OUT/$1.details.txt" 2>/dev/null
  [ "${2:-}" = source ] && ncu -i "$OUT/$1.ncu-rep" --page source --csv > "$OUT/$1.source.csv" 2>/dev/null
  rm -f "$OUT/$1.ncu-rep"
}
for WL in cfg5 cfg3 demo cfg2; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_$WL.csv" $NB > "$OUT/ncu_list_$WL.log" 2>&1; echo "list $WL rc=$?"
  python scripts/launch_summary.py "$OUT/launches_$WL.csv" > "$OUT/launches_${WL}_summary.txt" 2>&1
done
for WL in cfg5 cfg3; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  for k in k_jacobi_tb k_preproject k_gradient_mask; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o "$OUT/${WL}_$k" -f $NB > "$OUT/ncu_${WL}_$k.log" 2>&1; echo "$WL $k rc=$?"
    keep "${WL}_$k" $([ $WL = cfg3 ] && echo source)
  done
done
for WL in demo cfg2; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_smem -s 6 -c 1 -o "$OUT/${WL}_k_jacobi_smem" -f $NB > "$OUT/ncu_${WL}_smem.log" 2>&1; echo "$WL smem rc=$?"
  keep "${WL}_k_jacobi_smem" source
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dye_advect -s 3 -c 1 -o "$OUT/cfg3_k_dye_advect" -f python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > "$OUT/ncu_cfg3_dye.log" 2>&1
keep cfg3_k_dye_advect
# ---- per-tile trace of one Jacobi launch (idle / imbalance analysis)
NATRIX_TB_TRACE=$OUT/tb_trace_cfg3.csv timeout 200 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
NATRIX_TB_TRACE=$OUT/tb_trace_cfg5.csv timeout 200 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
du -sh "$OUT"; ls "$OUT" | wc -l
