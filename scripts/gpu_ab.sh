#!/usr/bin/env bash
# A/B timing of library variants built by scripts/ab_build.sh + the f32x2 microbenchmark.  usage: gpu_ab.sh <tag> <variant...>
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
if [ -f scripts/microbench/f32x2_rates.cu ] && [ ! -f "$OUT/f32x2_rates.txt" ]; then
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/f32x2_rates scripts/microbench/f32x2_rates.cu 2>/dev/null && /tmp/f32x2_rates > "$OUT/f32x2_rates.txt" 2>&1
  cat "$OUT/f32x2_rates.txt"
fi
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -3 "$OUT/pytest_gpu.log"
# a variant is the name of a library built by ab_build.sh, or @NAME=VALUE: the main library with that environment variable
STEPS=${STEPS:-20}; WLS=${WLS:-"cfg5 cfg3 demo"}
for REP in 1 2; do
for V in main "$@"; do
  LIB=""; EV="NATRIX_AB_NONE=1"; TAGV=$V
  case "$V" in
    main) ;;
    @*) EV="${V#@}"; TAGV=$(echo "${V#@}" | tr '=' '_');;
    *) LIB="$PWD/natrix_b200/_ab/lib_$V.so";;
  esac
  for WL in $WLS; do
    env NATRIX_B200_LIB=$LIB "$EV" timeout 300 python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu > "$OUT/bench_${TAGV}_${WL}_$REP.json" 2>> "$OUT/bench.err"
  done
done
done
python - "$OUT" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try: d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stage_ms"].items() if v > 0.004}, "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
