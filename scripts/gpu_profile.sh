#!/usr/bin/env bash
# ncu evidence for one build: launch lists of a bench run + --set full captures of the top kernels.
# Primary workload = bench.py default (config 5, 32768 x 4096 per GPU); config 3 (4096^2) alongside.
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for WL in cfg5 cfg3; do
  BENCH="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$OUT/launches_$WL.csv" $BENCH > "$OUT/ncu_list_$WL.log" 2>&1; echo "list $WL rc=$?"
  for k in k_jacobi_tb k_preproject k_gradient_mask; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o "$OUT/${WL}_$k" -f $BENCH > "$OUT/ncu_${WL}_$k.log" 2>&1; echo "$WL $k rc=$?"
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dye_advect -s 3 -c 1 -o "$OUT/cfg3_k_dye_advect" -f python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > "$OUT/ncu_cfg3_dye.log" 2>&1
python bench.py --steps 20 > "$OUT/bench.json" 2> "$OUT/bench.err"
python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
python bench.py --workload cfg3 --pipeline 0 --steps 5 --no-cpu > "$OUT/bench_cfg3_pipeline0.json" 2>&1
ls -la "$OUT" | head -40
