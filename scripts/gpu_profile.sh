#!/usr/bin/env bash
# ncu evidence for one build: launch list of a bench run + --set full captures of the top kernels.
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
BENCH="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" $BENCH > "$OUT/ncu_list.log" 2>&1; echo "list rc=$?"
for k in k_jacobi_tb k_preproject k_dye_advect k_gradient_mask; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o "$OUT/$k" -f $BENCH > "$OUT/ncu_$k.log" 2>&1; echo "$k rc=$?"
done
ls -la "$OUT"
