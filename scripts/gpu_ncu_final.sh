#!/usr/bin/env bash
# ncu --set full captures of the top kernels of the CURRENT build at the bench sizes (raw CSV + details kept, reports
# dropped: gpurun brings back at most 64 MiB), launch lists and per-tile traces.  usage: bash scripts/gpu_ncu_final.sh <tag>
set -u
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
keep() { ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.raw.csv" 2>/dev/null; ncu -i "$OUT/$1.ncu-rep" --page details > "$OUT/$1.details.txt" 2>/dev/null; rm -f "$OUT/$1.ncu-rep"; }
for WL in cfg5 cfg3; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_$WL.csv" $NB > /dev/null 2>&1
  python scripts/launch_summary.py "$OUT/launches_$WL.csv" > "$OUT/launches_${WL}_summary.txt" 2>&1
  for k in k_jacobi_tb k_preproject k_gradient_mask; do
    timeout 600 ncu --set full --clock-control none -k regex:$k -s 6 -c 1 -o "$OUT/${WL}_$k" -f $NB > /dev/null 2>&1; echo "$WL $k rc=$?"
    keep "${WL}_$k"
  done
  NATRIX_TB_TRACE=$OUT/tb_trace_$WL.csv timeout 200 $NB > /dev/null 2>&1
done
for WL in demo cfg2; do
  NB="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_$WL.csv" $NB > /dev/null 2>&1
  python scripts/launch_summary.py "$OUT/launches_$WL.csv" > "$OUT/launches_${WL}_summary.txt" 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:k_jacobi_smem -s 6 -c 1 -o "$OUT/${WL}_k_jacobi_smem" -f $NB > /dev/null 2>&1; echo "$WL smem rc=$?"
  keep "${WL}_k_jacobi_smem"
done
du -sh "$OUT"
