#!/usr/bin/env bash
# DRAM traffic and duration of k_jacobi_tb launches under a few settings (ncu, metrics only: fast)
set -u
OUT=${1:-gpurun_out/traffic}; mkdir -p "$OUT"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for WL in cfg5 cfg3; do
 for PROMO in 3 2 0; do
  NATRIX_TB_L2PROMO=$PROMO timeout 300 ncu --metrics $M --clock-control none -k regex:k_jacobi_tb -s 30 -c 3 --csv --log-file "$OUT/${WL}_promo$PROMO.csv" python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
  python - "$OUT/${WL}_promo$PROMO.csv" "$WL promo=$PROMO" <<'PY'
import csv, io, sys, collections
text = "".join(l for l in open(sys.argv[1]) if l.startswith('"'))
rows = list(csv.DictReader(io.StringIO(text)))
agg = collections.defaultdict(list)
for r in rows: agg[r["Metric Name"]].append((float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
print(sys.argv[2], {k: (round(sum(x for x, _ in v) / len(v), 2), v[0][1]) for k, v in agg.items()})
PY
 done
done
