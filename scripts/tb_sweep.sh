#!/usr/bin/env bash
# Jacobi kernel variants at 4096^2: launch shape x packed arithmetic (x obstacles on/off)
set -u
OUT=gpurun_out/${1:-sweep}
SHAPES=${2:-"0 1"}
PACKED=${3:-"0 1"}
mkdir -p "$OUT"
python scripts/tb_probe.py 1024 512 8 | tail -2
for shape in $SHAPES; do for packed in $PACKED; do for extra in "" "--no-obstacles"; do
  NATRIX_TB_SHAPE=$shape timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --packed $packed $extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('shape=$shape packed=$packed $extra', 'value=%.0f' % d['value'], 'ms/step=%.3f' % d['ms_per_step'], 'stage_ms=', d['stage_ms'])
    elif 'Error' in line or 'error' in line: print(line.strip())
" | tee -a "$OUT/sweep.txt"
done; done; done
