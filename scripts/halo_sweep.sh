#!/bin/bash
# A/B of the slab halo depth and of the exchange overlap: scripts/halo_sweep.sh <n_gpus> "<halo>:<overlap>:<workload> ..."
N=$1; shift
for spec in $1; do
  IFS=: read -r halo ov wl <<< "$spec"
  NATRIX_SLAB_HALO=$halo NATRIX_SLAB_OVERLAP=$ov timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps 6 --warmup 3 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('halo $halo overlap $ov', d['config']['workload'], d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['halo']['exchanges_per_step'],2), round(d['roofline']['jacobi_ms_per_step'],3), round(d['weak_base']['ms_per_step'],3))"
done
