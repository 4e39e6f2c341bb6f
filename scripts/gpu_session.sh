set -u
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2_$name.json 2> $OUT/bench_n2_$name.err; echo "bench $name rc=$?"
  python - $OUT/bench_n2_$name.json $name <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(l["value"]), round(l["ms_per_step"], 3), "base", round(l["weak_base"]["ms_per_step"], 3), "eff", round(l["value"] / (2 * l["weak_base"]["value"]), 4), "e2e", round(l["e2e"]["ms_per_step"], 3), l["halo"], "parity", l["slab_parity"]["bit_identical"])
PY
}
run default A=1
run xfirst_r2 NATRIX_SLAB_XFIRST=1 NATRIX_SLAB_RESERVE=2
run xfirst_r8 NATRIX_SLAB_XFIRST=1 NATRIX_SLAB_RESERVE=8
run xfirst_r2_cta2 NATRIX_SLAB_XFIRST=1 NATRIX_SLAB_RESERVE=2 NCCL_MAX_CTAS=2
run nooverlap NATRIX_SLAB_OVERLAP=0
run python NATRIX_SLAB_DRIVER=python
