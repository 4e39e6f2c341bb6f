#!/usr/bin/env bash
# N-GPU session (gpurun --gpus N): multi-GPU parity tests, the torchrun bench line of both arms, and
# compute-sanitizer (memcheck / racecheck / synccheck) over the pure-C multi-GPU host and the smoke run.
# usage: bash scripts/gpu_multi.sh <tag> <N> [sanitize]
set -u
TAG=${1:-r2multi}; N=${2:-2}; SAN=${3:-}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > "$OUT/pytest_multi.log" 2>&1; echo "pytest multi rc=$?"; tail -3 "$OUT/pytest_multi.log"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench n=$N rc=$?"; tail -c 1500 "$OUT/bench_n$N.json"
timeout 600 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref_n$N.json" 2> "$OUT/bench_ref_n$N.err"; echo "ref n=$N rc=$?"; cut -c1-400 "$OUT/bench_ref_n$N.json"
timeout 600 $TR scripts/slab_check.py 2048 $((512 * N)) 3 37 > "$OUT/slab_check_n$N.log" 2>&1; echo "slab_check rc=$?"; grep SLAB_CHECK "$OUT/slab_check_n$N.log" | cut -c1-200
if [ -n "$SAN" ]; then
  NCCL=$(python -c "import torch,os;print(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)),'nvidia','nccl','lib','libnccl.so.2'))")
  for TOOL in memcheck synccheck; do
    NATRIX_NCCL_LIB=$NCCL timeout 900 compute-sanitizer --tool $TOOL --report-api-errors no --error-exitcode 9 examples/_build/c_host_multi 2 3 > "$OUT/sanitizer_${TOOL}_c_host_multi.log" 2>&1
    echo "$TOOL c_host_multi rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|PASS|FAIL" "$OUT/sanitizer_${TOOL}_c_host_multi.log" | head -5
  done
  for TOOL in racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitizer_${TOOL}_smoke.log" 2>&1
    echo "$TOOL smoke rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/sanitizer_${TOOL}_smoke.log" | head -3
  done
fi
du -sh "$OUT"
