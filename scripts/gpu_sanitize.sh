#!/usr/bin/env bash
# compute-sanitizer memcheck over the smoke run and the small-grid parity tests (both pipelines, ragged
# sizes, impulses, obstacles, dye, frame rendering).  Output: gpurun_out/<tag>/sanitizer_*.log
set -u
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitizer_smoke.log" 2>&1
echo "sanitizer smoke rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds" "$OUT/sanitizer_smoke.log" | head -5
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "golden or ragged or impulse or render or rgba or depth" > "$OUT/sanitizer_tests.log" 2>&1
echo "sanitizer tests rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" "$OUT/sanitizer_tests.log" | head -8
