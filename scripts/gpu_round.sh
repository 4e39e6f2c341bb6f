#!/usr/bin/env bash
# One GPU session: smoke, parity tests, bench (both arms), ncu launch list + full capture of the
# Jacobi kernel.  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/.
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -5 "$OUT/smoke.log"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -25 "$OUT/pytest_gpu.log"
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
echo "== bench pipeline 0"; timeout 600 python bench.py --steps 5 --warmup 3 --pipeline 0 --no-cpu > "$OUT/bench_p0.json" 2> "$OUT/bench_p0.err"; cat "$OUT/bench_p0.json"
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; cat "$OUT/bench_ref.json"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu > "$OUT/ncu_list.log" 2>&1; echo "ncu list rc=$?"
echo "== ncu full (jacobi)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tb -s 20 -c 2 -o "$OUT/jacobi_full" -f python bench.py --steps 2 --warmup 3 --no-cpu > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"
ls -la "$OUT"
