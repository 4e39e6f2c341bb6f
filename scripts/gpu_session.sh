set -u
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg5", round(d["value"]), round(d["ms_per_step"], 3), d["stage_ms"], "e2e", round(d["e2e"]["value"]), d["plan_cache"], "cpu", d["cpu_baseline"])
for k in ("config3_4096", "config5_moving", "config4_16384"):
    s = d[k]; print(k, round(s["value"]), round(s["ms_per_step"], 3), s["stage_ms"], "e2e", round(s["e2e"]["value"]), s.get("plan_cache"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-600 $OUT/bench_ref.json
