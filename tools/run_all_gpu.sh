set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
timeout 300 python bench.py --variant p1 --no-cpu-baseline > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err; echo benchp1 rc=$?; tail -3 gpurun_out/bench_p1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json","gpurun_out/bench_p1.json"):
    try:
        d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["segment_ms"], d["roofline"]["frac"], d["roofline"]["algorithmic_gflop_per_image"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
