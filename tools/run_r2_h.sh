# full GPU tests; bench: default (cfg2), cfg4, cfg3, cfg5 (strong scaling, short)
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/gpu_tests.log | cut -c1-300
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
timeout 500 python bench.py --config cfg4 --steps 10 --warmup 3 --cpu-images 2 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo bench cfg4 rc=$?; tail -3 gpurun_out/bench_cfg4.err
timeout 500 python bench.py --config cfg3 --steps 10 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench cfg3 rc=$?; tail -3 gpurun_out/bench_cfg3.err
timeout 500 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo bench cfg5 rc=$?; tail -3 gpurun_out/bench_cfg5_n1.err
python - <<'PY'
import json
for f in ("bench_n1","bench_cfg4","bench_cfg3","bench_cfg5_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, d["metric"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(r.get("frac") or 0,3), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["scaling"])
    except Exception as e: print(f,"ERR",e)
PY
