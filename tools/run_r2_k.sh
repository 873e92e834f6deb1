set -x
timeout 1200 python -m pytest tests/test_p2_gpu.py tests/test_stock_mmdet_gpu.py tests/test_path_gpu.py tests/test_mpdet_gpu.py -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/gpu_tests.log | cut -c1-300
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
M="gpu__time_duration.sum"
timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/step_launches_p0.csv python tools/profile_step.py p0 > gpurun_out/ncu_step_p0.log 2>&1; echo ncu rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d.get("roofline",{})
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"))
PY
python tools/launch_table.py gpurun_out/step_launches_p0.csv gpurun_out/step_ops_p0.json | tail -32
