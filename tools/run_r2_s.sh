set -x
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_boundary_gpu.py tests/test_stock_mmdet_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches_p0.csv python tools/profile_step.py p0 > gpurun_out/ncu_step_p0.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d.get("roofline",{})
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"))
PY
python tools/launch_table.py gpurun_out/step_launches_p0.csv gpurun_out/step_ops_p0.json | tail -24 | head -12
