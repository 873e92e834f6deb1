set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log | cut -c1-300
timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_a.txt 2>&1; tail -1 gpurun_out/op_a.txt
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d.get("roofline",{})
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "wb", d["with_backbone"]["value"])
PY
