set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/gpu_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1; cut -c1-200 gpurun_out/parity_report.txt | tail -30
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d.get("roofline",{})
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "wb", d["with_backbone"]["value"])
PY
