# half-bias SiLU epilogue + 4 accumulator slots for the two-group TMA-store class: tests, op times (A/B knobs), bench
set -x
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_path_gpu.py tests/test_backbone_gpu.py tests/test_p1_gpu.py -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log | cut -c1-300
timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_times_new.txt 2>&1; tail -1 gpurun_out/op_times_new.txt
GLSDET_CONV_NO_NACC4=1 timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_times_nacc2.txt 2>&1; tail -1 gpurun_out/op_times_nacc2.txt
GLSDET_CONV_BRES_MT1=1 timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_times_mt1.txt 2>&1; tail -1 gpurun_out/op_times_mt1.txt
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d.get("roofline",{})
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"))
PY
paste gpurun_out/op_times_new.txt gpurun_out/op_times_nacc2.txt gpurun_out/op_times_mt1.txt | awk -F'\t' '{printf "%s | %s | %s\n", substr($1,1,70), substr($2,7,12), substr($3,7,12)}'
