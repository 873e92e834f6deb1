# Round-2 GPU run E: state check after re-entry: all GPU tests, smoke, parity report, uncapped bench, launch list + metrics of a step
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/gpu_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1; cat gpurun_out/parity_report.txt | cut -c1-220
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
timeout 500 python bench.py --no-cpu-baseline --max-det 1000 > gpurun_out/bench_n1_top1000.json 2> gpurun_out/bench_n1_top1000.err; echo bench rc=$?
M="gpu__time_duration.sum,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/step_launches_p0.csv python tools/profile_step.py p0 > gpurun_out/ncu_step_p0.log 2>&1; echo ncu rc=$?
python - <<'PY'
import json
for f in ("bench_n1","bench_n1_top1000"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), d["config"].get("kept_per_image"))
    except Exception as e: print(f,"ERR",e)
PY
