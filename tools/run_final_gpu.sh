# Final evidence run (one GPU): tests, bench lines, reference arm, launch list, per-launch conv table, top-kernel capture.
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log | cut -c1-200
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --variant p1 --no-cpu-baseline > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err
timeout 300 python bench.py --variant p2 --no-cpu-baseline > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err
timeout 300 python tools/conv_bench.py > gpurun_out/conv_bench_final.log 2>&1
timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_times_p0.txt 2>&1
timeout 300 python tools/backbone_op_times.py > gpurun_out/backbone_op_times.txt 2>&1
timeout 300 python tools/mpdet_bench.py 8 > gpurun_out/mpdet_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo ncu1 rc=$?
M="gpu__time_duration.sum,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum"
timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/step_launches_p0.csv python tools/profile_step.py p0 > gpurun_out/ncu_step_p0.log 2>&1; echo ncu2 rc=$?
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 3 --launch-count 1 -o gpurun_out/prof_final_head3x3_n256 -f python tools/conv_bench.py --cases head3x3_s4_n256 --iters 1 > gpurun_out/ncu_full.log 2>&1; echo ncu3 rc=$?
python - <<'PY'
import json
for f in ("bench_n1","bench_ref","bench_p1","bench_p2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), r.get("frac"), r.get("segment_ms"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
