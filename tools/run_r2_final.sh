# Round-2 final evidence on one B200: tests, smoke, parity report, every bench line, ncu launch lists, top-kernel capture.
set -x
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $O/gpu_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python tools/parity_report.py > $O/r2_parity_report.txt 2>&1
timeout 600 python bench.py > $O/r2_bench_n1.json 2> $O/bench_n1.err; echo rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/bench_ref.err; echo rc=$?
timeout 600 python bench.py --max-det 1000 --no-cpu-baseline > $O/r2_bench_n1_top1000.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --cpu-images 2 > $O/r2_bench_cfg4_n1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/r2_bench_cfg3_n1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_cfg5_n1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --variant p1 --no-cpu-baseline > $O/r2_bench_p1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --variant p2 --no-cpu-baseline > $O/r2_bench_p2.json 2>/dev/null; echo rc=$?
timeout 300 python tools/step_op_times.py p0 > $O/r2_op_times_p0.txt 2>&1
timeout 300 python tools/backbone_op_times.py > $O/r2_backbone_op_times.txt 2>&1
timeout 300 python tools/conv_trace.py > $O/r2_conv_trace_p0.txt 2>&1
M="gpu__time_duration.sum,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/r2_step_launches_p0.csv python tools/profile_step.py p0 > $O/ncu_step.log 2>&1; echo ncu rc=$?
python tools/launch_table.py $O/r2_step_launches_p0.csv $O/step_ops_p0.json $O/r2_step_launches_p0.txt $O/r2_conv_traffic.json > /dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo ncu bench rc=$?
python tools/summarize_launches.py $O/r2_launches_bench.csv 3 $O/r2_launches_bench.txt > /dev/null 2>&1 || echo "summarize failed"
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 39 --launch-count 1 -o $O/prof_r2_top -f python tools/profile_step.py p0 > $O/ncu_top.log 2>&1; echo ncu top rc=$?
ncu -i $O/prof_r2_top.ncu-rep --page details > $O/r2_top_kernel_ncu_full.txt 2>/dev/null
python - <<'PY'
import json
for f in ("r2_bench_n1","r2_bench_n1_top1000","r2_bench_cfg4_n1","r2_bench_cfg3_n1","r2_bench_cfg5_n1","r2_bench_p1","r2_bench_p2","r2_bench_reference_arm"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(r.get("frac") or 0,3), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
