# nano (DWConv) bring-up + full GPU suite + refreshed headline bench / launch list on one B200
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_nano_gpu.py -m gpu -q -x > $O/nano_tests.log 2>&1; echo "nano rc=$?"; tail -25 $O/nano_tests.log | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_nano_gpu.py > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -6 $O/gpu_tests.log | cut -c1-220
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/r2_bench_n1.json 2> $O/bench_n1.err; echo rc=$?
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_cfg3_n1.json 2>/dev/null; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo ncu bench rc=$?
python tools/summarize_launches.py $O/r2_launches_bench.csv 3 $O/r2_launches_bench.txt > /dev/null 2>&1 || echo "summarize failed"
python - <<'PY'
import json
for f in ("r2_bench_n1","r2_bench_cfg3_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(r.get("frac") or 0,3), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
