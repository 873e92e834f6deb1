# Round-2 closing run after the pair-form backbone block: tests, smoke, the bench lines that include the backbone (e2e), op times.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/gpu_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/r2_bench_n1.json 2> $O/bench_n1.err; echo rc=$?
timeout 600 python bench.py --max-det 1000 --no-cpu-baseline > $O/r2_bench_n1_top1000.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_cfg5_n1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --variant p1 --no-cpu-baseline > $O/r2_bench_p1.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --variant p2 --no-cpu-baseline > $O/r2_bench_p2.json 2>/dev/null; echo rc=$?
timeout 300 python tools/backbone_op_times.py > $O/r2_backbone_op_times.txt 2>&1
timeout 900 python tools/parity_report.py > $O/r2_parity_report.txt 2>&1
python - <<'PY'
import json
for f in ("r2_bench_n1","r2_bench_n1_top1000","r2_bench_cfg5_n1","r2_bench_p1","r2_bench_p2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(r.get("frac") or 0,3), "seg", r.get("segment_ms"), "post", d["config"].get("postprocess_ms"), "wb", (d.get("with_backbone") or {}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
