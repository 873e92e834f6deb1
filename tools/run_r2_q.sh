set -x
timeout 600 python -m pytest tests/test_boundary_gpu.py -m gpu -q -x 2>&1 | tail -4
GLSDET_BENCH_NO_GATHER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r2_bench_n2_nogather.json 2> gpurun_out/bench_n2.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox2.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_bench_n1_samebox2","r2_bench_n2_nogather","r2_bench_n2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(f,"ERR",e)
PY
