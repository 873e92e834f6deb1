set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/n2_a.json 2> gpurun_out/n2.err; echo rc=$?; tail -2 gpurun_out/n2.err
GLSDET_BENCH_NO_GATHER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/n2_b.json 2> gpurun_out/n2.err; echo rc=$?
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/n1_a.json 2>/dev/null
python - <<'PY'
import json
for f in ("n1_a","n2_a","n2_b"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(f,"ERR",e)
PY
