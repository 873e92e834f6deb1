set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -2 gpurun_out/bench_n2.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox_as_n2.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm_n2.json 2>/dev/null; echo ref rc=$?
python - <<'PY'
import json
for f in ("r2_bench_n1_samebox_as_n2","r2_bench_n2","r2_bench_reference_arm_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["scaling"])
    except Exception as e: print(f,"ERR",e)
PY
