set -x
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo rc=$?; tail -3 gpurun_out/bench_n$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 3 --warmup 2 --config cfg5 > gpurun_out/bench_cfg5_n$n.json 2> gpurun_out/bench_cfg5_n$n.err; echo rc=$?; tail -3 gpurun_out/bench_cfg5_n$n.err
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err
python - <<'PY'
import json
for f in ("bench_n1b","bench_n2","bench_cfg5_n2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d.get("roofline",{})
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "fp32img", round(d["e2e"]["from_fp32_images"]["value"],1), d["scaling"])
    except Exception as e: print(f,"ERR",e)
PY
