set -x
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/r2_bench_n$n.json 2> gpurun_out/bench_n$n.err; echo rc=$?; tail -2 gpurun_out/bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --config cfg5 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_n8.json 2> gpurun_out/bench_cfg5_n8.err; echo rc=$?
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2_bench_n1_samebox","r2_bench_n4","r2_bench_n8","r2_bench_cfg5_n8"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "fp32img", round(d["e2e"]["from_fp32_images"]["value"],1), d["scaling"])
    except Exception as e: print(f,"ERR",e)
PY
