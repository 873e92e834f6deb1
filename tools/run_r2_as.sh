O=gpurun_out
for i in 1 2; do
timeout 300 python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline > $O/as_cfg3_$i.json 2>/dev/null; echo rc=$?
done
GLSDET_MPDET_STREAMS=0 timeout 300 python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline > $O/as_cfg3_nostreams.json 2>/dev/null; echo rc=$?
python - <<'PY'
import json
for f in ("as_cfg3_1","as_cfg3_2","as_cfg3_nostreams"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3))
PY
