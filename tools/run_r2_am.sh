O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
timeout 300 python bench.py --no-cpu-baseline > $O/am_bench.json 2> $O/am_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/am_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'seg',d['roofline']['segment_ms'], {k:v for k,v in d.items() if 'post' in k})
PY
timeout 300 python tools/step_op_times.py p0 2>&1 | tail -45 | cut -c1-160
