O=gpurun_out
mkdir -p $O
GLSDET_HCSP_SPLIT=2 timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "golden or oracle or batch" 2>&1 | tail -3 | cut -c1-200
run() {
  GLSDET_HCSP_SPLIT=$1 timeout 300 python bench.py --no-cpu-baseline > $O/aq_bench_$1.json 2> $O/aq_bench_$1.err; 
  python - $1 <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/aq_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('split',sys.argv[1],'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'seg',round(d['roofline']['segment_ms'],4),'frac',round(d['roofline']['frac'],4),'launches',d['launches_per_step'])
PY
}
run 1; run 4; run 2; run 1; run 4; run 8
