O=gpurun_out
mkdir -p $O
for e in 0 1 0 1; do
if [ $e = 1 ]; then export GLSDET_CSP_SIDE=1; else unset GLSDET_CSP_SIDE; fi
timeout 300 python bench.py --no-cpu-baseline > $O/aj_$e.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/aj_$e.json"))
print("side=$e", round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
GLSDET_CSP_SIDE=1 timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -x 2>&1 | tail -2
