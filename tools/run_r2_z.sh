timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -3
for dbg in 0 4 0 4; do
GLSDET_CONV_DBG=$dbg timeout 300 python bench.py --no-cpu-baseline > gpurun_out/z_$dbg.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/z_$dbg.json"))
print("dbg=$dbg", round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
