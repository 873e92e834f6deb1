timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -3
C=csp1x1_c128_n128,csp1x1_c128_n128_post1,ffa1x1_c256_n128_post0,c3p3_c128_n128_pre1
echo "== res tma"; python tools/conv_bench.py --cases $C
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x 2>&1 | tail -3
for e in 0 1; do
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/z_$e.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/z_$e.json"))
print("run=$e", round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
