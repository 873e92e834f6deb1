# nano re-test with trained-like conditioning, depthwise bandwidth table, overlapped feature loads A/B
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_nano_gpu.py -m gpu -q > $O/nano_tests.log 2>&1; echo "nano rc=$?"; tail -25 $O/nano_tests.log | cut -c1-220
timeout 300 python tools/dw_bench.py > $O/r2_dw_bench.txt 2>&1; cat $O/r2_dw_bench.txt
for e in 0 1 0 1; do
if [ $e = 1 ]; then export GLSDET_OVERLAP_LOAD=1; else unset GLSDET_OVERLAP_LOAD; fi
timeout 300 python bench.py --no-cpu-baseline > $O/ab_$e.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$e.json"))
print("overlap=$e", round(d["value"],1), round(d["ms_per_step"],3), "seg", round(d["roofline"]["segment_ms"],3), "post", round(d["config"]["postprocess_ms"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
unset GLSDET_OVERLAP_LOAD
GLSDET_OVERLAP_LOAD=1 timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -x 2>&1 | tail -3
