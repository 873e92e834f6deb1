O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_boundary_gpu.py tests/test_facade_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-250
for i in 1 2; do
echo "--- plain"; GLSDET_NO_PAIR_CSP=1 timeout 200 python tools/backbone_op_times.py 2>&1 | sed -n 4,8p\;\$p
echo "--- pairs"; timeout 200 python tools/backbone_op_times.py 2>&1 | sed -n 4,8p\;\$p
done
