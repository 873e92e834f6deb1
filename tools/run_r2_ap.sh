O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_boundary_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-250
for i in 1 2; do
echo "--- prev"; GLSDET_LIB=glsdet_b200/lib/libglsdet_b200_prev.so timeout 200 python tools/nms_time.py 2>&1 | tail -4 | head -3
echo "--- new"; timeout 200 python tools/nms_time.py 2>&1 | tail -4 | head -3
done
