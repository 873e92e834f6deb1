O=gpurun_out
mkdir -p $O
for i in 1 2; do
echo "--- prev"; GLSDET_LIB=glsdet_b200/lib/libglsdet_b200_prev.so timeout 200 python tools/nms_time.py 2>&1 | tail -4
echo "--- new"; timeout 200 python tools/nms_time.py 2>&1 | tail -4
done
