set -x
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_path_gpu.py tests/test_backbone_gpu.py tests/test_p1_gpu.py -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log | cut -c1-300
timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_a.txt 2>&1
GLSDET_CONV_SPLIT_N=0 timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_b.txt 2>&1
paste gpurun_out/op_a.txt gpurun_out/op_b.txt | awk -F'\t' '{printf "%s | %s\n", substr($1,1,70), substr($2,7,12)}'
