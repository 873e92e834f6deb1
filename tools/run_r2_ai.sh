O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_fp32_gpu.py tests/test_p1_gpu.py tests/test_p2_gpu.py -m gpu -q 2>&1 | tail -25 | cut -c1-220
