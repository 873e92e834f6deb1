O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_nano_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/dw_bench.py 2>&1 | cut -c1-150
