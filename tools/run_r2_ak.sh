O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_nano_gpu.py -m gpu -q 2>&1 | tail -40 | cut -c1-250
