O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_nano_gpu.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-250
echo "--- tiled"; GLSDET_DW_STREAM=0 timeout 200 python tools/dw_bench.py 2>&1 | tee $O/r2_dw_bench_tiled.txt | cut -c1-200
echo "--- stream"; timeout 200 python tools/dw_bench.py 2>&1 | tee $O/r2_dw_bench_stream.txt | cut -c1-200
