set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_nano_gpu.py -m gpu -q > $O/nano_tests.log 2>&1; echo "nano rc=$?"; tail -15 $O/nano_tests.log | cut -c1-220
timeout 300 python tools/dw_bench.py > $O/r2_dw_bench.txt 2>&1; cat $O/r2_dw_bench.txt
