set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -12 $O/gpu_tests.log | cut -c1-220
timeout 300 python tools/dw_bench.py > $O/r2_dw_bench.txt 2>&1; tail -3 $O/r2_dw_bench.txt | cut -c1-160
