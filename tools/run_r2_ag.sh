O=gpurun_out
mkdir -p $O
timeout 600 python tools/fp32_time.py > $O/r2_fp32_mode.txt 2>&1; cat $O/r2_fp32_mode.txt | tail -4
