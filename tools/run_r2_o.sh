set -x
timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_a.txt 2>&1
GLSDET_CONV_2CTA=1 timeout 300 python tools/step_op_times.py p0 > gpurun_out/op_b.txt 2>&1
paste gpurun_out/op_a.txt gpurun_out/op_b.txt | awk -F'\t' '{printf "%s | %s\n", substr($1,1,70), substr($2,7,12)}'
