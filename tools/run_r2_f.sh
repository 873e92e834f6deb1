# ncu --set full of two small-K conv launches of a step: #27 (1x1 64->128 @256^2, TMA-store 16-warp class) and #23 (1x1 512->512 @64^2, 8-warp class)
set -x
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 27 --launch-count 1 -o gpurun_out/prof_r2_smallk_27 -f python tools/profile_step.py p0 > gpurun_out/ncu_27.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 23 --launch-count 1 -o gpurun_out/prof_r2_smallk_23 -f python tools/profile_step.py p0 > gpurun_out/ncu_23.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
