set -x
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 36 --launch-count 1 -o gpurun_out/prof_r2_top -f python tools/profile_step.py p0 > gpurun_out/ncu_top.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 27 --launch-count 1 -o gpurun_out/prof_r2_smallk -f python tools/profile_step.py p0 > gpurun_out/ncu_smallk.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:mask_tiles --launch-count 1 -o gpurun_out/prof_r2_mask -f python tools/profile_step.py p0 > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:resolve_kernel --launch-count 1 -o gpurun_out/prof_r2_resolve -f python tools/profile_step.py p0 > gpurun_out/ncu_resolve.log 2>&1; echo rc=$?
