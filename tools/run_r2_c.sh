set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/gpu_tests.log | cut -c1-300
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"mask_tiles|sort_keys|resolve_kernel" -o gpurun_out/prof_r2_nms -f python tools/profile_step.py p0 > gpurun_out/ncu_nms.log 2>&1; echo ncu rc=$?
