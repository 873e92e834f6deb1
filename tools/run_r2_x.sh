set -x
timeout 900 python -m pytest tests/test_mpdet_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cfg3_all.json 2> gpurun_out/cfg3_all.err; echo rc=$?; tail -2 gpurun_out/cfg3_all.err
python -c "
import json; d=json.load(open('gpurun_out/cfg3_all.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
GLSDET_MPDET_STREAMS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/mpdet_launches.csv python tools/profile_mpdet.py > gpurun_out/mpdet_ncu.log 2>&1; echo rc=$?
python tools/launch_table.py gpurun_out/mpdet_launches.csv > gpurun_out/mpdet_launch_table.txt; grep -A16 "^total" gpurun_out/mpdet_launch_table.txt
grep gfl_select gpurun_out/mpdet_launch_table.txt | head -15
