O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/an_step_launches_p0.csv python tools/profile_step.py p0 > $O/an_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_table.py $O/an_step_launches_p0.csv $O/step_ops_p0.json $O/an_step_launches_p0.txt $O/an_conv_traffic.json
sed -n 30,34p $O/an_step_launches_p0.txt | cut -c1-120
sed -n 56,75p $O/an_step_launches_p0.txt | cut -c1-120
