O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dw3 --launch-skip 2 --launch-count 1 -o $O/prof_dw3 -f python tools/dw_one.py > $O/ncu_dw3.log 2>&1; echo rc=$?
ncu -i $O/prof_dw3.ncu-rep --page details > $O/r2_ncu_dw3.txt 2>/dev/null
grep -E "Duration|Throughput|Pipe|Issue|IPC|Warp Cycles|Registers|Achieved Occupancy|Theoretical Occ|L1/TEX Hit|L2 Hit|dram__bytes|DRAM Throughput|Mem Busy|Max Bandwidth|Stall|Executed Ipc|No Eligible|Eligible" $O/r2_ncu_dw3.txt | head -60
