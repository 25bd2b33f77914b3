#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/t_all2.log 2>&1; echo "gpu suite rc=$?"; tail -10 gpurun_out/t_all2.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_ferguson_sorted" -c 1 -o gpurun_out/ferguson_full -f python scripts/dm_chain.py 2000 0 1 > gpurun_out/ncu_ferg.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ferguson_full.ncu-rep --page details > gpurun_out/ferguson_details.txt 2>&1
grep -n "Duration\|Executed Ipc Active\|FP64\|fp64\|Issue Slots Busy\|Achieved Occupancy\|highest-utilized" gpurun_out/ferguson_details.txt | head
