#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dm_dropin.py -x -q -m gpu > gpurun_out/t_dm.log 2>&1; echo "dm tests rc=$?"; tail -5 gpurun_out/t_dm.log
for cfg in "5000 100" "2000 0"; do
  set -- $cfg
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/dm_chain4_$1_$2.csv python scripts/dm_chain.py $1 $2 2 > gpurun_out/dm_chain4_$1_$2.log 2>&1
  echo "chain $cfg rc=$?"; tail -1 gpurun_out/dm_chain4_$1_$2.log
  python scripts/launch_table.py gpurun_out/dm_chain4_$1_$2.csv 2 | grep -i "ferguson\|colsum\|total"
done
