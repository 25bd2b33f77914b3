#!/bin/bash
mkdir -p gpurun_out
for cfg in "5000 100" "2000 0"; do
  set -- $cfg
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/dm_chain_$1_$2.csv python scripts/dm_chain.py $1 $2 2 > gpurun_out/dm_chain_$1_$2.log 2>&1
  echo "chain $cfg rc=$?"; cat gpurun_out/dm_chain_$1_$2.log | tail -3
  timeout 60 python scripts/dm_chain.py $1 $2 3 2>&1 | tail -3
done
