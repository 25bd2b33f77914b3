#!/bin/bash
# Round-2 GPU check: GPU suite, smoke, rotation check, bench line, ncu launch list of one C4 PD.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?" | tee gpurun_out/summary.txt
tail -14 gpurun_out/t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 120 python scripts/rotate_check.py > gpurun_out/rotate_check.log 2>&1; tail -4 gpurun_out/rotate_check.log
timeout 120 python scripts/one_pd.py 2000 256 3 > gpurun_out/one_pd.log 2>&1; tail -1 gpurun_out/one_pd.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_one_pd.csv python scripts/one_pd.py 2000 256 2 > gpurun_out/ncu_one_pd.log 2>&1
python scripts/launch_table.py gpurun_out/launches_one_pd.csv 2 > gpurun_out/launch_table_one_pd.txt 2>&1; cat gpurun_out/launch_table_one_pd.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-600 gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_n1.err
cat gpurun_out/summary.txt
