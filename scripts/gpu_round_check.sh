#!/bin/bash
# One gpurun call that re-validates the round on a B200: whole GPU suite, smoke, config-3 / DM-chain / S2 timings,
# the bench line.  Every stage writes under gpurun_out/ and is bounded by its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?" | tee gpurun_out/summary.txt
tail -14 gpurun_out/t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 150 python scripts/configs_check.py c3 > gpurun_out/c3.log 2>&1; echo "c3 rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/c3.log
timeout 60 python scripts/dm_chain.py 2000 0 3 > gpurun_out/dm_chain.log 2>&1; tail -1 gpurun_out/dm_chain.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-300 gpurun_out/bench_n1.json
cat gpurun_out/summary.txt
