#!/bin/bash
# One gpurun call: new kNN tests, whole GPU suite, config-3 timings, bench line, ncu of the selection kernel.
# Every stage writes under gpurun_out/ and is bounded by its own timeout, so a clamped call still leaves results.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_knn_fused.py -x -q -m gpu > gpurun_out/t_knn.log 2>&1; echo "knn tests rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/t_knn.log
timeout 420 python -m pytest tests -q -m gpu --deselect tests/test_gpu_knn_fused.py --durations=8 > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/summary.txt
tail -14 gpurun_out/t_all.log
timeout 120 python scripts/configs_check.py c3 > gpurun_out/c3.log 2>&1; echo "c3 rc=$?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/c3.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 240 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-400 gpurun_out/bench_n1.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_knn_select" -c 1 -o gpurun_out/knn_select_full -f python scripts/one_pd.py 5000 256 1 100 > gpurun_out/ncu_knn.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/summary.txt
ncu -i gpurun_out/knn_select_full.ncu-rep --page details > gpurun_out/knn_select_details.txt 2>&1
cat gpurun_out/summary.txt
