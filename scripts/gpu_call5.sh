#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_dm_dropin.py -x -q -m gpu -k "sidecar or worker or trimming" > gpurun_out/t_rec.log 2>&1; echo "record tests rc=$?"; tail -4 gpurun_out/t_rec.log
df -h /tmp | tail -1
timeout 300 python scripts/dropin_e2e.py 4 2000 256 /tmp/dropin_e2e --trim > gpurun_out/dropin_e2e.log 2>&1; echo "dropin e2e rc=$?"; tail -6 gpurun_out/dropin_e2e.log
timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_final.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['single_pd_config_3'], d['clocks'])"
