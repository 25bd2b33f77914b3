#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_s2_tessellation.py -x -q -m gpu > gpurun_out/t_s2.log 2>&1; echo "s2 tests rc=$?"; tail -4 gpurun_out/t_s2.log
timeout 200 python scripts/s2_timing.py 1000000 4071 > gpurun_out/s2_timing.log 2>&1; echo "s2 timing rc=$?"; cat gpurun_out/s2_timing.log | tail -3
