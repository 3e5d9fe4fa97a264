#!/bin/bash
# crop: parity tests, then ablation timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crop.py tests/test_sweep.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
bash scripts/gpu_r2_abl.sh "$@"
