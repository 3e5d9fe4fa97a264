#!/bin/bash
# crop: parity tests + ncu of the crop kernels in one call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crop.py tests/test_sweep.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
timeout 300 python scripts/bench_configs.py crop 2>gpurun_out/w_crop.err | tee gpurun_out/w_crop.json
bash scripts/gpu_r2_x.sh
