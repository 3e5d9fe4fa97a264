#!/bin/bash
# crop: parity tests, per-stage timing of the 200-frame sweep, optional stage statistics (diagnostic build)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crop.py tests/test_sweep.py tests/test_pipeline.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
timeout 300 python scripts/bench_configs.py crop 2>gpurun_out/w_crop.err | tee gpurun_out/w_crop.json
