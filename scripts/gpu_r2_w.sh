#!/bin/bash
# crop: parity tests, per-stage timing of the 200-frame sweep, phases of the end-to-end sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crop.py tests/test_sweep.py tests/test_pipeline.py tests/test_trackops.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
timeout 300 python scripts/bench_configs.py crop 2>gpurun_out/w_crop.err | tee gpurun_out/w_crop.json
timeout 300 python scripts/sweep_phases.py 2>&1 | tail -9 | tee gpurun_out/w_sweep_phases.txt
