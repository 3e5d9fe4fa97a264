#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/k_build.log 2>&1
timeout 900 python -m pytest tests/test_crop.py tests/test_sweep.py tests/test_trackops.py -m gpu -q > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python scripts/bench_configs.py crop > gpurun_out/k_crop.json 2> gpurun_out/k_crop.err; echo "crop rc=$?"
tail -4 gpurun_out/k_tests.log; cat gpurun_out/k_crop.json
