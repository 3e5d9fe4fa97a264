#!/bin/bash
# round 2, GPU call E: training tests, crop tests, then the whole GPU suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_crop.py tests/test_losses.py -m gpu -q -s > gpurun_out/e_train.log 2>&1; echo "train+crop rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e_tests.log 2>&1; echo "all rc=$?"
grep -n "worst\|passed\|failed" gpurun_out/e_train.log | tail -12; tail -5 gpurun_out/e_tests.log
