#!/bin/bash
# round 2, GPU call D: training-step tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_losses.py -m gpu -q -s > gpurun_out/d_train.log 2>&1; echo "train rc=$?"
tail -40 gpurun_out/d_train.log
