#!/bin/bash
# round 2, GPU call F: training, crop and trackops tests; launch list in the profiler range
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_crop.py tests/test_trackops.py -m gpu -q -s > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"
AL3D_CUDA_PROFILER_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/f_ncu_b.log 2>&1; echo "ncu list rc=$?"
grep -n "worst\|passed\|failed\|^FAILED" gpurun_out/f_tests.log | tail -14
