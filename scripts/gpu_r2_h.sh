#!/bin/bash
# round 2, GPU call H: crop v2 tests + crop bench + per-stage times + ncu of the new hits kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1
timeout 900 python -m pytest tests/test_crop.py tests/test_sweep.py tests/test_trackops.py -m gpu -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python scripts/bench_configs.py crop > gpurun_out/h_crop.json 2> gpurun_out/h_crop.err; echo "crop rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_ -s 12 -c 5 -o gpurun_out/h_crop_full -f python scripts/bench_configs.py crop --frames 200 > gpurun_out/h_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/h_tests.log; cat gpurun_out/h_crop.json; tail -3 gpurun_out/h_crop.err
