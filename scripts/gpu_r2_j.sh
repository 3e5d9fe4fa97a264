#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/j_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python scripts/bench_configs.py crop > gpurun_out/j_crop.json 2> gpurun_out/j_crop.err; echo "crop rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_ -s 12 -c 5 -o gpurun_out/j_crop_full -f python scripts/bench_configs.py crop --frames 200 > gpurun_out/j_ncu.log 2>&1; echo "ncu rc=$?"
tail -6 gpurun_out/j_tests.log; cat gpurun_out/j_crop.json
