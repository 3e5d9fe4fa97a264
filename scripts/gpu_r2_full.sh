#!/bin/bash
# full GPU suite, smoke, default bench line (N=1), crop bench + ncu of the crop kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/full_tests.log
tail -5 gpurun_out/full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/full_smoke.log
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/full_bench.json
