#!/bin/bash
# round 2, GPU call I: full GPU suite (crop v2, fused heads), crop bench, bench N=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python scripts/bench_configs.py crop > gpurun_out/i_crop.json 2> gpurun_out/i_crop.err; echo "crop rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench1.json 2> gpurun_out/i_bench1.err; echo "bench rc=$?"
tail -6 gpurun_out/i_tests.log; cat gpurun_out/i_crop.json; tail -3 gpurun_out/i_crop.err; tail -c 400 gpurun_out/i_bench1.err
