#!/bin/bash
# round 2, final call: full GPU suite, smoke, default bench + reference arm, ncu launch list of the timed region
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
AL3D_CUDA_PROFILER_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/final_ncu.log 2>&1; echo "ncu list rc=$?"
cut -c1-300 gpurun_out/final_bench.json; cut -c1-400 gpurun_out/final_ref.json
