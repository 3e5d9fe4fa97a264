#!/bin/bash
# round 2, GPU call T: full GPU suite, smoke, default bench + reference arm, ncu launch list + full capture of the pipelined split kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/t_bench1.json 2> gpurun_out/t_bench1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_ref.json 2> gpurun_out/t_ref.err; echo "ref rc=$?"
AL3D_CUDA_PROFILER_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/t_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/t_ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:split_ -s 6 -c 3 -o gpurun_out/t_split_full -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/t_ncu_f.log 2>&1; echo "ncu full rc=$?"
cut -c1-400 gpurun_out/t_bench1.json; cut -c1-400 gpurun_out/t_ref.json
