#!/bin/bash
# round 2, GPU call C: new bench line at N=1 (bf16x3 headline + fast mode + crop), ncu launch list + full capture of the split kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench1.json 2> gpurun_out/c_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --tracks 1024 --no-cpu-baseline --no-crop > gpurun_out/c_bench_1024.json 2> gpurun_out/c_bench_1024.err; echo "bench1024 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/c_ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:split_ -s 6 -c 3 -o gpurun_out/c_split_full -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-crop --no-fast-mode > gpurun_out/c_ncu_f.log 2>&1; echo "ncu full rc=$?"
tail -c 600 gpurun_out/c_bench1.json
