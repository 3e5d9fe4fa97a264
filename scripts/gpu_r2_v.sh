#!/bin/bash
# round 2, GPU call V: tensor-core training GEMMs -- tests, step time in both GEMM modes, ncu launch list of one x3 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/bench_train.py --gemm f32 > gpurun_out/v_train_f32.json 2> gpurun_out/v_train_f32.err; tail -1 gpurun_out/v_train_f32.json | cut -c1-400
timeout 300 python scripts/bench_train.py --gemm x3 --no-cpu > gpurun_out/v_train_x3.json 2> gpurun_out/v_train_x3.err; tail -1 gpurun_out/v_train_x3.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/v_train_launches.csv \
    python scripts/bench_train.py --gemm x3 --no-cpu --steps 2 --warmup 2 > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
