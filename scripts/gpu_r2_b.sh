#!/bin/bash
# round 2, GPU call B: split-precision kernels: tensor-core unit tests, parity suite, per-tensor errors, timing
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "split" > gpurun_out/b_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/b_tc.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 300 python scripts/measure_parity.py --out gpurun_out/b_parity.jsonl > gpurun_out/b_parity.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision bf16x3 > gpurun_out/b_bench_x3.json 2> gpurun_out/b_bench_x3.err; echo "bench rc=$?"
tail -3 gpurun_out/b_tc.log; tail -5 gpurun_out/b_tests.log
