#!/bin/bash
# round 2, GPU call Q: pipelined split_tail_kernel -- correctness, stress, timeline, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -8
python scripts/split_timeline.py run > gpurun_out/q_timeline.txt 2>&1; tail -62 gpurun_out/q_timeline.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-crop --no-e2e --no-fast-mode > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; cut -c1-700 gpurun_out/q_bench.json
