#!/bin/bash
# round 2, GPU call S: interleaved pair-mode chain kernel -- correctness, stress, timeline, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -8
python scripts/split_timeline.py run chain > gpurun_out/s_timeline_chain.txt 2>&1; tail -75 gpurun_out/s_timeline_chain.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-crop --no-e2e --no-fast-mode > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s_bench.json')); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['clocks'])"
