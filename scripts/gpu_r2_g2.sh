#!/bin/bash
# ncu --set full of the compaction / gather kernels inside the bench step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'mask_compact|gather_fg' -c 4 -o gpurun_out/g_gather python bench.py --no-crop --steps 2 --warmup 1 > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
