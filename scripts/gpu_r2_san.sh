#!/bin/bash
# compute-sanitizer over the crop tests: memcheck (global / shared out-of-bounds, misaligned), racecheck (the per-warp shared
# queues and the TMA staging buffers), synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool" >> gpurun_out/san_crop.txt
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_crop.py -x -q -m gpu -k "not waymo_sized" 2>&1 | grep -v "^$" | tail -8 >> gpurun_out/san_crop.txt
done
cat gpurun_out/san_crop.txt
