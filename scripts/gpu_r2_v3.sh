#!/bin/bash
# crop: chunk-size sweep (host-side knob)
mkdir -p gpurun_out
: > gpurun_out/v3_variants.txt
for ch in "$@"; do
  echo "chunk $ch" >> gpurun_out/v3_variants.txt
  AL3D_CROP_CHUNK=$ch python scripts/bench_configs.py crop 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})" >> gpurun_out/v3_variants.txt
done
cat gpurun_out/v3_variants.txt
