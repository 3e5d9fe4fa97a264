#!/bin/bash
# crop_hits_kernel build variants timed on the box: each argument is a quoted list of -D flags
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/v_build.log 2>&1
: > gpurun_out/v_variants.txt
for flags in "$@"; do
  cd 3dal_pytorch_b200
  OBJS=$(ls csrc/_obj/*.o | grep -v stress | grep -v crop.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I../include $flags -c csrc/crop.cu -o /tmp/crop_v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libal3d_cropv.so $OBJS /tmp/crop_v.o -lcuda
  cd ..
  echo "variant $flags" >> gpurun_out/v_variants.txt
  AL3D_LIB=libal3d_cropv.so python scripts/bench_configs.py crop 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})" >> gpurun_out/v_variants.txt
done
rm -f 3dal_pytorch_b200/libal3d_cropv.so
cat gpurun_out/v_variants.txt
