#!/bin/bash
# crop kernel variants, built on the box: (min blocks per SM, cell mapping)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/l_build.log 2>&1
cd 3dal_pytorch_b200
OBJS=$(ls csrc/_obj/*.o | grep -v stress | grep -v crop.o)
for v in "3 0" "4 0" "2 0"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I../include -DCROP_MINB=$1 $( [ "$2" = "1" ] && echo -DCROP_CELL_FLOOR ) -c csrc/crop.cu -o /tmp/crop_v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libal3d_cropv.so $OBJS /tmp/crop_v.o -lcuda
  cd ..
  echo "variant minb=$1 floor=$2" >> gpurun_out/l_variants.txt
  AL3D_LIB=libal3d_cropv.so python scripts/bench_configs.py crop 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms'], d['stage_ms'])" >> gpurun_out/l_variants.txt
  cd 3dal_pytorch_b200
done
cd ..; rm -f 3dal_pytorch_b200/libal3d_cropv.so
cat gpurun_out/l_variants.txt
