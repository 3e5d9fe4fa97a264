#!/bin/bash
# crop_hits_kernel ablations (diagnostic builds; results are wrong on purpose, only the time of the hits pass is read)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/v_build.log 2>&1
cat > /tmp/time_hits.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
crop = importlib.import_module("3dal_pytorch_b200.crop")
synth = importlib.import_module("3dal_pytorch_b200.synth")
frames = synth.lidar_frames(200, seed=3)
plan = crop.CropPlan([torch.from_numpy(f["points"]).cuda() for f in frames], [crop.detector_to_waymo(f["det_boxes"]) for f in frames],
                     [f["pose"] for f in frames])
plan.grid()
for _ in range(3): plan.hits_pass()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): plan.hits_pass()
e1.record(); torch.cuda.synchronize()
print("hits_pass ms %.4f" % (e0.elapsed_time(e1) / 20))
PY
: > gpurun_out/abl.txt
for flags in "$@"; do
  cd 3dal_pytorch_b200
  OBJS=$(ls csrc/_obj/*.o | grep -v stress | grep -v crop.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I../include $flags -c csrc/crop.cu -o /tmp/crop_v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libal3d_cropv.so $OBJS /tmp/crop_v.o -lcuda
  cd ..
  echo "variant $flags: $(AL3D_LIB=libal3d_cropv.so python /tmp/time_hits.py 2>&1 | tail -1)" >> gpurun_out/abl.txt
done
rm -f 3dal_pytorch_b200/libal3d_cropv.so
cat gpurun_out/abl.txt
