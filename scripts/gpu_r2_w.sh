#!/bin/bash
# crop v3 (three-stage queues): parity tests, per-stage timing of the 200-frame sweep, stage statistics (diagnostic build)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_crop.py tests/test_sweep.py -x -q -m gpu > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
timeout 300 python scripts/bench_configs.py crop 2>gpurun_out/w_crop.err | tee gpurun_out/w_crop.json
if [ "$1" = "stats" ]; then
  cd 3dal_pytorch_b200
  OBJS=$(ls csrc/_obj/*.o | grep -v stress | grep -v crop.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I../include -DCROP_STATS -c csrc/crop.cu -o /tmp/crop_s.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libal3d_crops.so $OBJS /tmp/crop_s.o -lcuda
  cd ..
  AL3D_LIB=libal3d_crops.so python - <<'PY' 2>&1 | tee gpurun_out/w_stats.txt
import ctypes, importlib, sys, torch
sys.path.insert(0, ".")
crop = importlib.import_module("3dal_pytorch_b200.crop")
synth = importlib.import_module("3dal_pytorch_b200.synth")
_lib = importlib.import_module("3dal_pytorch_b200._lib")
frames = synth.lidar_frames(20, seed=3)
plan = crop.CropPlan([torch.from_numpy(f["points"]).cuda() for f in frames], [crop.detector_to_waymo(f["det_boxes"]) for f in frames],
                     [f["pose"] for f in frames])
l = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
plan.run(); l.al3d_crop_stats(out, 1)
plan.run(); l.al3d_crop_stats(out, 1)
names = ["candidates", "pairs", "expand batches", "serial expand batches", "test batches", "test batches w/ exact", "pairs in margin", "hits"]
n = int(plan.pts_all.shape[0])
print("points", n)
for k, v in zip(names, out): print("%-24s %10d  %.4f per point" % (k, v, v / n))
PY
  rm -f 3dal_pytorch_b200/libal3d_crops.so
fi
