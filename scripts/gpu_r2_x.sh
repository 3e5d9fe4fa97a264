#!/bin/bash
# ncu --set full of the crop kernels (one sweep of 200 frames)
mkdir -p gpurun_out
cat > /tmp/crop_once.py <<'PY'
import importlib, sys, torch
sys.path.insert(0, ".")
crop = importlib.import_module("3dal_pytorch_b200.crop")
synth = importlib.import_module("3dal_pytorch_b200.synth")
frames = synth.lidar_frames(200, seed=3)
plan = crop.CropPlan([torch.from_numpy(f["points"]).cuda() for f in frames], [crop.detector_to_waymo(f["det_boxes"]) for f in frames],
                     [f["pose"] for f in frames])
plan.run(); plan.run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
plan.run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:crop_ -o gpurun_out/x_crop python /tmp/crop_once.py > gpurun_out/x_ncu.log 2>&1
tail -3 gpurun_out/x_ncu.log
ls -la gpurun_out/x_crop.ncu-rep
