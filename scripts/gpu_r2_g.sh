#!/bin/bash
# round 2, GPU call G (2 GPUs): remaining tests, training bench N=1/2, strong-scaling bench N=2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_crop.py tests/test_trackops.py -m gpu -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python scripts/bench_train.py --steps 10 > gpurun_out/g_train1.json 2> gpurun_out/g_train1.err; echo "train1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_train.py --steps 10 > gpurun_out/g_train2.json 2> gpurun_out/g_train2.err; echo "train2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g_bench2.json 2> gpurun_out/g_bench2.err; echo "bench2 rc=$?"
tail -4 gpurun_out/g_tests.log; tail -c 700 gpurun_out/g_train1.json; tail -c 500 gpurun_out/g_train2.json; tail -c 300 gpurun_out/g_train1.err
