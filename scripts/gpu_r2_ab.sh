#!/bin/bash
# round 2, GPU call AB (8 GPUs): default bench at N = 8 (as the driver runs it, with the train_step sub-line) and the
# training step alone at N = 8 / 4 / 2 / 1 (BASELINE configs[4])
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
tr 8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/ab_bench8.json 2> gpurun_out/ab_bench8.err; echo "bench8 rc=$?"
for n in 8 4 2; do tr $n scripts/bench_train.py --gemm x6 --no-cpu 2>/dev/null | tail -1; done > gpurun_out/ab_train_scaling.jsonl
timeout 300 python scripts/bench_train.py --gemm x6 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/ab_train_scaling.jsonl
wc -l gpurun_out/ab_bench8.json; cut -c1-260 gpurun_out/ab_train_scaling.jsonl
python -c "
import json; d=json.loads(open('gpurun_out/ab_bench8.json').read()); print(d['value'], d['e2e']['value'], json.dumps(d['train_step'])[:700])"
