#!/bin/bash
# round 2, GPU call U (8 GPUs): strong-scaling bench at N = 8 (default flags, as the driver runs it), 4, 2, 1
mkdir -p gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 "$@"; }
run 8 > gpurun_out/u_bench8.json 2> gpurun_out/u_bench8.err; echo "n8 rc=$?"
run 4 --no-crop --no-fast-mode > gpurun_out/u_bench4.json 2> gpurun_out/u_bench4.err; echo "n4 rc=$?"
run 2 --no-crop --no-fast-mode > gpurun_out/u_bench2.json 2> gpurun_out/u_bench2.err; echo "n2 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-crop --no-fast-mode --no-cpu-baseline > gpurun_out/u_bench1.json 2> gpurun_out/u_bench1.err; echo "n1 rc=$?"
for n in 8 4 2 1; do python -c "
import json; d=json.load(open('gpurun_out/u_bench$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else None, d['clocks'], d['scaling'])"; done
tail -3 gpurun_out/u_bench8.err
