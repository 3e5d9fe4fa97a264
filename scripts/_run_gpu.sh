cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/f4_bench_n2.json 2> gpurun_out/f4_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/f4_bench_n2.json | head -c 1500; tail -3 gpurun_out/f4_bench_n2.err
