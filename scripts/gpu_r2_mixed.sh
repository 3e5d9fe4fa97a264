# The gpurun commands behind the mixed-mode evidence under profiles/ (r2_*mixed*); run from the repo root on the GPU box.
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/f_tests.log 2>&1                                   # profiles/r2_mixed_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1                  # profiles/r2_mixed_smoke.log
python bench.py > gpurun_out/f_bench.json                                                         # profiles/r2_bench_n1_mixed.json (first mixed build)
python bench.py --steps 10 --no-configs --no-crop --no-cpu-baseline > gpurun_out/f_bench2.json    # profiles/r2_bench_n1_mixed_final.json
AL3D_MIXED_D2=1 python bench.py --precision mixed --no-configs --no-crop --no-cpu-baseline --no-e2e --steps 5 \
    > gpurun_out/f_bench_d2x1.json                                                                # profiles/r2_bench_n1_mixed_d2_single_f16.json
python scripts/measure_parity.py --precisions mixed --out gpurun_out/f_parity_mixed.jsonl         # appended to profiles/r2_parity_per_tensor.jsonl
AL3D_CUDA_PROFILER_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-mode --no-crop \
    --no-configs                                                                                  # profiles/r2_mixed_launches_timed_region.csv
AL3D_CUDA_PROFILER_RANGE=1 ncu --profile-from-start off --set full --clock-control none -k regex:"split_(chain_pair|tail|chain)_kernel" \
    -c 3 --csv --page raw --log-file gpurun_out/f_ncu_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    --no-fast-mode --no-crop --no-configs                                                         # profiles/r2_mixed_ncu_full.csv
# 2 GPUs (gpurun --gpus 2): profiles/r2_bench_n2_mixed.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    --steps 10 --warmup 3 > gpurun_out/f_bench_n2.json
