cd /root/repo
AL3D_CUDA_PROFILER_RANGE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:"split_(chain_pair|tail|chain)_kernel" -c 3 --csv --page raw --log-file gpurun_out/f3_ncu_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-mode --no-crop --no-configs > gpurun_out/f3_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/f3_ncu_full.log
