cd /root/repo
timeout 60 python -m pytest tests/test_pipeline.py tests/test_gpu_graphs.py -q -m gpu -k "mixed or pipeline" > gpurun_out/f6_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/f6_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f6_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/f6_smoke.log
