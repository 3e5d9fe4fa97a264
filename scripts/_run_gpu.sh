cd /root/repo
timeout 200 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -q -m gpu -k "mixed" -x > gpurun_out/m_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/m_tests.log
timeout 120 python bench.py --precision mixed --no-configs --no-crop --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/m_bench_d2x2.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
AL3D_MIXED_D2=1 timeout 120 python bench.py --precision mixed --no-configs --no-crop --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/m_bench_d2x1.json 2>> gpurun_out/m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("m_bench_d2x2","m_bench_d2x1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["clocks"], d["roofline"]["executed_frac"], d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/m_bench.err
