cd /root/repo
timeout 38 python bench.py --steps 10 --no-configs --no-crop --no-cpu-baseline > gpurun_out/f7_bench.json 2> gpurun_out/f7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/f7_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["kernel_ms"], d["clocks"], d["e2e"])
    print({k:(v["value"],v["ms_per_step"]) for k,v in d.items() if k in ("bf16x3_mode","fast_mode") and v})
except Exception as e: print("ERR", e)
PY
