#!/bin/bash
# gather kernels: parity tests + the bench line (gather object)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_pipeline.py tests/test_gpu_graphs.py -x -q -m gpu > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -4 gpurun_out/g_tests.log
timeout 900 python bench.py --no-crop > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['kernel_ms']); print(json.dumps(d['gather'], indent=1)); print(d['e2e'])"
