#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "split" > gpurun_out/m_tc.log 2>&1; echo "tc rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q > gpurun_out/m_par.log 2>&1; echo "parity+stress rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-crop > gpurun_out/m_bench1.json 2> gpurun_out/m_bench1.err; echo "bench rc=$?"
tail -4 gpurun_out/m_tc.log; tail -4 gpurun_out/m_par.log; python -c "
import json;d=json.loads(open('gpurun_out/m_bench1.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['kernel_ms'],d['roofline']['executed_frac'])"
