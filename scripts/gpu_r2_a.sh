#!/bin/bash
# round 2, GPU call A: GPU test suite (incl. protocol stress), parity per tensor, bench x3 (deadlock soak)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 300 python scripts/measure_parity.py --out gpurun_out/a_parity.jsonl > gpurun_out/a_parity.log 2>&1
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench$i.json 2> gpurun_out/a_bench$i.err; echo "bench$i rc=$?"
done
tail -3 gpurun_out/a_tests.log
