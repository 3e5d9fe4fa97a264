#!/bin/bash
# round 2, GPU call O: secondary configs in the default precision, compute-sanitizer on the tensor-core kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/o_build.log 2>&1
timeout 900 python scripts/bench_configs.py all > gpurun_out/o_configs.jsonl 2> gpurun_out/o_configs.err; echo "configs rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "split_seg_matches_fp64 or split_chain_trunks or split_tail_pair" > gpurun_out/o_memcheck_split.log 2>&1; echo "memcheck split rc=$?"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_crop.py tests/test_trackops.py tests/test_gpu_train.py -m gpu -x -q > gpurun_out/o_memcheck_misc.log 2>&1; echo "memcheck misc rc=$?"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "test_split_chain_trunks_match_fp64 or test_chain_maxpool_trunks" > gpurun_out/o_racecheck.log 2>&1; echo "racecheck rc=$?"
cat gpurun_out/o_configs.jsonl | cut -c1-400; tail -5 gpurun_out/o_memcheck_split.log; tail -5 gpurun_out/o_memcheck_misc.log; tail -8 gpurun_out/o_racecheck.log
