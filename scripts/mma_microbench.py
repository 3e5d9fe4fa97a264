"""Development aid: tcgen05.mma issue / execution rates on this GPU (see csrc/microbench.cu).
Usage (GPU box): python scripts/mma_microbench.py > gpurun_out/mma_microbench.txt"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as ge

ge.build()
lib = importlib.import_module("3dal_pytorch_b200._lib")

out = torch.zeros(8, dtype=torch.int64, device="cuda:0")
src = torch.randint(0, 255, (1 << 20,), dtype=torch.uint8, device="cuda:0")
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
print("N mode commit_every ctas background | issue_cycles total_cycles cycles_per_mma first8_issue bg_blocks bg_loads")
n = 2048
for ctas in (1, n_sm):
    for bg in (0, 4, 1, 2, 3):
        for mode in (0, 1):
            for N in (64, 128, 256):
                for ce in (0, 8, 4):
                    if bg not in (0, 4) and ce != 8:
                        continue
                    for _ in range(2):
                        out.zero_()
                        lib.check(lib.lib().al3d_mma_microbench(N, n, ce, mode, ctas, bg, src.data_ptr(), out.data_ptr(), None),
                                  "microbench")
                    torch.cuda.synchronize()
                    o = out.cpu().tolist()
                    print(N, "SS" if mode == 0 else "TS", ce, ctas, bg, "|", o[0], o[1], "%.1f" % (o[1] / n), o[2], o[3], o[4])
