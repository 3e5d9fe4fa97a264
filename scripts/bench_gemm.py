"""Time the split-precision tensor-core GEMMs (csrc/gemm_split.cu) on the layer shapes of the training step
(64 objects x 4096 points) next to the fp32 SIMT kernels they replace."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tr = importlib.import_module("3dal_pytorch_b200.train")
ops = importlib.import_module("3dal_pytorch_b200.ops")
DEV = "cuda:0"


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    M = 64 * 4096
    for N, K in [(64, 64), (128, 64), (1024, 128), (512, 64), (256, 512), (128, 256), (128, 128), (128, 1024), (512, 256)]:
        a = torch.randn((M, K), device=DEV)
        w = torch.randn((N, K), device=DEV) / K ** 0.5
        dy = torch.randn((M, N), device=DEV)
        dw = torch.empty((N, K), device=DEV)
        out = torch.empty((M, N), device=DEV)
        row = {"M": M, "N": N, "K": K, "gflop": 2.0 * M * N * K / 1e9}
        row["nt_x3_ms"] = timed(lambda: tr.linear_split(a, w, out=out, parts=2))
        row["nt_x6_ms"] = timed(lambda: tr.linear_split(a, w, out=out, parts=3))
        row["nt_f32_ms"] = timed(lambda: ops.linear(a, w, None, act=ops.ACT_NONE, out=out))
        row["tn_x3_ms"] = timed(lambda: tr.wgrad_split(dy, a, dw, parts=2))
        row["tn_x6_ms"] = timed(lambda: tr.wgrad_split(dy, a, dw, parts=3))
        tr.set_gemm_mode("f32")
        row["tn_f32_ms"] = timed(lambda: tr.wgrad(dy, a, dw))
        row["nt_x3_tflops"] = row["gflop"] / row["nt_x3_ms"]
        row["nt_x6_tflops"] = row["gflop"] / row["nt_x6_ms"]
        print(json.dumps(row))


if __name__ == "__main__":
    main()
