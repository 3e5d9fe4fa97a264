#!/usr/bin/env python
"""BASELINE.json configs[4]: static one-box training step (forward with batch-statistics BatchNorm + dropout, fused loss,
backward, ONE flat-bucket NCCL gradient all-reduce, fused Adam), 64 tracks x 4096 points per GPU.

    python scripts/bench_train.py [--steps 10] [--batch 64]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py   # N GPUs

Prints one JSON line (rank 0): ms per step (CUDA events, max over ranks), objects/s, the all-reduce share, and -- at
N = 1 -- the same step of the reference algorithm (oracle autograd on the host cores) on a bounded sample."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__ as ge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--model", default="static_one", choices=["static_one", "static_two", "dynamic"],
                    help="static_one: the fused TrainStep; static_two / dynamic: AutogradTrainStep (the models' own autograd Functions)")
    ap.add_argument("--gemm", default="x6", choices=["x6", "x3", "f32"],
                    help="layer GEMMs: bf16x6 tensor-core kernels (fp32-grade, default), bf16x3 tensor-core kernels, fp32 SIMT kernels")
    args = ap.parse_args()
    ge.build()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    synth = importlib.import_module("3dal_pytorch_b200.synth")
    sm = importlib.import_module("3dal_pytorch_b200.static_model")
    tr = importlib.import_module("3dal_pytorch_b200.train")
    spec = importlib.import_module("3dal_pytorch_b200.spec")
    bs, n = args.batch, args.points
    tr.set_gemm_mode(args.gemm)
    g = torch.Generator(device=dev); g.manual_seed(5 + rank)
    if args.model == "static_one":
        sd = synth.random_state_dict("static_one", seed=synth.REFERENCE_SEED)
        model = sm.StaticModelOneBoxEst().to(dev).train()
        model.load_state_dict(sd)
        step = tr.TrainStep(model, lr=1e-3, weight_decay=1e-4, dropout_p=0.5)
        d = synth.static_tracks_device(bs, n=n, seed=100 + rank, device=dev)
        pts, init_box = d["pts_pm"].transpose(2, 1), d["init_box"]
        labels = ((torch.rand((bs, n), device=dev, generator=g) < 0.3).float(), torch.randn((bs, 3), device=dev, generator=g) * 0.3,
                  torch.randint(0, 12, (bs,), device=dev, generator=g), torch.randn((bs,), device=dev, generator=g) * 0.1,
                  torch.randint(0, 3, (bs,), device=dev, generator=g), torch.randn((bs, 3), device=dev, generator=g) * 0.2)
        run_step = lambda: step.step(pts, init_box, labels)
        grads = step.grads
    else:
        import numpy as np
        dm = importlib.import_module("3dal_pytorch_b200.dynamic_model")
        losses_mod = importlib.import_module("3dal_pytorch_b200.losses")
        sd = synth.random_state_dict(args.model, seed=synth.REFERENCE_SEED)
        if args.model == "dynamic":
            t = synth.dynamic_tracks(min(bs, 64), seed=100 + rank)
            rep = -(-bs // t["pts_pm"].shape[0])
            pts = torch.from_numpy(np.tile(t["pts_pm"], (rep, 1, 1))[:bs]).to(dev).transpose(2, 1)
            aux = torch.from_numpy(np.tile(t["box_sm"], (rep, 1, 1))[:bs]).to(dev).transpose(2, 1)
            gt = torch.from_numpy(np.tile(t["bbox_gt"], (rep, 1))[:bs]).to(dev)
            model, crit = dm.DynamicModel().to(dev), losses_mod.DynamicModelLoss()
        else:
            d = synth.static_tracks_device(bs, n=n, seed=100 + rank, device=dev)
            pts, aux, gt = d["pts_pm"].transpose(2, 1), d["init_box"], d["bbox_gt"]
            model, crit = sm.StaticModelTwoBoxEst().to(dev), losses_mod.FrustumPointNetLossTwoBoxEst()
        model.load_state_dict(sd)
        n = pts.shape[2]
        labels = [(torch.rand((bs, n), device=dev, generator=g) < 0.3).float(), torch.randn((bs, 3), device=dev, generator=g) * 0.3,
                  torch.randint(0, 12, (bs,), device=dev, generator=g), torch.randn((bs,), device=dev, generator=g) * 0.1,
                  torch.randint(0, 3, (bs,), device=dev, generator=g), torch.randn((bs, 3), device=dev, generator=g) * 0.2]
        step = tr.AutogradTrainStep(model, crit, lr=1e-3, weight_decay=1e-4)
        run_step = lambda: step.step((pts, aux, gt), labels)
        grads = step.grads
    for _ in range(args.warmup):
        run_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(args.steps):
        losses.append(run_step()["total_loss"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # the all-reduce alone (same bucket), for its share of the step
    ar_ms = 0.0
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        a0.record()
        for _ in range(10):
            dist.all_reduce(grads.flat)
        a1.record(); torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 10
        t = torch.tensor([ms, ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar_ms = float(t[0]), float(t[1])
    if rank == 0:
        flop = 3.0 * spec.flops_per_object(args.model, n)          # fwd + dgrad + wgrad on the factored forward count
        line = {"bench": "train_step", "model": args.model, "n_gpus": world, "batch_per_gpu": bs, "points": n,
                "ms_per_step": ms, "objects_per_s": world * bs / (ms * 1e-3), "model_tflops_per_gpu": bs * flop / (ms * 1e-3) / 1e12,
                "grad_bucket_floats": int(grads.flat.numel()), "allreduce_ms": ar_ms,
                "loss_first_last": [float(losses[0]), float(losses[-1])], "dtype": {"f32": "f32 (SIMT GEMMs)", "x3": "bf16x3 tensor-core GEMMs (fp32 in / out, fp32 accumulate)",
                          "x6": "bf16x6 tensor-core GEMMs (fp32 in / out, fp32 accumulate)"}[args.gemm], "gemm": args.gemm,
                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
        if world == 1 and not args.no_cpu and args.model == "static_one":
            from oracle import train as otrain
            cb = 8
            torch.set_num_threads(os.cpu_count() or 1)
            cp = d["pts_pm"][:cb].transpose(2, 1).cpu().contiguous()
            cl = tuple(t[:cb].cpu() for t in labels)
            otrain.static_one_step(sd, cp, init_box[:cb].cpu(), cl)
            t0 = time.perf_counter()
            otrain.static_one_step(sd, cp, init_box[:cb].cpu(), cl)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"objects_per_s": cb / dt, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "%d tracks x %d pts, forward + autograd backward of the fp32 oracle" % (cb, n)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
