#!/usr/bin/env python
"""Benchmark of the auto-labeling hot path (metric of BASELINE.json: auto-labeled objects/sec).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2]): static one-box Frustum-PointNet forward + box decode on
`--tracks` synthetic tracks x 4096 points PER GPU (weak scaling; independent tracks, no data-path
collective -- only the final all_gather of the (tracks,7) boxes), random-init BN-calibrated weights,
bf16 tensor-core mode.  One "step" = one pass over the whole batch.  Inputs (403 MB/step at the
default size) are larger than the 126 MB L2, so no explicit L2 flush is needed between steps.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

N_POINTS = 4096
CPU_SAMPLE_TRACKS = 32


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def _traffic(kernel):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(kernel)
    return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.rows, self.proc = [], None
        try:
            ident = str(torch.cuda.get_device_properties(dev).uuid)
            if not ident.startswith("GPU-"):
                ident = "GPU-" + ident
        except Exception:
            ident = str(torch.cuda.current_device())
        self.cmd = ["nvidia-smi", "-i", ident, "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50"]

    def start(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def wait_first_sample(self, keep_busy, timeout=5.0):
        """nvidia-smi takes a while to start on an 8-GPU box: keep the GPU under load (extra untimed steps)
        until the first sample has arrived, so that the timed region is covered."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            keep_busy()
            torch.cuda.synchronize()

    def mark(self):
        return time.time()

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r[1:] for r in self.rows if t_begin is None or (t_begin - 0.05 <= r[0] <= t_end + 0.1)]
        if not rows:                      # region shorter than the sampling period: use the loaded samples around it
            rows = [r[1:] for r in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[3:7]):
                if v == "Active":
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_objects_per_s(sd_cpu, min_seconds=10.0, max_iters=8):
    """The reference algorithm (oracle port, fp32 torch on the host cores) on a bounded sample."""
    from oracle import models
    synth = importlib.import_module("3dal_pytorch_b200.synth")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = synth.static_tracks(CPU_SAMPLE_TRACKS, n=N_POINTS, seed=1)
    pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1)
    init_box = torch.from_numpy(tr["init_box"])

    def step():
        out = models.static_one_forward(sd_cpu, pts, init_box, policy="strided")
        models.decode_box(out["center"], out["heading_scores"], out["heading_residuals"], out["size_scores"],
                          out["size_residuals"], init_box[:, 6])

    step()
    times = []
    t_all = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_all < min_seconds or len(times) < 2):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return CPU_SAMPLE_TRACKS / statistics.median(times), cores, times


def run_reference(args, rank):
    if rank != 0:
        return
    synth = importlib.import_module("3dal_pytorch_b200.synth")
    sd = synth.random_state_dict("static_one", seed=synth.REFERENCE_SEED)
    from oracle import models
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = synth.static_tracks(CPU_SAMPLE_TRACKS, n=N_POINTS, seed=1)
    pts = torch.from_numpy(tr["pts_pm"]).transpose(2, 1)
    init_box = torch.from_numpy(tr["init_box"])

    def step():
        out = models.static_one_forward(sd, pts, init_box, policy="strided")
        models.decode_box(out["center"], out["heading_scores"], out["heading_residuals"], out["size_scores"],
                          out["size_residuals"], init_box[:, 6])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = CPU_SAMPLE_TRACKS * args.steps / dt
    sample = "%d tracks x %d pts per step, oracle port (fp32 torch CPU restatement of the reference forward + decode)" % (
        CPU_SAMPLE_TRACKS, N_POINTS)
    line = {"impl": "reference", "metric": "auto-labeled objects/sec", "value": val, "unit": "objects/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "static one-box Frustum-PointNet forward + box decode, %d tracks x %d pts per GPU"
                                   " (BASELINE.json configs[2])" % (args.tracks, N_POINTS),
                       "tracks_per_gpu": args.tracks, "points": N_POINTS, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "objects/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=8192, help="tracks per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import __graft_entry__ as ge
    ge.build()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        # a short collective timeout: a protocol bug should end the run with an error, not sit until the driver's limit
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    synth = importlib.import_module("3dal_pytorch_b200.synth")
    sm = importlib.import_module("3dal_pytorch_b200.static_model")
    eb = importlib.import_module("3dal_pytorch_b200.engine_bf16")
    lib = importlib.import_module("3dal_pytorch_b200._lib")
    pipeline = importlib.import_module("3dal_pytorch_b200.pipeline")
    spec = importlib.import_module("3dal_pytorch_b200.spec")

    T = args.tracks
    sd = synth.random_state_dict("static_one", seed=synth.REFERENCE_SEED)
    model = sm.StaticModelOneBoxEst().to(dev).eval()
    model.precision = args.precision
    model.load_state_dict(sd)
    data = synth.static_tracks_device(T, n=N_POINTS, seed=1000 + rank, device=dev)
    pts = data["pts_pm"].transpose(2, 1)                  # strided (T,3,n) view, as the eval scripts pass it
    init_box = data["init_box"]
    # calibrate the segmentation head on a subsample so the mask / gather stages do real work
    with torch.no_grad():
        lg = model(pts[:256], init_box[:256], None)["logits"]
    synth.calibrate_seg_margin(sd, lg, fg_fraction=0.125)
    model.load_state_dict(sd)
    labeler = pipeline.StaticAutoLabeler(model, chunk_tracks=min(T, int(os.environ.get("AL3D_E2E_CHUNK", "2048"))),
                                         first_chunk_tracks=int(os.environ.get("AL3D_E2E_FIRST", "0")) or None)
    gathered = torch.empty((world * T, 7), device=dev, dtype=torch.float32) if world > 1 else None

    def local_step():
        return labeler.label_device(pts, init_box)

    def step():
        boxes = local_step()
        if world > 1:
            dist.all_gather_into_tensor(gathered, boxes)
        return boxes

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    fg = model(pts[:256], init_box[:256], None)["mask"].float().sum(1)

    # ---------------- timed region: inputs resident in HBM
    if os.environ.get("AL3D_CUDA_PROFILER_RANGE") == "1":      # for `ncu --profile-from-start off`
        torch.cuda.cudart().cudaProfilerStart()
    sampler = ClockSampler(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    # NOT `step`: every rank leaves this loop after its own number of iterations (when its own nvidia-smi has produced a
    # sample), and a collective inside it would be called a different number of times per rank -- a deadlock at N >= 4,
    # where nvidia-smi start-up times differ most.
    sampler.wait_first_sample(local_step)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = sampler.mark()
    eb.KERNEL_EVENTS = {}                   # per-kernel CUDA events and the launch count cover the K timed steps only
    launches0 = lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_end = sampler.mark()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_begin, t_end)
    if os.environ.get("AL3D_CUDA_PROFILER_RANGE") == "1":
        torch.cuda.cudart().cudaProfilerStop()
    launches = lib.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    eb.check_abort("bench timed region")      # a tensor-core kernel that gave up on a wait would make the number meaningless
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in eb.KERNEL_EVENTS.items()}
    eb.KERNEL_EVENTS = None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * T * args.steps / (ms * 1e-3)

    # ---------------- e2e: pinned host buffers in, boxes out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pts_host = data["pts_pm"].cpu().pin_memory()
        box_host = init_box.cpu().pin_memory()
        out_host = torch.empty((T, 7), dtype=torch.float32).pin_memory()
        for _ in range(2):
            labeler.label_host(pts_host, box_host, out_host)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            labeler.label_host(pts_host, box_host, out_host)
        f1.record()
        torch.cuda.synchronize()
        ems = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t0))
        eb.check_abort("bench e2e region")
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * T * args.steps / (ems * 1e-3), "unit": "objects/s",
               "h2d_bytes_per_step": int(pts_host.numel() * 4 + box_host.numel() * 4),
               "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": ems / args.steps}

    if rank == 0:
        peaks = _peaks()
        # dominant kernel: segmentation pass 2 (dconv1..dconv5).  Algorithmic MACs per point are the
        # factored count of SURVEY.md 8d; the conv1-2 recompute is not credited.
        macs_pt = {"seg_pass2_kernel": 64 * 512 + 512 * 256 + 256 * 128 + 128 * 128 + 128 * 2,
                   "seg_pass1_kernel": 3 * 64 + 64 * 64 + 64 * 64 + 64 * 128 + 128 * 1024}
        macs_pt["split_tail_kernel"] = macs_pt["seg_pass2_kernel"]
        macs_pt["split_chain_kernel[last=1024]"] = macs_pt["seg_pass1_kernel"]
        dom = max((k for k in kernel_ms if k in macs_pt), key=lambda k: kernel_ms[k], default=None)
        roofline = None
        if dom is not None:
            flops = 2.0 * macs_pt[dom] * T * N_POINTS
            ach = flops / (kernel_ms[dom] * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                        "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": _traffic(dom),
                        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                        "kernel_ms": kernel_ms[dom], "flops_per_launch": flops}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, times = cpu_oracle_objects_per_s({k: t.detach().cpu() for k, t in sd.items()})
            cpu = {"value": v, "unit": "objects/s", "cores": cores, "kind": "port",
                   "sample": "%d tracks x %d pts, %d timed passes of the fp32 oracle port on the host cores" % (
                       CPU_SAMPLE_TRACKS, N_POINTS, len(times))}
        flop_obj = spec.flops_per_object("static_one", N_POINTS)
        line = {
            "metric": "auto-labeled objects/sec", "value": value, "unit": "objects/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (split bf16 hi+lo operands, fp32 accumulate)", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "static one-box Frustum-PointNet forward + box decode, %d tracks x %d pts per GPU"
                                   " (BASELINE.json configs[2])" % (T, N_POINTS),
                       "tracks_per_gpu": T, "points": N_POINTS, "parallelism": "tracks sharded, dp%d" % world,
                       "weights": "random-init, BN randomised, seg margin calibrated",
                       "l2": "inputs %.0f MB/step > 126 MB L2, no flush needed" % (T * N_POINTS * 12 / 1e6),
                       "fg_points_per_object_median": float(fg.median().item())},
            "model_tflops": value * flop_obj / 1e12 / world, "flop_per_object": flop_obj,
            "kernel_ms": kernel_ms, "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
