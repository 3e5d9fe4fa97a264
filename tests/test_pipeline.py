"""Host-facing pipeline (pipeline.StaticAutoLabeler): chunk schedule (CPU) and label_host == label_device (GPU)."""
import importlib

import pytest
import torch

from helpers import synth

pipeline = importlib.import_module("3dal_pytorch_b200.pipeline")
sm = importlib.import_module("3dal_pytorch_b200.static_model")


@pytest.mark.parametrize("T,chunk,first", [(8192, 2048, None), (100, 64, 16), (5, 64, None), (64, 64, 64), (1000, 256, 1)])
def test_chunk_schedule_covers_all_tracks_once(T, chunk, first):
    lab = pipeline.StaticAutoLabeler(None, chunk_tracks=chunk, first_chunk_tracks=first)
    bounds = lab._schedule(T)
    assert bounds[0][0] == 0 and bounds[-1][1] == T
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))            # contiguous, no overlap
    assert all(0 < t1 - t0 <= chunk for t0, t1 in bounds)                   # staging buffers hold `chunk` tracks
    assert bounds[0][1] - bounds[0][0] <= max(1, first or chunk // 4)       # short first chunk: its copy is not hidden


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "mixed", "bf16"])
def test_label_host_equals_label_device(prec):
    dev = "cuda:0"
    T = 150                                                                  # not a multiple of the chunk sizes
    model = sm.StaticModelOneBoxEst().to(dev).eval()
    model.load_state_dict(synth.random_state_dict("static_one", seed=4))
    model.precision = prec
    data = synth.static_tracks(T, seed=11)                                   # point-major (T, n, 3) host arrays
    pts_host = torch.from_numpy(data["pts_pm"]).float().pin_memory()
    box_host = torch.from_numpy(data["init_box"]).float().pin_memory()
    lab = pipeline.StaticAutoLabeler(model, chunk_tracks=64, first_chunk_tracks=16)
    out_host = lab.label_host(pts_host, box_host)
    ref = lab.label_device(pts_host.to(dev).transpose(2, 1), box_host.to(dev))
    torch.cuda.synchronize()
    assert out_host.shape == (T, 7)
    # every track is an independent unit: chunking must not change a single bit
    assert torch.equal(out_host, ref.cpu())
    # a second call reuses the staging buffers (double buffering across calls must not leak state)
    assert torch.equal(lab.label_host(pts_host, box_host), out_host)
