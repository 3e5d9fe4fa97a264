"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, box gather, global RNG replay."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sharding = importlib.import_module("3dal_pytorch_b200.sharding")
engine = importlib.import_module("3dal_pytorch_b200.engine")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(total, rank, world)
        g = torch.Generator().manual_seed(0)
        all_boxes = torch.randn(total, 7, generator=g)                     # what one process would produce
        counts = torch.randint(0, 900, (total,), generator=g)
        got = sharding.gather_boxes(all_boxes[lo:hi].clone(), total)
        assert torch.equal(got, all_boxes), "gathered boxes differ from the single-process result"
        np.random.seed(123)
        mine = sharding.global_choice_tables(counts[lo:hi], total, 512)
        np.random.seed(123)
        ref = engine.choice_table_numpy_legacy(counts.numpy(), 512)
        assert np.array_equal(mine, ref[lo:hi]), "sharded RNG replay differs from the single-process stream"
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_two_rank_gather_and_rng_replay(tmp_path, total):
    mp.spawn(_worker, args=(2, _free_port(), total, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_ranges_partition_exactly():
    for total in (0, 1, 7, 8, 8191, 8192):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def _grad_worker(rank, world, port, out_dir):
    """Training configuration (BASELINE.json configs[4]): every rank fills its flat gradient bucket, ONE all-reduce sums
    them, the optimiser applies 1 / world -- the result equals the gradient of the concatenated batch's mean loss."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        train = importlib.import_module("3dal_pytorch_b200.train")
        sm = importlib.import_module("3dal_pytorch_b200.static_model")
        torch.manual_seed(0)
        model = sm.StaticModelOneBoxEst()                       # parameter container only: no forward on the CPU
        bucket = train.GradBucket(model)
        assert bucket.flat.numel() == sum(p.numel() for p in model.parameters()) == 1481513
        g = torch.Generator().manual_seed(100)
        per_rank = [torch.randn(bucket.flat.numel(), generator=g) for _ in range(world)]
        bucket.flat.copy_(per_rank[rank])
        bucket.attach()
        w = dict(model.named_parameters())["box_est.fc3.weight"]
        assert w.grad.data_ptr() == bucket.view(w).data_ptr()  # p.grad is a view of the bucket
        scale = train.allreduce_gradients(bucket)
        assert scale == 1.0 / world
        assert torch.allclose(bucket.flat * scale, sum(per_rank) / world)
        assert torch.equal(w.grad, bucket.view(w))
        open(os.path.join(out_dir, "g%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_flat_bucket_gradient_allreduce(tmp_path):
    mp.spawn(_grad_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "g0") and os.path.exists(tmp_path / "g1")
