"""Crop: host plane setup (CPU) and the CUDA crop against the oracle / golden frame (GPU)."""
import importlib
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, synth
from oracle import crop as ocrop

crop = importlib.import_module("3dal_pytorch_b200.crop")


def _random_boxes(rng, nb, spread=20.0):
    return np.concatenate([rng.normal(0, spread, (nb, 2)), rng.normal(0.5, 0.5, (nb, 1)), rng.uniform(0.4, 11, (nb, 3)),
                           rng.uniform(-7, 7, (nb, 1))], 1).astype(np.float32)


def test_host_planes_bitwise_equal_to_oracle():
    rng = np.random.default_rng(0)
    boxes = _random_boxes(rng, 2000)
    planes, aabb = crop.box_planes_host(boxes)
    assert np.array_equal(planes, ocrop.box_planes(boxes))
    corners = ocrop.box_corners(boxes)
    assert np.all(aabb[:, :3] < corners.min(1)) and np.all(aabb[:, 3:] > corners.max(1))
    assert np.array_equal(crop.detector_to_waymo(boxes), ocrop.detector_to_waymo(boxes))


def test_cell_list_bound_covers_the_grid_kernel_count():
    """crop.cell_cap_bound sizes the coarse cell -> box lists without a counting run on the device: it must never be below
    what crop_grid_kernel registers (replayed here in numpy float32: padded rectangles -> frame extent -> cell ranges)."""
    rng = np.random.default_rng(5)
    G = crop.GRID
    frames = [_random_boxes(rng, 200, spread=40.0), _random_boxes(rng, 3, spread=0.3), _random_boxes(rng, 1),
              np.tile(_random_boxes(rng, 1), (10, 1)), _random_boxes(rng, 60, spread=2000.0), np.zeros((0, 7), np.float32),
              crop.detector_to_waymo(synth.lidar_frames(1, n_points=10, seed=2)[0]["det_boxes"])]
    box_off = np.concatenate([[0], np.cumsum([len(b) for b in frames])])
    bound = crop.cell_cap_bound(np.concatenate(frames, 0), box_off, G)
    worst = 0
    for b in frames:
        if len(b) == 0:
            continue
        _, aabb = crop.box_planes_host(b)
        x0, y0 = aabb[:, 0].min(), aabb[:, 1].min()
        ex = max(np.float32(aabb[:, 3].max() - x0), np.float32(1e-3)); ey = max(np.float32(aabb[:, 4].max() - y0), np.float32(1e-3))
        ix, iy = np.float32(G) / ex * np.float32(0.999), np.float32(G) / ey * np.float32(0.999)
        cell = lambda v, v0, inv: np.clip(np.floor((v - v0) * inv), 0, G - 1)
        nx = cell(aabb[:, 3], x0, ix) - cell(aabb[:, 0], x0, ix) + 1
        ny = cell(aabb[:, 4], y0, iy) - cell(aabb[:, 1], y0, iy) + 1
        worst = max(worst, int((nx * ny).sum()))
    assert bound >= worst, (bound, worst)


def _check_against_oracle(points, boxes, poses, res):
    off = res["offsets"].cpu().numpy()
    idx = res["indices"].cpu().numpy()
    xyz = res["xyz"].cpu().numpy()
    glob = res["xyz_global"].cpu().numpy() if res["xyz_global"] is not None else None
    box_off = res["box_off"]
    for f in range(len(points)):
        ridx, rxyz = ocrop.crop_frame(np.asarray(points[f], dtype=np.float32), np.asarray(boxes[f], np.float32).reshape(-1, 7),
                                      poses[f] if poses is not None else np.eye(4))
        for b in range(len(ridx)):
            k = box_off[f] + b
            got = idx[off[k]:off[k + 1]]
            assert np.array_equal(got, ridx[b]), (f, b, len(got), len(ridx[b]))           # bit-exact, ascending
            assert np.array_equal(xyz[off[k]:off[k + 1]], np.asarray(points[f])[ridx[b], :3], equal_nan=True)
            if glob is not None and len(got):
                ref = rxyz[b]
                assert np.allclose(glob[off[k]:off[k + 1]], ref, rtol=1e-12, atol=1e-9, equal_nan=True)   # f64 transform: FP, not bitwise
    assert off[-1] == len(idx)


@pytest.mark.gpu
def test_crop_golden_frame():
    z = np.load(os.path.join(GOLDEN, "crop_frame.npz"))
    box = crop.detector_to_waymo(z["det_boxes"])
    assert np.array_equal(box, z["waymo_boxes"])
    res = crop.crop_frames([z["points"]], [box], [z["pose"]])
    off = res["offsets"].cpu().numpy()
    assert np.array_equal(off, z["offsets"])
    assert np.array_equal(res["indices"].cpu().numpy(), z["indices"])
    assert np.allclose(res["xyz_global"].cpu().numpy(), z["xyz_global"], rtol=1e-12, atol=1e-9)
    m = crop.points_in_rbbox(z["points"], box)
    ref = ocrop.points_in_boxes(z["points"], box)
    assert m.dtype == np.bool_ and np.array_equal(m, ref)


@pytest.mark.gpu
def test_crop_ragged_batch_and_edge_cases():
    rng = np.random.default_rng(3)
    frames = synth.lidar_frames(3, n_points=30000, n_boxes=60, seed=5)
    points = [f["points"] for f in frames] + [np.zeros((0, 3), np.float32), rng.normal(0, 5, (777, 3)).astype(np.float32)]
    boxes = [crop.detector_to_waymo(f["det_boxes"]) for f in frames] + [_random_boxes(rng, 5), np.zeros((0, 7), np.float32)]
    poses = [f["pose"] for f in frames] + [np.eye(4), np.eye(4)]
    # heavy overlap (several boxes over the same points), a NaN point and a point exactly on a face
    dense = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    dense[17] = np.nan
    dense[18] = [1.0, 0.0, 0.0]
    ob = np.array([[0, 0, 0, 2, 2, 2, 0.0], [0.1, 0, 0, 2, 2, 2, 0.3], [0, 0.1, 0, 3, 1, 2, -0.4], [5, 5, 0, 1, 1, 1, 0.0]], np.float32)
    points.append(dense); boxes.append(ob); poses.append(frames[0]["pose"])
    res = crop.crop_frames(points, boxes, poses, hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle(points, boxes, poses, res)
    # the chunk size is a scheduling choice: the largest chunks (what a 200-frame sweep runs with), odd ones and tiny ones
    # give the same lists
    for ch in (16384, 5000, 130):
        alt = crop.crop_frames(points, boxes, poses, hit_cap=16384, chunk=ch)
        assert int(alt["overflow"].item()) == 0
        for k in ("indices", "offsets", "xyz", "xyz_global"):
            assert np.array_equal(alt[k].cpu().numpy(), res[k].cpu().numpy(), equal_nan=k.startswith("xyz")), (ch, k)
    # NaN point is "inside" every box, exactly like the reference predicate
    off = res["offsets"].cpu().numpy(); idx = res["indices"].cpu().numpy()
    k = res["box_off"][5] + 3
    assert 17 in idx[off[k]:off[k + 1]]


@pytest.mark.gpu
def test_crop_overflow_is_reported():
    rng = np.random.default_rng(1)
    pts = rng.uniform(-0.5, 0.5, (4096, 3)).astype(np.float32)
    boxes = np.tile(np.array([[0, 0, 0, 4, 4, 4, 0.0]], np.float32), (10, 1))     # every point inside 10 boxes
    with pytest.raises(OverflowError):
        crop.crop_frames([pts], [boxes], hit_cap=1024)                           # 256 points x 10 boxes per warp segment
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=8192)             # enough room: no limit on boxes per point
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)


@pytest.mark.gpu
def test_crop_waymo_sized_frame_property():
    """Full-size frame (180k points, 200 boxes): index lists are ascending, disjoint from outside points and
    consistent with a dense recomputation on a random subset of boxes."""
    fr = synth.lidar_frames(1, seed=9)[0]
    box = crop.detector_to_waymo(fr["det_boxes"])
    res = crop.crop_frames([fr["points"]], [box], [fr["pose"]])
    off = res["offsets"].cpu().numpy(); idx = res["indices"].cpu().numpy()
    for b in range(box.shape[0]):
        seg = idx[off[b]:off[b + 1]]
        assert np.all(np.diff(seg) > 0)
    sel = np.random.default_rng(0).choice(box.shape[0], 12, replace=False)
    ref = ocrop.points_in_boxes(fr["points"], box[sel])
    for j, b in enumerate(sel):
        assert np.array_equal(idx[off[b]:off[b + 1]], np.nonzero(ref[:, j])[0])


@pytest.mark.gpu
def test_device_plane_setup_is_bit_identical_to_numpy():
    """crop_box_setup_kernel (explicit round-to-nearest float32 operations in numpy's order, sin / cos from the host)
    against the host twin, which test_host_planes_bitwise_equal_to_oracle pins to the reference arithmetic."""
    rng = np.random.default_rng(11)
    boxes = np.concatenate([_random_boxes(rng, 5000), _random_boxes(rng, 500, spread=2000.0)], 0)
    boxes[7, 6] = 0.0
    boxes[8, 3:6] = [1e-3, 40.0, 0.25]
    planes, aabb = crop.box_planes_device(boxes, "cuda:0")
    hp, ha = crop.box_planes_host(boxes)
    assert np.array_equal(planes.cpu().numpy(), hp)
    assert np.array_equal(aabb.cpu().numpy(), ha)


@pytest.mark.gpu
def test_crop_few_large_overlapping_boxes():
    """near-duplicate detections cover (almost) every BEV cell each; the cell lists are sized from a counting
    run, so this neither overflows nor truncates."""
    rng = np.random.default_rng(2)
    pts = rng.uniform(-3, 3, (20000, 3)).astype(np.float32)
    boxes = np.array([[0, 0, 0, 4.0, 2.0, 1.5, 0.3], [0.02, 0.01, 0, 4.0, 2.0, 1.5, 0.31], [0.5, 0.2, 0, 4.2, 1.9, 1.5, 0.25]], np.float32)
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)
    m = crop.points_in_rbbox(pts, boxes)
    assert np.array_equal(m, ocrop.points_in_boxes(pts, boxes))


@pytest.mark.gpu
def test_crop_many_candidate_rectangles_but_few_hits():
    """A point may sit in the padded rectangles of more than 8 boxes as long as it is inside at most 8 of them: twelve
    thin, rotated boxes fanned around the origin overlap as rectangles far more than as boxes."""
    rng = np.random.default_rng(4)
    ang = np.linspace(0, np.pi, 12, endpoint=False).astype(np.float32)
    boxes = np.stack([np.zeros(12), np.zeros(12), np.zeros(12), np.full(12, 8.0), np.full(12, 0.2), np.full(12, 1.0), ang], 1).astype(np.float32)
    pts = rng.uniform(-4, 4, (30000, 3)).astype(np.float32)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > 1.0]                      # keep clear of the hub where all twelve overlap
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)


@pytest.mark.gpu
def test_crop_points_within_the_rounding_margin_of_faces():
    """The hits kernel classifies a (point, box) pair in the box's own frame and only evaluates the float32 plane
    equations inside a rounding margin around the faces.  Points placed at 1e-7 .. 1e-1 m on both sides of the faces of
    rotated boxes far from the origin (where float32 rounding is largest) must land exactly where the reference puts them."""
    rng = np.random.default_rng(12)
    nb = 120
    boxes = np.concatenate([rng.uniform(-300, 300, (nb, 2)), rng.normal(0.5, 0.5, (nb, 1)),
                            rng.uniform(0.3, 12, (nb, 3)), rng.uniform(-7, 7, (nb, 1))], 1).astype(np.float32)
    boxes[:6, 3:6] = [[8.0, 0.2, 1.0], [0.2, 8.0, 1.0], [20.0, 2.5, 0.3], [0.5, 0.5, 0.5], [30.0, 0.15, 4.0], [1.0, 1.0, 1.0]]
    pts = []
    for b in boxes.astype(np.float64):
        n = 400
        u = rng.uniform(-0.5, 0.5, (n, 3)) * b[3:6]
        ax = rng.integers(0, 3, n)
        side = rng.choice([-1.0, 1.0], n)
        eps = side * 10.0 ** rng.uniform(-7, -1, n) * rng.choice([-1.0, 1.0], n)
        u[np.arange(n), ax] = side * 0.5 * b[3 + ax] + eps
        c, s = np.cos(b[6]), np.sin(b[6])
        w = np.stack([u[:, 0] * c + u[:, 1] * s, -u[:, 0] * s + u[:, 1] * c, u[:, 2]], 1) + b[:3]
        pts.append(w)
    pts = np.concatenate(pts, 0)[rng.permutation(nb * 400)].astype(np.float32)
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)
    ref = ocrop.points_in_boxes(pts, boxes)
    assert 0.2 < ref.any(1).mean() < 0.8                                   # the sample really straddles the faces


@pytest.mark.gpu
@pytest.mark.parametrize("offset", [1.0, 60.0, 800.0, 20000.0])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_crop_margin_classification_fuzz(offset, seed):
    """Same property over box sizes from 5 cm to 40 m, aspect ratios up to 800 : 1 and scene offsets up to 20 km (where the
    float32 grid of the coordinates is ~2 mm and most pairs fall inside the margin or the box is declared irregular):
    whatever path a pair takes -- classified in the box frame, or evaluated exactly -- the lists equal the oracle's."""
    rng = np.random.default_rng(100 + seed)
    nb = 64
    dims = 10.0 ** rng.uniform(np.log10(0.05), np.log10(40.0), (nb, 3))
    boxes = np.concatenate([rng.uniform(-1, 1, (nb, 2)) * offset + offset, rng.normal(0.5, 0.5, (nb, 1)), dims,
                            rng.uniform(-7, 7, (nb, 1))], 1).astype(np.float32)
    pts = []
    for b in boxes.astype(np.float64):
        n = 300
        u = rng.uniform(-0.6, 0.6, (n, 3)) * b[3:6]                          # some inside, some outside
        k = rng.random(n) < 0.6                                              # most of them pushed next to a face
        ax = rng.integers(0, 3, n)
        eps = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-8, 0, n)
        u[np.arange(n)[k], ax[k]] = (rng.choice([-1.0, 1.0], n) * 0.5 * b[3 + ax] + eps)[k]
        c, s_ = np.cos(b[6]), np.sin(b[6])
        pts.append(np.stack([u[:, 0] * c + u[:, 1] * s_, -u[:, 0] * s_ + u[:, 1] * c, u[:, 2]], 1) + b[:3])
    pts = np.concatenate(pts, 0)[rng.permutation(nb * 300)].astype(np.float32)
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=32768)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)


@pytest.mark.gpu
def test_crop_infinite_points_and_irregular_boxes():
    """inf / NaN / huge coordinates and boxes with zero or negative dimensions take the exact predicate -- whatever the reference's float32 arithmetic says (inf * 0 = NaN never rejects) is the answer."""
    rng = np.random.default_rng(13)
    pts = rng.uniform(-6, 6, (9000, 3)).astype(np.float32)
    pts[:, 2] = rng.uniform(-1.5, 1.5, 9000)
    inf = np.float32(np.inf)
    pts[5] = [inf, inf, 0.0]
    pts[6] = [inf, 0.0, 0.0]
    pts[7] = [0.0, -inf, 0.2]
    pts[8] = [np.nan, 0.0, 0.0]
    pts[9] = [1e30, 1e30, 0.0]
    pts[10] = [0.0, 0.0, inf]
    pts[4000] = [-inf, inf, -inf]
    boxes = np.array([[0, 0, 0, 4, 2, 2, 0.0],            # axis aligned: inf * 0 products
                      [1, 1, 0, 3, 3, 2, np.pi / 2],
                      [0.5, -1, 0, 4, 2, 2, 0.7],
                      [2, 2, 0, 0.0, 2, 2, 0.3],          # zero length
                      [-2, 1, 0, -3, 2, 2, 0.1],          # negative length
                      [-2, -2, 0, -3, -2, 2, 0.4],        # two negative dimensions
                      [0, 3, 0, 1e-4, 1e-4, 1e-4, 0.2],   # margin not small against the box
                      [3, -3, 0, 2, 2, 2, 0.0]], np.float32)
    res = crop.crop_frames([pts], [boxes], [np.eye(4)], hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle([pts], [boxes], [np.eye(4)], res)
    m = crop.points_in_rbbox(pts, boxes)
    assert np.array_equal(m, ocrop.points_in_boxes(pts, boxes))


@pytest.mark.gpu
def test_crop_frames_with_many_boxes_use_the_global_tables():
    """Up to 256 boxes per frame the hits kernel keeps the frame's box records and packed cell table in shared memory;
    beyond that it reads them from global memory (8-byte cell entries).  Both must give the reference's lists -- also
    in one batch, where the largest frame decides."""
    rng = np.random.default_rng(21)
    frames = synth.lidar_frames(2, n_points=40000, n_boxes=600, seed=6)
    points = [f["points"] for f in frames] + [rng.normal(0, 8, (5000, 3)).astype(np.float32)]
    boxes = [crop.detector_to_waymo(f["det_boxes"]) for f in frames] + [_random_boxes(rng, 7, spread=6.0)]
    poses = [f["pose"] for f in frames] + [np.eye(4)]
    res = crop.crop_frames(points, boxes, poses, hit_cap=16384)
    assert int(res["overflow"].item()) == 0
    _check_against_oracle(points, boxes, poses, res)
    # frames that already live on the device are used where they lie, whatever their row width (x y z intensity elongation)
    wide = [torch.from_numpy(np.concatenate([p, rng.random((len(p), 2)).astype(np.float32)], 1)).cuda() for p in points]
    res5 = crop.crop_frames(wide, boxes, poses, hit_cap=16384)
    plain = [torch.from_numpy(p).cuda() for p in points]
    res3 = crop.crop_frames(plain, boxes, poses, hit_cap=16384)
    for k in ("indices", "offsets", "xyz", "xyz_global"):
        assert torch.equal(res5[k], res[k]) and torch.equal(res3[k], res[k]), k
