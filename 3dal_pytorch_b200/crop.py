"""Points-in-rotated-box crop on the GPU (host side of csrc/crop.cu).

Drop-in for ``box_np_ops.points_in_rbbox`` (det3d/core/bbox/box_np_ops.py:641-647) and for the per-box
crop loop of ``_create_pd_detection`` (det3d/datasets/waymo/waymo_common.py:167-171), batched over
frames.  The box -> corner -> plane arithmetic is done here on the host with numpy float32 element-wise
operations in the reference's own order (corners_nd / rotation_3d_in_axis / center_to_corner_box3d
box_np_ops.py:55-85,146-179,241-262; corner_to_surfaces_3d :650-670; surface_equ_3d_jitv2
geometry.py:351-377): it is 6x4 floats per box, and it keeps numpy's float32 sin/cos in the loop so the
device predicate sees bit-identical plane equations.  The N x B point tests, the ordered compaction,
the gather and the float64 pose transform run in libal3d.so.
"""
import os

import numpy as np
import torch

from . import _lib, ops

GRID = 64                     # BEV cells per side
AABB_PAD = 0.05               # metres; far above the ~1e-4 m rounding slack of the plane test
_SIGNS = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]],
                  dtype=np.float32) - np.float32(0.5)
_QUADS = np.array([[0, 1, 2, 3], [7, 6, 5, 4], [0, 3, 7, 4], [1, 5, 6, 2], [0, 4, 5, 1], [3, 2, 6, 7]])


def detector_to_waymo(box3d):
    """CenterPoint [x,y,z,w,l,h,r2] -> Waymo [x,y,z,l,w,h,r1 = -r2 - pi/2] (waymo_common.py:110-111)."""
    b = np.array(box3d, copy=True)
    b[:, -1] = -b[:, -1] - np.pi / 2
    return b[:, [0, 1, 2, 4, 3, 5, -1]]


def box_planes_device(boxes, device, return_inputs=False):
    """(B,7) f32 [x,y,z,l,w,h,heading] -> planes (B,6,4) f32, padded rectangles (B,6) f32, both CUDA tensors, computed by
    crop_box_setup_kernel.  Only sin / cos of the heading are evaluated here (numpy float32, like the reference:
    box_np_ops.py:163-164); tests/test_crop.py checks the result bit for bit against box_planes_host."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 7)
    n = boxes.shape[0]
    sincos = np.stack([np.sin(boxes[:, 6]), np.cos(boxes[:, 6])], 1).astype(np.float32)
    d_boxes = torch.from_numpy(boxes).to(device)
    d_sc = torch.from_numpy(sincos).to(device)
    planes = torch.empty((n, 6, 4), device=device, dtype=torch.float32)
    aabb = torch.empty((n, 6), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().al3d_crop_box_setup(d_boxes.data_ptr(), d_sc.data_ptr(), n, AABB_PAD, 1e-5, planes.data_ptr(),
                                                  aabb.data_ptr(), ops._stream()), "crop_box_setup")
    if return_inputs:
        return planes, aabb, d_boxes, d_sc
    return planes, aabb


def box_planes_host(boxes):
    """Host twin of box_planes_device (numpy, the reference's own operations): used by the tests as the checker.
    (B,7) f32 [x,y,z,l,w,h,heading] -> planes (B,6,4) f32 (inward normals, d) and padded
    rectangles (B,6) f32 [xmin,ymin,zmin,xmax,ymax,zmax]."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 7)
    f0, f1 = np.float32(0), np.float32(1)
    loc = boxes[:, None, 3:6] * _SIGNS[None]
    s, c = np.sin(boxes[:, 6])[:, None], np.cos(boxes[:, 6])[:, None]
    x, y, z = loc[..., 0], loc[..., 1], loc[..., 2]
    corners = np.stack([x * c + y * s + z * f0, x * (-s) + y * c + z * f0, x * f0 + y * f0 + z * f1], -1).astype(np.float32)
    corners += boxes[:, None, 0:3]
    q = corners[:, _QUADS, :]                                 # (B,6,4,3)
    a = q[:, :, 0] - q[:, :, 1]
    b = q[:, :, 1] - q[:, :, 2]
    nx = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    ny = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    nz = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    d = -q[:, :, 0, 0] * nx - q[:, :, 0, 1] * ny - q[:, :, 0, 2] * nz
    planes = np.stack([nx, ny, nz, d], -1).astype(np.float32)
    # min / max over the 8 corners as pairwise element-wise reductions (numpy's reduction over a middle axis is ~20x
    # slower); max|corner| = max(|min|, |max|).  Exact operations: same values as corners.min(1) / .max(1) / abs().max(1).
    cmin, cmax = corners[:, 0], corners[:, 0]
    for k in range(1, 8):
        cmin = np.minimum(cmin, corners[:, k])
        cmax = np.maximum(cmax, corners[:, k])
    pad = np.float32(AABB_PAD) + np.float32(1e-5) * np.maximum(np.abs(cmin), np.abs(cmax))
    aabb = np.concatenate([cmin - pad, cmax + pad], axis=1).astype(np.float32)
    return np.ascontiguousarray(planes), np.ascontiguousarray(aabb)


def _check_overflow(flag, what):
    code = int(flag.item())
    if code:
        names = {1: "cell list capacity", 3: "hit capacity of a chunk segment (raise hit_cap)", 4: "output capacity"}
        raise OverflowError("crop %s: %s (code %d)" % (what, names.get(code, "?"), code))


_TORCH_OF = {np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
             np.dtype(np.float64): torch.float64}


def _upload(dev, arrays):
    """One host -> device copy for a set of small numpy tables: dict name -> array (or None) in, dict name -> device tensor
    (a view of one buffer, every table 16-byte aligned) out."""
    off, total = {}, 0
    for k, a in arrays.items():
        if a is None:
            continue
        arrays[k] = a = np.ascontiguousarray(a)
        off[k] = total
        total += (a.nbytes + 15) // 16 * 16
    buf = np.zeros((max(total, 16),), np.uint8)
    for k, o in off.items():
        buf[o:o + arrays[k].nbytes] = arrays[k].reshape(-1).view(np.uint8)
    d = torch.from_numpy(buf).to(dev)
    out = {}
    for k, a in arrays.items():
        out[k] = None if a is None else d[off[k]:off[k] + a.nbytes].view(_TORCH_OF[a.dtype]).view(a.shape)
    return out


def cell_cap_bound(boxes, box_off, G):
    """Upper bound on the entries of one frame's coarse cell -> box lists (crop_grid_kernel), from the boxes alone.
    A box whose padded rectangle is at most 2 r wide spans at most floor(2 r / cell) + 2 cells per axis, at most G; the
    frame's cell size is its extent / G and the extent is at least the spread of the box centres.  r is taken per FRAME
    (its largest box): a few reductions over the boxes, loose by the ratio of the largest to the typical box."""
    boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 7)
    nb = np.diff(np.asarray(box_off))
    if boxes.shape[0] == 0:
        return 1
    with np.errstate(all="ignore"):
        start = np.asarray(box_off)[:-1][nb > 0]             # strictly increasing: reduceat is exact
        x, y = boxes[:, 0], boxes[:, 1]
        xmin, xmax = np.minimum.reduceat(x, start), np.maximum.reduceat(x, start)
        ymin, ymax = np.minimum.reduceat(y, start), np.maximum.reduceat(y, start)
        # half diagonal <= (|l| + |w|) / 2; padding = AABB_PAD + 1e-5 max|corner| in the kernel, taken generously here
        r = 0.5 * np.maximum.reduceat(np.abs(boxes[:, 3]) + np.abs(boxes[:, 4]), start).astype(np.float64)
        r = r + 2 * AABB_PAD + 1e-4 * (np.maximum(np.abs(xmin), np.abs(xmax)) + np.maximum(np.abs(ymin), np.abs(ymax)) + r)
        ex = np.maximum((xmax - xmin).astype(np.float64), 1e-3)
        ey = np.maximum((ymax - ymin).astype(np.float64), 1e-3)
        nx, ny = np.floor(2 * r * G / ex) + 2, np.floor(2 * r * G / ey) + 2
        nx = np.where(nx >= 1, np.minimum(nx, G), G)         # NaN / inf / negative: the whole axis
        ny = np.where(ny >= 1, np.minimum(ny, G), G)
        per_frame = nb[nb > 0] * nx * ny
    return int(max(per_frame.max(), 1))


class CropPlan:
    """Device-resident state of a batch crop: inputs uploaded, scratch allocated.  ``run()`` only launches the
    kernels (grid build -> hits -> scan -> fill), so it can be timed / replayed without host work."""

    def __init__(self, points, boxes, poses=None, device="cuda", want_xyz=True, hit_cap=None, chunk=None):
        lib = _lib.lib()
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.dev, self.want_xyz = dev, want_xyz
        F = self.F = len(points)
        # ---- points.  Frames that already live on the device as contiguous (N, >=3) float32 tensors of one row width are
        #      used where they lie (the kernels take a table of frame pointers); anything else is packed into (sum N, 3) on
        #      the host and uploaded in one copy.
        in_place = F > 0 and all(isinstance(p, torch.Tensor) and p.is_cuda and p.device.index == dev.index and p.dtype == torch.float32
                                 and p.dim() == 2 and p.shape[1] >= 3 and p.is_contiguous() for p in points) \
            and len({int(p.shape[1]) for p in points}) == 1
        n_pts = [int(p.shape[0]) for p in points]
        pt_off = np.concatenate([[0], np.cumsum(n_pts)]).astype(np.int64)
        self.n_points = int(pt_off[-1])
        if in_place:
            self._frames = list(points)                      # keeps the memory alive
            self.pt_stride = int(points[0].shape[1])
            ptrs = np.array([p.data_ptr() for p in points], dtype=np.int64)
            self.pts_all = None
        else:
            host = [np.asarray(p.detach().cpu() if isinstance(p, torch.Tensor) else p, dtype=np.float32) for p in points]
            host = [h.reshape(h.shape[0], -1)[:, :3] if h.shape[0] else np.zeros((0, 3), np.float32) for h in host]
            self.pts_all = torch.from_numpy(np.ascontiguousarray(np.concatenate(host, 0)) if F else np.zeros((0, 3), np.float32)).to(dev)
            self.pt_stride = 3
            ptrs = self.pts_all.data_ptr() + pt_off[:-1] * 12
        # ---- boxes
        if isinstance(boxes, np.ndarray) and boxes.ndim == 3:        # (F, B, 7): the same number of boxes in every frame, one array
            nb = [int(boxes.shape[1])] * F
            all_boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 7)
        else:
            nb = [int(np.asarray(b).reshape(-1, 7).shape[0]) for b in boxes]
            all_boxes = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float32).reshape(-1, 7) for b in boxes], 0)) if sum(nb) \
                else np.zeros((0, 7), np.float32)
        self.box_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        TB = self.TB = int(self.box_off[-1])
        sincos = np.stack([np.sin(all_boxes[:, 6]), np.cos(all_boxes[:, 6])], 1).astype(np.float32)    # numpy float32, like the reference
        self.max_boxes = max(1, max(nb) if nb else 1)
        CH = self.chunk_pts = int(chunk) if chunk else self.chunk_points(self.n_points)
        if not 1 <= CH <= lib.al3d_crop_chunk_points():
            raise ValueError("chunk=%d not in [1, %d]" % (CH, lib.al3d_crop_chunk_points()))
        # chunk table (frame, first point, points, chunk index in frame), frame-major -- vectorised: a sweep has thousands of chunks
        n_arr = np.asarray(n_pts, dtype=np.int64).reshape(-1)
        n_ch = (n_arr + CH - 1) // CH
        frame_chunk_off = np.concatenate([[0], np.cumsum(n_ch)]).astype(np.int64)
        self.n_chunks = int(frame_chunk_off[-1])
        f_of = np.repeat(np.arange(F, dtype=np.int64), n_ch)
        k_of = np.arange(self.n_chunks, dtype=np.int64) - frame_chunk_off[f_of]
        first = k_of * CH
        chunks = np.stack([f_of, first, np.minimum(CH, n_arr[f_of] - first), k_of], 1).astype(np.int32) if self.n_chunks else np.zeros((0, 4), np.int32)
        # hits per WARP segment (an eighth of a chunk, consecutive points): half of its points unless told otherwise (the
        # synthetic Waymo-shaped frames put ~10 % of the points inside a box)
        self.hit_cap = int(hit_cap or max(256, CH // 16))
        self.n_seg = 8                               # warps per chunk CTA (csrc/crop.cu kCropWarps)
        # ---- every small host table in ONE upload
        host_poses = None
        if poses is not None:
            if isinstance(poses, np.ndarray):
                host_poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(F, 4, 4)
            else:
                host_poses = np.stack([np.asarray(P, dtype=np.float64).reshape(4, 4) for P in poses]) if F else np.zeros((0, 4, 4))
        d = _upload(dev, {"ptrs": ptrs, "box_off": self.box_off, "fco": frame_chunk_off, "poses": host_poses, "chunks": chunks,
                          "boxes": all_boxes, "sincos": sincos})
        self.d_frame_ptrs, self.d_box_off, self.d_fco, self.d_chunks = d["ptrs"], d["box_off"], d["fco"], d["chunks"]
        self.d_boxes, self.d_sincos, self.d_poses = d["boxes"], d["sincos"], d["poses"]
        # ---- plane equations, padded rectangles and box-frame records on the device
        f32 = lambda *sh: torch.empty(sh, device=dev, dtype=torch.float32)
        i32 = lambda *sh: torch.empty(sh, device=dev, dtype=torch.int32)
        self.d_planes, self.d_aabb, self.d_local = f32(max(TB, 1), 6, 4)[:TB], f32(max(TB, 1), 6)[:TB], f32(max(TB, 1), 12)
        if TB:
            with torch.cuda.device(dev):
                _lib.check(lib.al3d_crop_box_setup(self.d_boxes.data_ptr(), self.d_sincos.data_ptr(), TB, AABB_PAD, 1e-5, self.d_planes.data_ptr(),
                                                   self.d_aabb.data_ptr(), ops._stream()), "crop_box_setup")
                _lib.check(lib.al3d_crop_box_local(self.d_boxes.data_ptr(), self.d_sincos.data_ptr(), TB, self.d_local.data_ptr(),
                                                   ops._stream()), "crop_box_local")
        self.meta = f32(max(F, 1), 8)
        self.occ = i32(max(F, 1), lib.al3d_crop_occ_words())
        self.cell_start = i32(max(F, 1), GRID * GRID + 1)
        self.cell4 = i32(max(F, 1), GRID * GRID, 2)          # packed cell entries (up to three box ids in 8 bytes)
        self.overflow = torch.zeros((1,), device=dev, dtype=torch.int32)
        # CSR cell lists: sized by an upper bound computed from the boxes on the host (no counting run, no read-back); the
        # overflow flag of the grid kernel stays as the safety net
        self.cell_cap = cell_cap_bound(all_boxes, self.box_off, GRID)
        self.cell_boxes = i32(max(F, 1), self.cell_cap)
        self.hits = torch.empty((max(self.n_chunks, 1), self.n_seg, self.hit_cap, lib.al3d_crop_hit_bytes() // 4), device=dev, dtype=torch.int32)
        self.n_hits = i32(max(self.n_chunks, 1), self.n_seg)
        self.cbc = i32(max(self.n_chunks, 1), self.max_boxes)
        self.box_total = i32(max(TB, 1))
        self.offsets = torch.zeros((TB + 1,), device=dev, dtype=torch.int64)
        self.capacity = None
        self.out_idx = self.out_xyz = self.out_glob = None
        # algorithmic bytes: every point read once (12 B) + plane equations; the writes are added per run
        self.read_bytes = self.n_points * 12 + TB * 96

    @staticmethod
    def chunk_points(total_points):
        """Points per work chunk (one CTA of the hits kernel): as large as the kernel takes (the per-CTA set-up and the
        ranking of its hits are amortised over the chunk) while the job still fills every SM several times over."""
        most = _lib.lib().al3d_crop_chunk_points()
        if os.environ.get("AL3D_CROP_CHUNK"):                # tuning experiments (scripts/gpu_r2_v2.sh)
            return int(min(most, max(1, int(os.environ["AL3D_CROP_CHUNK"]))))
        want = int(total_points) // 1184                     # 148 SMs x 4 resident CTAs x 2
        # measured on the 200-frame sweep (ms per sweep): 4096 0.483, 8192 0.421, 12288 0.403, 16384 0.389, 20480 0.391,
        # 24576 0.396, 32768 0.403 (fewer, longer CTAs: the last wave is emptier)
        return int(min(most, 16384, max(2048, want // 1024 * 1024)))

    def _alloc_outputs(self, capacity):
        dev = self.dev
        self.capacity = int(capacity)
        c = max(self.capacity, 1)
        self.out_idx = torch.empty((c,), device=dev, dtype=torch.int32)[: self.capacity]
        self.out_xyz = torch.empty((c, 3), device=dev, dtype=torch.float32)[: self.capacity] if self.want_xyz else None
        self.out_glob = torch.empty((c, 3), device=dev, dtype=torch.float64)[: self.capacity] if self.d_poses is not None else None

    def grid(self):
        lib, st, p = _lib.lib(), ops._stream(), (lambda t: t.data_ptr())
        _lib.check(lib.al3d_crop_build_grid(p(self.d_aabb), p(self.d_boxes), p(self.d_sincos), p(self.d_box_off), self.F, GRID,
                                            p(self.meta), p(self.cell_start), p(self.cell_boxes), self.cell_cap, p(self.cell4),
                                            self.max_boxes, p(self.occ), p(self.overflow), st), "crop_build_grid")

    def hits_pass(self):
        lib, st, p = _lib.lib(), ops._stream(), (lambda t: t.data_ptr())
        _lib.check(lib.al3d_crop_hits(p(self.d_frame_ptrs), self.pt_stride, p(self.d_planes), p(self.d_local), p(self.d_box_off), GRID, p(self.meta),
                                      p(self.cell_start), p(self.cell_boxes), self.cell_cap, p(self.cell4), p(self.occ), p(self.d_chunks), self.n_chunks,
                                      p(self.hits), self.hit_cap, p(self.n_hits), p(self.cbc), self.max_boxes, p(self.overflow), st),
                   "crop_hits")

    def scan(self):
        lib, st, p = _lib.lib(), ops._stream(), (lambda t: t.data_ptr())
        _lib.check(lib.al3d_crop_scan(p(self.d_box_off), p(self.d_fco), self.F, self.TB, p(self.cbc), self.max_boxes,
                                      p(self.box_total), p(self.offsets), st), "crop_scan")

    def count(self):
        """grid build + hits + scan (everything that does not need the output size)."""
        self.grid()
        self.hits_pass()
        self.scan()

    def fill(self):
        lib, st, p = _lib.lib(), ops._stream(), (lambda t: t.data_ptr() if t is not None else None)
        _lib.check(lib.al3d_crop_fill(p(self.d_box_off), self.F, p(self.d_chunks), self.n_chunks,
                                      p(self.hits), self.hit_cap, p(self.n_hits), p(self.cbc), self.max_boxes, p(self.offsets),
                                      p(self.d_poses), self.capacity, p(self.out_idx), p(self.out_xyz), p(self.out_glob),
                                      p(self.overflow), st), "crop_fill")

    def run(self, capacity=None):
        """One full crop.  Without ``capacity`` (and without a previous run) the total is read back once to size the
        outputs; afterwards the same buffers are reused and nothing synchronises with the host."""
        self.count()
        if capacity is not None and self.capacity != capacity:
            self._alloc_outputs(capacity)
        if self.capacity is None:
            if int(self.overflow.item()) == 1:               # cell lists larger than the host bound (never seen): size them exactly
                self.cell_cap = max(int(self.cell_start[:, GRID * GRID].max().item()), 1)
                self.cell_boxes = torch.empty((max(self.F, 1), self.cell_cap), device=self.dev, dtype=torch.int32)
                self.overflow.zero_()
                self.count()
            _check_overflow(self.overflow, "hits")
            self._alloc_outputs(int(self.offsets[-1].item()))
        self.fill()
        return self.result()

    def result(self):
        return {"indices": self.out_idx, "offsets": self.offsets, "box_off": self.box_off, "xyz": self.out_xyz,
                "xyz_global": self.out_glob, "overflow": self.overflow, "read_bytes": self.read_bytes}


def crop_frames(points, boxes, poses=None, device="cuda", want_xyz=True, hit_cap=None, capacity=None, chunk=None):
    """points: list of (N_f, >=3) f32 arrays / CUDA tensors; boxes: list of (B_f, 7) f32 Waymo-convention
    arrays; poses: optional list of 4x4 f64 vehicle->global matrices.

    Returns dict(indices (T,) i32 CUDA -- point index within its frame, ascending per box;
                 offsets (sum B_f + 1,) i64 CUDA -- box b of frame f owns indices[offsets[k]:offsets[k+1]], k = box_off[f]+b;
                 box_off (F+1,) i64 host; xyz (T,3) f32; xyz_global (T,3) f64 when poses are given).
    With ``capacity`` set no host synchronisation happens (the output is over-allocated); otherwise the total
    count is read back once to size the outputs exactly."""
    return CropPlan(points, boxes, poses, device=device, want_xyz=want_xyz, hit_cap=hit_cap, chunk=chunk).run(capacity)


def points_in_rbbox(points, rbbox, z_axis=2, origin=(0.5, 0.5, 0.5), device="cuda"):
    """Same signature and (N, B) bool numpy result as box_np_ops.points_in_rbbox (float32 points)."""
    if z_axis != 2 or tuple(origin) != (0.5, 0.5, 0.5):
        raise NotImplementedError("only the lidar convention used by the 3DAL path (z_axis=2, origin 0.5) is built")
    pts = np.asarray(points)
    if pts.dtype != np.float32:
        raise TypeError("GPU crop takes float32 points (the reference's float64 specialisation is label-only)")
    rb = np.asarray(rbbox, dtype=np.float32).reshape(-1, 7)
    N, B = pts.shape[0], rb.shape[0]
    if N == 0 or B == 0:
        return np.zeros((N, B), dtype=bool)
    try:
        res = crop_frames([pts], [rb], device=device, want_xyz=False)
    except OverflowError:
        res = crop_frames([pts], [rb], device=device, want_xyz=False, hit_cap=CropPlan.chunk_points(N) // 8 * B)
    _check_overflow(res["overflow"], "fill")
    mask = torch.zeros((N, B), device=res["indices"].device, dtype=torch.uint8)
    _lib.check(_lib.lib().al3d_crop_dense_mask(res["indices"].data_ptr(), res["offsets"].data_ptr(), B, mask.data_ptr(),
                                               ops._stream()), "crop_dense_mask")
    return mask.cpu().numpy().astype(bool)
