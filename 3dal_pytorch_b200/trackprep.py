"""Device-side track preparation (host side of csrc/trackprep.cu): turns ragged per-track crops into the
fixed-size canonical-frame tensors the models consume, replacing the numpy work of STATICTRACK /
DYNAMICTRACK ``__getitem__`` (tools/static_model.py:529-572, tools/dynamic_model.py:419-509).

The random resample (``np.random.choice(..., replace=True)``) stays on the host as an index table, drawn
either with the reference's legacy RNG calls (parity) or with a deterministic strided rule; the gather and
all float64 transforms run on the GPU.
"""
import numpy as np
import torch

from . import _lib, ops, spec


def resample_choice_static(counts, offsets, npoints, policy="numpy_legacy"):
    """(bs, npoints) i64 absolute row indices: np.random.choice(N, npoints, replace=True) per track
    (tools/static_model.py:546) or the strided rule (j*N)//npoints."""
    out = np.full((len(counts), npoints), -1, dtype=np.int64)
    for i, (n, off) in enumerate(zip(counts, offsets)):
        n = int(n)
        if n <= 0:
            continue
        ch = np.random.choice(n, npoints, replace=True) if policy == "numpy_legacy" else (np.arange(npoints) * n) // npoints
        out[i] = off + ch
    return out


def resample_choice_dynamic(frame_counts, frame_offsets, npoints, policy="numpy_legacy"):
    """frame_counts / frame_offsets (bs, 5): rows of each window frame in the concatenated point array, count
    <= 0 for a missing or empty frame (zero points, tools/dynamic_model.py:431-439).  -> (bs, 5*npoints) i64."""
    bs, F = np.asarray(frame_counts).shape
    out = np.full((bs, F * npoints), -1, dtype=np.int64)
    for i in range(bs):
        for j in range(F):
            n = int(frame_counts[i][j])
            if n > 0:
                ch = np.random.choice(n, npoints, replace=True) if policy == "numpy_legacy" else (np.arange(npoints) * n) // npoints
                out[i, j * npoints:(j + 1) * npoints] = frame_offsets[i][j] + ch
    return out


def prep_points(src_xyz, choice, inv_pose, init_box, heading_col=6, c_out=3, time_block=0, time_center=0):
    """src_xyz (rows,3) f64 CUDA, choice (bs,n_out) i64, inv_pose (bs,4,4) f64, init_box (bs,k) f64 ->
    (bs, n_out, c_out) f32 point-major canonical-frame points."""
    ops._need_cuda(src_xyz, choice, inv_pose, init_box)
    bs, n_out = choice.shape
    for t in (src_xyz, choice, inv_pose, init_box):
        assert t.is_contiguous()
    assert src_xyz.dtype == torch.float64 and inv_pose.dtype == torch.float64 and init_box.dtype == torch.float64
    out = torch.empty((bs, n_out, c_out), device=src_xyz.device, dtype=torch.float32)
    _lib.check(_lib.lib().al3d_track_points_prep(src_xyz.data_ptr(), choice.data_ptr(), bs, n_out, inv_pose.data_ptr(),
                                                 init_box.data_ptr(), init_box.shape[1], heading_col, c_out, time_block,
                                                 time_center, out.data_ptr(), ops._stream()), "track_points_prep")
    return out


def prep_boxseq(box, inv_pose, center_step=50):
    """box (bs,steps,8) f64 global -> (bs,steps,8) f32 relative box sequence, init_box (bs,8) f64."""
    ops._need_cuda(box, inv_pose)
    bs, steps, _ = box.shape
    assert box.is_contiguous() and inv_pose.is_contiguous() and box.dtype == torch.float64
    out = torch.empty((bs, steps, 8), device=box.device, dtype=torch.float32)
    init_box = torch.empty((bs, 8), device=box.device, dtype=torch.float64)
    _lib.check(_lib.lib().al3d_boxseq_prep(box.data_ptr(), bs, steps, center_step, inv_pose.data_ptr(), out.data_ptr(),
                                           init_box.data_ptr(), ops._stream()), "boxseq_prep")
    return out, init_box
