"""Foreground-mask gather.  TEST INFRASTRUCTURE (oracle).

Restates gather_object_pts / point_cloud_masking (tools/static_model.py:23-62,
tools/dynamic_model.py:24-63).  Two index policies:

  "numpy_legacy"  the reference's behaviour: the global numpy legacy RNG is consumed
                  sequentially over the batch (np.random.choice + np.random.shuffle).
  "strided"       the deterministic device rule of the B200 build ("fast mode"): for L
                  foreground points and n_pts slots, slot j reads pos[(j * L) // n_pts] when
                  L >= n_pts and pos[j % L] otherwise; no shuffle.  It is what the reference
                  computes when np.random.choice / np.random.shuffle are replaced by that rule
                  (tests/test_oracle_vs_reference.py does exactly that monkey-patch).
"""
import numpy as np


def mask_from_logits(logits):
    """tools/static_model.py:59: strict `<`; NaN compares False."""
    return logits[..., 0] < logits[..., 1]


def choice_numpy_legacy(L, n_pts):
    """tools/static_model.py:38-45 for one object with L > 0 foreground points."""
    if L >= n_pts:
        choice = np.random.choice(L, n_pts, replace=False)
    else:
        extra = np.random.choice(L, n_pts - L, replace=True)
        choice = np.concatenate((np.arange(L), extra))
    np.random.shuffle(choice)
    return choice


def choice_strided(L, n_pts):
    j = np.arange(n_pts, dtype=np.int64)
    if L >= n_pts:
        return (j * L) // n_pts
    return j % L


def build_choice_table(counts, n_pts, policy):
    """(bs, n_pts) int64 table of positions into each object's ascending foreground list.
    Rows of empty objects are zero and unused."""
    table = np.zeros((len(counts), n_pts), dtype=np.int64)
    for i, L in enumerate(counts):
        L = int(L)
        if L > 0:
            table[i] = choice_numpy_legacy(L, n_pts) if policy == "numpy_legacy" else choice_strided(L, n_pts)
    return table


def gather_object_pts(pts, mask, n_pts, policy="numpy_legacy"):
    """pts (bs,C,n) float array, mask (bs,n) bool -> object_pts (bs,C,n_pts) f32, indices (bs,n_pts) i64.
    Objects without foreground stay all-zero (tools/static_model.py:32-37)."""
    pts = np.asarray(pts)
    mask = np.asarray(mask)
    bs, C, _ = pts.shape
    out = np.zeros((bs, C, n_pts), dtype=np.float32)
    indices = np.zeros((bs, n_pts), dtype=np.int64)
    for i in range(bs):
        pos = np.nonzero(mask[i])[0]
        if len(pos) > 0:
            ch = choice_numpy_legacy(len(pos), n_pts) if policy == "numpy_legacy" else choice_strided(len(pos), n_pts)
            indices[i] = pos[ch]
            out[i] = pts[i][:, indices[i]]
    return out, indices
