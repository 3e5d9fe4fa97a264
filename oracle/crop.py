"""Points-in-rotated-box crop.  TEST INFRASTRUCTURE (oracle).

Restates the reference chain behind ``box_np_ops.points_in_rbbox`` (call site
det3d/datasets/waymo/waymo_common.py:168):

  box -> 8 corners            center_to_corner_box3d / corners_nd / rotation_3d_in_axis
                              det3d/core/bbox/box_np_ops.py:241-262, 55-85, 146-179
  corners -> 6 inward quads   corner_to_surfaces_3d      box_np_ops.py:650-670
  quads -> planes (n, d)      surface_equ_3d_jitv2       geometry.py:351-377
  point test                  _points_in_convex_polygon_3d_jit   geometry.py:241-276
  crop materialisation        waymo_common.py:169-171 (gather in ascending index order, then the
                              homogeneous float64 pose transform)

All box arithmetic is float32, un-fused, in the reference's association order.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# corner order of corners_nd for ndim=3 after the [0,1,3,2,4,5,7,6] re-order, origin 0.5
_CORNER_SIGNS = np.array(
    [[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]], dtype=np.float32
) - np.float32(0.5)
# corner ids of the six quads (box_np_ops.py:660-667)
_QUADS = np.array([[0, 1, 2, 3], [7, 6, 5, 4], [0, 3, 7, 4], [1, 5, 6, 2], [0, 4, 5, 1], [3, 2, 6, 7]])


def build(force=False):
    """Compile crop_ref.c into oracle/_build/libal3d_oracle.so (gcc)."""
    so = os.path.join(_HERE, "_build", "libal3d_oracle.so")
    src = os.path.join(_HERE, "crop_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        lib.al3d_ref_planes.argtypes = [vp, i64, vp]
        lib.al3d_ref_inside.argtypes = [vp, i64, i64, vp, i64, vp]
        lib.al3d_ref_inside_f64.argtypes = [vp, i64, i64, vp, i64, vp]
        _LIB = lib
    return _LIB


def box_corners(boxes):
    """(B,7) f32 [x,y,z,l,w,h,heading] -> (B,8,3) f32 corners (origin 0.5, rotation about z)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    local = boxes[:, None, 3:6] * _CORNER_SIGNS[None]                      # (B,8,3) f32
    s = np.sin(boxes[:, 6])[:, None]
    c = np.cos(boxes[:, 6])[:, None]
    x, y, z = local[..., 0], local[..., 1], local[..., 2]
    # einsum("aij,jka->aik") with rot_mat_T = [[c,-s,0],[s,c,0],[0,0,1]]: a plain 3-term f32 sum
    rx = x * c + y * s + z * np.float32(0)
    ry = x * (-s) + y * c + z * np.float32(0)
    rz = x * np.float32(0) + y * np.float32(0) + z * np.float32(1)
    out = np.stack([rx, ry, rz], axis=-1).astype(np.float32)
    out += boxes[:, None, 0:3]
    return out


def box_planes(boxes):
    """(B,7) f32 -> (B,6,4) f32 inward plane equations (nx,ny,nz,d) via the C restatement."""
    surfaces = np.ascontiguousarray(box_corners(boxes)[:, _QUADS, :], dtype=np.float32)  # (B,6,4,3)
    planes = np.empty((surfaces.shape[0], 6, 4), dtype=np.float32)
    _lib().al3d_ref_planes(surfaces.ctypes.data, surfaces.shape[0], planes.ctypes.data)
    return planes


def points_in_boxes(points, boxes):
    """points (N,>=3) f32 or f64, boxes (B,7) f32 -> (N,B) bool, == box_np_ops.points_in_rbbox."""
    planes = box_planes(boxes)
    pts = np.ascontiguousarray(points)
    out = np.empty((pts.shape[0], planes.shape[0]), dtype=np.uint8)
    if pts.dtype == np.float64:
        _lib().al3d_ref_inside_f64(pts.ctypes.data, pts.shape[0], pts.shape[1], planes.ctypes.data, planes.shape[0],
                                   out.ctypes.data)
    else:
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        _lib().al3d_ref_inside(pts.ctypes.data, pts.shape[0], pts.shape[1], planes.ctypes.data, planes.shape[0],
                               out.ctypes.data)
    return out.astype(bool)


def detector_to_waymo(box3d):
    """waymo_common.py:110-111: heading = -r2 - pi/2 (in the box dtype), swap the two footprint dims."""
    b = np.array(box3d, copy=True)
    b[:, -1] = -b[:, -1] - np.pi / 2
    return b[:, [0, 1, 2, 4, 3, 5, -1]]


def crop_frame(points, waymo_boxes, pose):
    """Per box: ascending indices of the points inside and their float64 global-frame xyz
    (waymo_common.py:168-171).  Returns (list of int64 index arrays, list of (k,3) f64 arrays)."""
    inside = points_in_boxes(points, waymo_boxes)
    idx, xyz = [], []
    for b in range(waymo_boxes.shape[0]):
        sel = np.nonzero(inside[:, b])[0]
        o = points[sel].T
        o = pose @ np.concatenate([o, np.ones((1, o.shape[1]))], axis=0)
        idx.append(sel)
        xyz.append(o[:3, :].T)
    return idx, xyz
