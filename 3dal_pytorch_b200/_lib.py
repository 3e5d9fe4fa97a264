"""ctypes binding of libal3d.so (the C ABI declared in include/al3d.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import of any model
module fails with an explicit error telling how to build it (``python -c "import
__graft_entry__ as g; g.build()"``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AL3D_LIB selects another build of the same library inside the package directory (the protocol stress tests load
# libal3d_stress.so, the build with delay-injection hooks, in a subprocess)
LIB_PATH = os.path.join(_HERE, os.path.basename(os.environ.get("AL3D_LIB", "libal3d.so")))

_vp, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argtypes ; every function returns int (0 = ok) unless listed in _RESTYPES
_SIGNATURES = {
    "al3d_abi_version": [],
    "al3d_last_error": [],
    "al3d_device_supports_tcgen05": [],
    "al3d_pointwise_first_f32": [_vp, _i64, _i64, _i64, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp],
    "al3d_linear_f32": [_vp, _i64, _i64, _i, _vp, _i64, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp, _vp],
    "al3d_mask_compact": [_vp, _vp, _i, _i, _vp, _vp, _vp],
    "al3d_gather_fg": [_vp, _i64, _i64, _i64, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "al3d_parse_heads": [_vp, _i, _vp, _i64] + [_vp] * 8 + [_vp],
    "al3d_decode_boxes": [_vp] * 6 + [_i64, _i, _vp, _vp, _vp],
    "al3d_twostage_retransform": [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "al3d_fc_chain": [_vp, _i, _vp],
    "al3d_crop_box_setup": [_vp, _vp, _i64, _f, _f, _vp, _vp, _vp],
    "al3d_crop_chunk_points": [],
    "al3d_crop_occ_words": [],
    "al3d_crop_box_local": [_vp, _vp, _i64, _vp, _vp],
    "al3d_crop_build_grid": [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp],
    "al3d_crop_hits": [_vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp],
    "al3d_crop_scan": [_vp, _vp, _i, _i64, _vp, _i, _vp, _vp, _vp],
    "al3d_crop_fill": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "al3d_crop_hit_bytes": [],
    "al3d_crop_dense_mask": [_vp, _vp, _i, _vp, _vp],
    "al3d_track_points_prep": [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "al3d_boxseq_prep": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "al3d_track_regroup": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "al3d_motion_features": [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp],
    "al3d_track_labels": [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "al3d_box_writeback": [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "al3d_match_iou3d": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "al3d_loss_forward": [_vp, _vp, _i64] + [_vp] * 10 + [_i, _vp, _i, _vp, _vp],
    "al3d_train_ws_floats": [_i64, _i],
    "al3d_wgrad_ws_floats": [_i64, _i, _i],
    "al3d_bn_train_forward": [_vp, _i64, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "al3d_bn_train_backward": [_vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "al3d_group_colsum": [_vp, _i64, _i, _i64, _vp, _vp, _vp],
    "al3d_group_max_forward": [_vp, _i64, _i64, _i, _vp, _vp, _vp],
    "al3d_group_max_backward": [_vp, _vp, _i64, _i64, _i, _vp, _vp],
    "al3d_wgrad_f32": [_vp, _i64, _vp, _i64, _i64, _i, _i, _vp, _vp, _i64, _i, _vp],
    "al3d_gemm_split_ws_bytes": [_i, _i, _i],
    "al3d_gemm_split_nt": [_vp, _i64, _i, _i, _vp, _i64, _i, _vp, _vp, _i, _i, _i, _vp, _i64, _i, _vp, _vp],
    "al3d_gemm_split_tn_ws_bytes": [_i64, _i, _i, _i],
    "al3d_gemm_split_tn": [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _vp, _vp, _i64, _i, _vp],
    "al3d_loss_backward": [_vp, _vp, _i64] + [_vp] * 10 + [_i, _vp, _vp, _vp, _vp],
    "al3d_seg_correct": [_vp, _vp, _i64, _vp, _vp],
    "al3d_adam_step": [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _vp],
    "al3d_chain_maxpool_bf16": [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp],
    "al3d_seg_pass1_bf16": [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp],
    "al3d_seg_pass2_bf16": [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "al3d_chain_maxpool_bf16x3": [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp],
    "al3d_seg_pass2_bf16x3": [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp, _vp, _vp],
    "al3d_umma_selftest": [_vp, _vp, _i, _i, _vp, _i, _vp],
    "al3d_umma_selftest_ts": [_vp, _vp, _i, _i, _vp, _vp],
    "al3d_umma_selftest_pair": [_vp, _vp, _i, _i, _vp, _vp],
    "al3d_umma_selftest_pair_ss": [_vp, _vp, _i, _vp, _vp],
    "al3d_mma_microbench": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "al3d_tc_abort_code": [_vp],
    "al3d_tc_status_word_host": [_vp],
    "al3d_tc_configure": [_i, _i],
    "al3d_set_debug_buffer": [_vp],
}
_RESTYPES = {"al3d_last_error": ctypes.c_char_p, "al3d_gemm_split_ws_bytes": ctypes.c_int64, "al3d_gemm_split_tn_ws_bytes": ctypes.c_int64}

_lib = None


def exported_symbols():
    """Names include/al3d.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libal3d.so is not built (%s missing). Build it with "
                "`python -c \"import __graft_entry__ as g; g.build()\"`; there is no fallback path." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if l.al3d_abi_version() != 3:
            raise RuntimeError("libal3d.so ABI version mismatch")
        _lib = l
    return _lib


# kernels launched through this binding since import (bench.py reports it as gpu_launches)
LAUNCHES = 0
_NO_LAUNCH = ("train_ws_floats", "wgrad_ws_floats", "gemm_split_ws_bytes", "gemm_split_tn_ws_bytes", "tc_abort_code", "tc_status_word_host", "tc_configure", "set_debug_buffer", "crop_chunk_points", "crop_occ_words", "crop_hit_bytes")


def check(rc, what=""):
    global LAUNCHES
    if rc == 0 and what not in _NO_LAUNCH:
        LAUNCHES += 1
    if rc != 0:
        msg = lib().al3d_last_error()
        raise RuntimeError("libal3d: %s%s" % (what + ": " if what else "", msg.decode() if msg else "error %d" % rc))
