"""Heading / size <-> class codecs.  TEST INFRASTRUCTURE (oracle).

Restates tools/utils.py:53-79 of the reference.  Arithmetic is done in whatever type the caller
passes (the reference calls these with python floats / numpy f64 in the datasets and with 0-dim
torch f32 tensors inside StaticModelTwoBoxEst.forward, tools/static_model.py:202).
"""
import numpy as np

NUM_HEADING_BIN = 12
NUM_SIZE_CLUSTER = 3
# tools/utils.py:10-14 (== tools/static_model.py:17-21): car / truck / ped-cyclist anchors (l, w, h)
MEAN_SIZE_ARR = np.array([[4.8, 1.8, 1.5], [10.0, 2.6, 3.2], [2.0, 1.0, 1.6]])

TWO_PI = 2 * np.pi


def angle2class(angle, num_class=NUM_HEADING_BIN):
    """tools/utils.py:53-60.  angle -> (bin id, residual); bins are centred on k*2pi/num_class."""
    a = angle % TWO_PI
    per = TWO_PI / float(num_class)
    shifted = (a + per / 2) % TWO_PI
    cid = int(shifted / per)
    return cid, shifted - (cid * per + per / 2)


def class2angle(pred_cls, residual, num_class=NUM_HEADING_BIN, to_label_format=True):
    """tools/utils.py:69-75.  Inverse of angle2class; wraps to (-pi, pi] when asked."""
    per = TWO_PI / float(num_class)
    ang = pred_cls * per + residual
    if to_label_format and ang > np.pi:
        ang = ang - TWO_PI
    return ang


def size2class(lwh):
    """tools/utils.py:62-67.  Nearest anchor (L2) and the residual to it."""
    d = np.linalg.norm(lwh[np.newaxis, ...] - MEAN_SIZE_ARR, axis=1)
    cid = int(np.argmin(d))
    return cid, lwh - MEAN_SIZE_ARR[cid]


def class2size(pred_cls, residual):
    """tools/utils.py:77-79."""
    return MEAN_SIZE_ARR[pred_cls] + residual


def angle2class_f32(angle_f32, num_class=NUM_HEADING_BIN):
    """The f32 evaluation the reference performs when `angle` is a 0-dim torch f32 tensor
    (tools/static_model.py:202 -> tools/utils.py:53-60): every scalar constant is rounded to f32
    before use and `%` is torch.remainder (fmod + sign fix-up).  Returns (int, np.float32)."""
    f = np.float32
    two_pi = f(TWO_PI)
    per = f(TWO_PI / float(num_class))
    half = f((TWO_PI / float(num_class)) / 2)

    def rem(x, m):
        r = np.fmod(f(x), m)
        if r != 0 and (r < 0) != (m < 0):
            r = f(r + m)
        return f(r)

    a = rem(angle_f32, two_pi)
    shifted = rem(f(a + half), two_pi)
    cid = int(f(shifted / per))
    # class_id * angle_per_class + angle_per_class / 2 is python-float (f64) arithmetic in the
    # reference; the subtraction from the f32 tensor rounds that constant to f32 first.
    centre = f(cid * (TWO_PI / float(num_class)) + (TWO_PI / float(num_class)) / 2)
    return cid, f(shifted - centre)
