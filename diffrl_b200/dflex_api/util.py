"""numpy helpers behind ``import dflex as df`` that asset loaders and envs call (quaternions are
(x, y, z, w); a transform is a ``(position, quaternion)`` pair of arrays).

Interface mirror of the reference's ``dflex/dflex/util.py`` (names + argument meaning); written
independently.  Arithmetic is kept in the same operation order where it decides the bits of the
finalized Model tensors (the golden-model parity test compares them exactly), including the
reference's element-wise ``R * I * R.T`` in :func:`transform_inertia` (util.py:235-239).
"""
import cProfile
import math
import timeit

import numpy as np


# ------------------------------------------------------------------ small vector helpers
def length(a):
    return np.linalg.norm(a)


def length_sq(a):
    return np.dot(a, a)


def normalize(v):
    n = np.linalg.norm(v)
    return v if n == 0.0 else v / n


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


# ------------------------------------------------------------------ quaternions
def quat(i, j, k, w):
    return np.array([i, j, k, w])


def quat_identity():
    return np.array((0.0, 0.0, 0.0, 1.0))


def quat_inverse(q):
    return np.array((-q[0], -q[1], -q[2], q[3]))


def quat_from_axis_angle(axis, angle):
    half = angle * 0.5
    v = np.array(axis) * math.sin(half)
    return np.array((v[0], v[1], v[2], math.cos(half)))


def quat_rotate(q, x):
    x = np.array(x)
    im = np.array((q[0], q[1], q[2]))
    return x * (2.0 * q[3] * q[3] - 1.0) + np.cross(im, x) * q[3] * 2.0 + im * np.dot(im, x) * 2.0


def quat_multiply(a, b):
    return np.array((a[3] * b[0] + b[3] * a[0] + a[1] * b[2] - b[1] * a[2],
                     a[3] * b[1] + b[3] * a[1] + a[2] * b[0] - b[2] * a[0],
                     a[3] * b[2] + b[3] * a[2] + a[0] * b[1] - b[0] * a[1],
                     a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]))


def quat_to_matrix(q):
    cols = [quat_rotate(q, np.array(e)) for e in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))]
    return np.array(cols).T


def rpy2quat(roll, pitch, yaw):
    cy, sy = math.cos(yaw * 0.5), math.sin(yaw * 0.5)
    cr, sr = math.cos(roll * 0.5), math.sin(roll * 0.5)
    cp, sp = math.cos(pitch * 0.5), math.sin(pitch * 0.5)
    w = cy * cr * cp + sy * sr * sp
    x = cy * sr * cp - sy * cr * sp
    y = cy * cr * sp + sy * sr * cp
    z = sy * cr * cp - cy * sr * sp
    return (x, y, z, w)


quat_rpy = rpy2quat


def quat_from_matrix(m):
    """Rotation matrix -> unit quaternion (largest-pivot branch selection)."""
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr >= 0.0:
        h = math.sqrt(tr + 1.0)
        w = 0.5 * h
        h = 0.5 / h
        x, y, z = (m[2, 1] - m[1, 2]) * h, (m[0, 2] - m[2, 0]) * h, (m[1, 0] - m[0, 1]) * h
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        if i == 0:
            h = math.sqrt((m[0, 0] - (m[1, 1] + m[2, 2])) + 1.0)
            x = 0.5 * h
            h = 0.5 / h
            y, z, w = (m[0, 1] + m[1, 0]) * h, (m[2, 0] + m[0, 2]) * h, (m[2, 1] - m[1, 2]) * h
        elif i == 1:
            h = math.sqrt((m[1, 1] - (m[2, 2] + m[0, 0])) + 1.0)
            y = 0.5 * h
            h = 0.5 / h
            z, x, w = (m[1, 2] + m[2, 1]) * h, (m[0, 1] + m[1, 0]) * h, (m[0, 2] - m[2, 0]) * h
        else:
            h = math.sqrt((m[2, 2] - (m[0, 0] + m[1, 1])) + 1.0)
            z = 0.5 * h
            h = 0.5 / h
            x, y, w = (m[2, 0] + m[0, 2]) * h, (m[1, 2] + m[2, 1]) * h, (m[1, 0] - m[0, 1]) * h
    return normalize(quat(x, y, z, w))


# ------------------------------------------------------------------ rigid transforms
def transform(x, r):
    return (np.array(x), np.array(r))


def transform_identity():
    return (np.array((0.0, 0.0, 0.0)), quat_identity())


def transform_inverse(t):
    q_inv = quat_inverse(t[1])
    return (-quat_rotate(q_inv, t[0]), q_inv)


def transform_vector(t, v):
    return quat_rotate(t[1], v)


def transform_point(t, p):
    return np.array(t[0]) + quat_rotate(t[1], p)


def transform_multiply(t, u):
    return (quat_rotate(t[1], u[0]) + t[0], quat_multiply(t[1], u[1]))


def transform_flatten(t):
    return np.array([*t[0], *t[1]])


def transform_expand(t):
    return (np.array(t[0:3]), np.array(t[3:7]))


def transform_flatten_list(xforms):
    return [transform_flatten(t) for t in xforms]


def transform_expand_list(xforms):
    return [transform_expand(t) for t in xforms]


def transform_inertia(m, I, p, q):
    """Inertia of a shape expressed about a point offset by ``p`` (Steiner).  NOTE: like the
    reference this multiplies ELEMENT-WISE (ndarray ``*``), which only equals R I R^T for diagonal I
    and axis-aligned R; kept because it defines the reference's body_I_m values."""
    R = quat_to_matrix(q)
    return R * I * R.T + m * (np.dot(p, p) * np.eye(3) - np.outer(p, p))


def spatial_matrix_from_inertia(I, m):
    G = np.zeros((6, 6))
    G[0:3, 0:3] = I
    G[3, 3] = G[4, 4] = G[5, 5] = m
    return G


# ------------------------------------------------------------------ timing
_log = []


def log(s):
    print(s)
    _log.append(s)


class ScopedTimer:
    """``with df.ScopedTimer(name, active=True, detailed=False):`` -- the envs wrap their simulate / reset / render
    sections in it (``envs/cartpole_swing_up.py:114-148``), almost always with ``active=False``.

    An inactive timer costs nothing.  An active one reports the wall-clock time of the block; on a CUDA build it brackets
    the block with a device synchronisation so that the figure covers the kernels the block launched, not just their
    enqueueing (the fused step returns before the GPU has run it).  ``detailed`` adds a cProfile table.  Nested timers
    indent their report by depth."""

    _depth = 0
    enabled = True
    sync_cuda = True

    def __init__(self, name, active=True, detailed=False):
        self.name = name
        self.active = bool(active) and ScopedTimer.enabled
        self.detailed = bool(detailed)
        self.elapsed_ms = None
        self._t0 = self._profile = None

    @staticmethod
    def _sync():
        if ScopedTimer.sync_cuda:
            try:
                import torch
                if torch.cuda.is_available() and torch.cuda.is_initialized():
                    torch.cuda.synchronize()
            except Exception:
                pass

    def __enter__(self):
        if not self.active:
            return self
        ScopedTimer._depth += 1
        if self.detailed:
            self._profile = cProfile.Profile()
            self._profile.enable()
        self._sync()
        self._t0 = timeit.default_timer()
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        if not self.active:
            return False
        self._sync()
        self.elapsed_ms = 1e3 * (timeit.default_timer() - self._t0)
        if self._profile is not None:
            self._profile.disable()
            self._profile.print_stats(sort="tottime")
        ScopedTimer._depth -= 1
        log("%s%s took %.2f ms" % ("\t" * ScopedTimer._depth, self.name, self.elapsed_ms))
        return False
