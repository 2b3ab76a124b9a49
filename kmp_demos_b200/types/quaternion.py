"""``quaternion`` - drop-in for the reference's ``kmp.types.quaternion.quaternion`` (kmp/types/quaternion.py).

Every numerical map (constructor normalisation, exp, log, product, conjugate, from_rotation_vector) is a
batched device kernel (kmpx_quat_*); the scalar class calls them with one element, the ``*_batch`` functions
with (M,4)/(M,3) arrays (what demo2's per-sample loops and create_dataset amount to, demos.py:274-275,292-293).
Component-wise +, -, /, abs are host glue on 4 numbers followed by the device normalisation."""
from __future__ import annotations

import numpy as np

from .. import _lib


def _dev(a, cols):
    torch = _lib.require_cuda()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, cols)), device='cuda')


def normalize_batch(Q):
    torch = _lib.require_cuda()
    q = _dev(Q, 4)
    out = torch.empty_like(q)
    _lib.call('kmpx_quat_normalize', _lib.ptr(q), _lib.ptr(out), q.shape[0], _lib.stream())
    return out.cpu().numpy()


def exp_batch(W):
    """(M,3) tangent vectors -> (M,4) unit quaternions (quaternion.py:60-80)."""
    torch = _lib.require_cuda()
    w = _dev(W, 3)
    out = torch.empty((w.shape[0], 4), dtype=torch.float64, device='cuda')
    _lib.call('kmpx_quat_exp', _lib.ptr(w), _lib.ptr(out), w.shape[0], _lib.stream())
    return out.cpu().numpy()


def log_batch(Q):
    """(M,4) quaternions -> (M,3) tangent vectors (quaternion.py:82-96)."""
    torch = _lib.require_cuda()
    q = _dev(Q, 4)
    out = torch.empty((q.shape[0], 3), dtype=torch.float64, device='cuda')
    _lib.call('kmpx_quat_log', _lib.ptr(q), _lib.ptr(out), q.shape[0], _lib.stream())
    return out.cpu().numpy()


def mul_batch(P, Q, conj_q=False):
    """Hamilton products P[i]*Q[i] (or P[i]*Q[0] when Q holds one quaternion), optionally with ~Q
    (quaternion.py:182-202,219-227); results normalised."""
    torch = _lib.require_cuda()
    p, q = _dev(P, 4), _dev(Q, 4)
    bcast = int(q.shape[0] == 1 and p.shape[0] != 1)
    if not bcast and q.shape[0] != p.shape[0]:
        raise ValueError('mul_batch: operand counts differ')
    out = torch.empty_like(p)
    _lib.call('kmpx_quat_mul', _lib.ptr(p), _lib.ptr(q), bcast, int(conj_q), _lib.ptr(out), p.shape[0], _lib.stream())
    return out.cpu().numpy()


def from_rotvec_batch(R):
    torch = _lib.require_cuda()
    r = _dev(R, 3)
    out = torch.empty((r.shape[0], 4), dtype=torch.float64, device='cuda')
    _lib.call('kmpx_quat_from_rotvec', _lib.ptr(r), _lib.ptr(out), r.shape[0], _lib.stream())
    return out.cpu().numpy()


class quaternion:
    """Quaternion with scalar part ``v`` and vector part ``u`` (shape (3,)); normalised on construction."""

    def __init__(self, v, u) -> None:
        self.v = v
        self.u = u
        self.normalize()

    @classmethod
    def _from_unit(cls, arr):
        q = cls.__new__(cls)
        q.v = float(arr[0])
        q.u = np.array(arr[1:], dtype=np.float64)
        return q

    @classmethod
    def from_rotation_vector(cls, rotation_vector):
        return cls._from_unit(from_rotvec_batch(rotation_vector)[0])

    @classmethod
    def from_array(cls, array):
        return cls(array[0], array[1:])

    @classmethod
    def exp(cls, w):
        return cls._from_unit(exp_batch(w)[0])

    def log(self):
        return log_batch(self.as_array())[0]

    def abs(self):
        return np.concatenate(([np.abs(self.v)], np.abs(self.u)))

    def as_array(self):
        return np.concatenate(([self.v], self.u))

    def normalize(self):
        arr = normalize_batch(np.concatenate(([self.v], np.asarray(self.u, dtype=np.float64))))[0]
        self.v = float(arr[0])
        self.u = arr[1:].copy()
        return self

    def derivative(self, wh, qa, dt):
        xi1 = ((self.exp(wh * dt / 2) * self) * ~qa).log()
        xi2 = (self * ~qa).log()
        return (xi1 - xi2) / dt

    def __add__(self, q):
        return quaternion(self.v + q.v, self.u + q.u)

    def __sub__(self, q):
        return quaternion(self.v - q.v, self.u - q.u)

    def __mul__(self, q):
        if isinstance(q, quaternion):
            return quaternion._from_unit(mul_batch(self.as_array(), q.as_array())[0])
        return quaternion(self.v * q, self.u * q)

    def __truediv__(self, q):
        return quaternion(self.v / q, self.u / q)

    def __invert__(self):
        return quaternion(self.v, -self.u)

    def __neg__(self):
        return quaternion(-self.v, -self.u)

    def __pos__(self):
        return quaternion(self.v, self.u)

    def __getitem__(self, index):
        return self.v if index == 0 else self.u[index - 1]

    def __setitem__(self, index, value):
        if index == 0:
            self.v = value
        else:
            self.u[index - 1] = value
