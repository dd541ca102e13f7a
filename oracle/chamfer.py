# -*- coding: utf-8 -*-
"""ORACLE -- test infrastructure, not product code.  ctypes wrapper of oracle/chamfer_ref.c plus a
numpy restatement used to cross-check the C build.  PARITY UNPINNED (see chamfer_ref.c header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libchamfer_ref.so')
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, 'chamfer_ref.c')
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):      # never check against a stale build
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.chamfer_ref.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4
        _lib.chamfer_ref.restype = None
    return _lib


def chamfer_c(xyz1, xyz2, nthreads=None):
    """xyz1 [B,N,3], xyz2 [B,M,3] float32 numpy -> dist1, dist2, idx1, idx2."""
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    d1 = np.empty((B, N), np.float32); i1 = np.empty((B, N), np.int32)
    d2 = np.empty((B, M), np.float32); i2 = np.empty((B, M), np.int32)
    _load().chamfer_ref(xyz1.ctypes.data, xyz2.ctypes.data, d1.ctypes.data, i1.ctypes.data, d2.ctypes.data,
                        i2.ctypes.data, B, N, M, nthreads or os.cpu_count() or 1)
    return d1, d2, i1, i2


def chamfer_numpy(xyz1, xyz2):
    """Same arithmetic with numpy fp32 elementwise ops (no FMA: each op is a separate rounding).
    np.argmin returns the first minimum == strict '<' ascending scan."""
    xyz1 = np.asarray(xyz1, np.float32); xyz2 = np.asarray(xyz2, np.float32)

    def one(q, r):
        dx = q[:, None, 0] - r[None, :, 0]
        dy = q[:, None, 1] - r[None, :, 1]
        dz = q[:, None, 2] - r[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz
        i = d.argmin(1).astype(np.int32)
        return d[np.arange(len(q)), i], i

    d1, i1, d2, i2 = [], [], [], []
    for b in range(xyz1.shape[0]):
        a, ai = one(xyz1[b], xyz2[b]); c, ci = one(xyz2[b], xyz1[b])
        d1.append(a); i1.append(ai); d2.append(c); i2.append(ci)
    return np.stack(d1), np.stack(d2), np.stack(i1), np.stack(i2)


def chamfer_distance(xyz1, xyz2):
    """CD = mean(dist1) + mean(dist2), per batch (fp64 means)."""
    d1, d2, _, _ = chamfer_c(xyz1, xyz2)
    return d1.astype(np.float64).mean(1) + d2.astype(np.float64).mean(1)
