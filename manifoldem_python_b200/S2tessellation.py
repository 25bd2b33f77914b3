"""Drop-in for the S2 tessellation that produces the per-PD particle lists the distance stage consumes
(modules/S2tessellation.py:67-160, called from Data.op, modules/Data.py:96).

op(q, shAngWidth, PDsizeTh, visual, thres, *fig) -> (CG1, CG, nG, S2, S20_th, S20, NC)

Host: the bin centres (Lovisolo & da Silva's ring construction as modules/distribute3Sphere.py evaluates it, so the
centres are bit-identical), the projection directions of the quaternions, occupancy thresholds.  Device (C ABI
`mem_s2_assign_host`): the nearest bin centre of every direction — the reference's ball-tree query (:59-63).
Grouping the particles by bin is one stable argsort instead of the reference's nG scans of the index array.
"""
import math

import numpy as np

from . import _lib
from .getDistanceCTF_local_Conj9combinedS2 import _ctx


def sphere_points(K, max_iter=100):
    """K points spread over the unit sphere in rings of constant polar angle (distribute3Sphere.op :20-50):
    ring spacing delta, 2 pi sin(w1) / delta points per ring; delta is rescaled by sqrt(found / K) until exactly K
    points come out (at most max_iter passes, then the first K rows of the buffer, as the reference does).
    Returns (points (K,3), passes)."""
    delta = math.exp(math.log(4 * math.pi / K) / 2.)
    buf = np.zeros((2 * K, 3))
    found, passes = 0, 0
    while found != K and passes < max_iter:
        passes += 1
        found = 0
        for w1 in np.arange(0.5 * delta, np.pi, delta):
            c1, s1 = math.cos(w1), math.sin(w1)
            step = delta / s1
            for w2 in np.arange(0.5 * step, 2 * np.pi, step):
                buf[found] = (c1, s1 * math.cos(w2), s1 * math.sin(w2))
                found += 1
        delta = delta * math.exp(math.log(float(found) / K) / 2.)
    return buf[:K], passes


def get_S2(q):
    """(:39-57) projection direction of every quaternion, (3, n)."""
    q = np.asarray(q)
    if q.shape[0] <= 3:
        raise AssertionError('subroutine get_S2: q has wrong dimensions')
    return 2 * np.vstack((q[1, :] * q[3, :] - q[0, :] * q[2, :],
                          q[0, :] * q[1, :] + q[2, :] * q[3, :],
                          q[0, :] ** 2 + q[3, :] ** 2 - 0.5))


def classS2(X, Q, ctx=None):
    """(:59-63) X (nG,3) bin centres, Q (n,3) directions -> (IND (n,1) nearest centre, NC = bincount)."""
    lib = _lib.load()
    ctx = ctx or _ctx()
    X = np.ascontiguousarray(X, dtype=np.float64)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    if X.ndim != 2 or X.shape[1] != 3 or Q.ndim != 2 or Q.shape[1] != 3:
        raise ValueError('classS2 expects (nG,3) centres and (n,3) directions')
    idx = np.empty(Q.shape[0], dtype=np.int32)
    _lib.check(lib.mem_s2_assign_host(ctx.handle, X.ctypes.data, X.shape[0], Q.ctypes.data, Q.shape[0], idx.ctypes.data))
    IND = idx.astype(np.int64).reshape(-1, 1)
    return IND, np.bincount(IND[:, 0])


def _groups(ind, n_bins):
    """Particle indices of every bin, ascending inside a bin — what (IND == i).nonzero()[0] returns for each i."""
    order = np.argsort(ind, kind='stable')
    cuts = np.searchsorted(ind[order], np.arange(n_bins + 1))
    return [order[cuts[i]:cuts[i + 1]] for i in range(n_bins)]


def op(q, shAngWidth, PDsizeTh, visual, thres, *fig):
    nG = np.floor(4 * np.pi / (shAngWidth ** 2)).astype(int)                   # :69
    S20, _ = sphere_points(int(nG))
    S20 = S20.T                                                                # (3, nG)
    S2 = get_S2(q)
    IND, NC = classS2(S20.T, S2.T)
    groups = _groups(IND[:, 0], S20.shape[1])
    CG1 = np.empty(len(groups), dtype=object)                                  # :81-90
    for i, a in enumerate(groups):
        CG1[i] = a
    # lower threshold on one half of the bins (:92-113): NC is as long as the highest occupied bin + 1
    mid = np.floor(S20.shape[1] / 2).astype(int)
    NC1, NC2 = NC[:mid], NC[mid:]
    if len(NC1) >= len(NC2):
        NIND = [pd for pd, occ in enumerate(NC1) if occ >= PDsizeTh]
    else:
        NIND = [mid + pd for pd, occ in enumerate(NC2) if occ >= PDsizeTh]
    S20_th = S20[:, NIND]
    CG = [groups[i][:thres] if len(groups[i]) > thres else groups[i] for i in NIND]   # :126-132 upper threshold
    return (CG1, CG, nG, S2, S20_th, S20, NC)
