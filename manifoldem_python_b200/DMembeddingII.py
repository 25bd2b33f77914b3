"""Drop-in for the diffusion-map front end (modules/DMembeddingII.py:86-185).

op(D, k, tune, prefsigma) -> (lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared)

On the GPU (C ABI): kNN lists (a15), OR-symmetrised graph (a16), the 1,501-point Ferguson sweep (a17 — 97 %
of the reference's time), Gaussian kernel + alpha = 1 + symmetric normalisations (a18).  On the host, with the
same SciPy calls as the reference: the 4-parameter tanh fit (curve_fit, fergusonE.py:45-57) and the ARPACK
eigen-solve (eigsh, sembeddingonFly.py:27-35), so sigma and the eigenvectors follow the same code path.
D is mutated in place (diagonal <- -inf) like the reference (:48).
"""
import ctypes as C

import numpy as np
from scipy.optimize import curve_fit, OptimizeWarning
from scipy.sparse.linalg import eigsh, ArpackNoConvergence
import warnings

from . import _lib
from .getDistanceCTF_local_Conj9combinedS2 import _cfg, _ctx

warnings.simplefilter(action='ignore', category=OptimizeWarning)


def fun(xx, aa0, aa1, aa2, aa3):
    """fergusonE.fun :21-23."""
    return aa3 + aa2 * np.tanh(aa0 * xx + aa1)


def _fit(logEps, logSumWij, a0):
    """fergusonE.op :45-57 — retry from random starts while sum sqrt|diag pcov| > 100."""
    resnorm = np.inf
    while resnorm > 100:
        popt, pcov = curve_fit(fun, logEps, logSumWij, p0=a0)
        resnorm = sum(np.sqrt(np.fabs(np.diag(pcov))))
        a0 = 1 * (np.random.rand(4, 1) - .5)
        residuals = logSumWij - fun(logEps, popt[0], popt[1], popt[2], popt[3])
        ss_res = np.sum(residuals ** 2)
        ss_tot = np.sum((logSumWij - np.mean(logSumWij)) ** 2)
        R_squared = 1 - (ss_res / ss_tot)
    return popt, resnorm, R_squared


def graph_and_sweep(D, k, ctx=None, nS=None, knn=None, want_lists=True):
    """Device part up to the Ferguson curve.  `D`: host (nS,nS) array, or a `_lib.DeviceArray` holding the
    float32 D that the distance stage left on the device (then D never visits the host), or None with
    `knn` = (idx, val) device arrays that the distance stage selected itself (pd_stage.run_pd_resident(knn_k=k,
    keep_D=False): D was never assembled; the lists are consumed and freed here).
    Returns (M_dev DeviceArray (nS,nS) float64 graph, logEps, logSumWij, idx (nS,k) int32, val (nS,k) float64);
    idx, val are None with want_lists=False (op() does not need them: at k = nS = 2,000 their download is 48 MB)."""
    lib = _lib.load()
    ctx = ctx or _ctx()
    Dd = None
    if knn is not None:
        idx_d, val_d = knn
        nS = idx_d.shape[0]
        if idx_d.shape != (nS, k) or val_d.shape != (nS, k) or idx_d.dtype != np.int32 or val_d.dtype != np.float64:
            raise TypeError('knn lists must be (nS,k) int32 / float64 device arrays')
    else:
        nS = D.shape[0]
        idx_d = _lib.DeviceArray(ctx, (nS, k), np.int32)
        val_d = _lib.DeviceArray(ctx, (nS, k), np.float64)
        if isinstance(D, _lib.DeviceArray):
            if D.dtype not in (np.float32, np.float64) or D.shape != (nS, nS):
                raise TypeError('resident D must be a square float32 or float64 device array')
            knn = lib.mem_knn_device_f32 if D.dtype == np.float32 else lib.mem_knn_device
            _lib.check(knn(ctx.handle, D.ptr, nS, k, idx_d.ptr, val_d.ptr, None))
        else:
            Dd = _lib.DeviceArray(ctx, (nS, nS), np.float64, np.ascontiguousarray(D, dtype=np.float64))
            _lib.check(lib.mem_knn_device(ctx.handle, Dd.ptr, nS, k, idx_d.ptr, val_d.ptr, None))
    M = _lib.DeviceArray(ctx, (nS, nS), np.float64)
    _lib.check(lib.mem_graph_dense_device(ctx.handle, idx_d.ptr, val_d.ptr, nS, k, M.ptr, None))
    # k < nS: sweep over the edges only (row-major compaction), not over the nS^2 slots of the dense graph
    vals, n_vals, compact = M, nS * nS, None
    if k < nS:
        compact = _lib.DeviceArray(ctx, (min(nS * nS, 2 * nS * k),), np.float64)
        cnt = C.c_int64()
        _lib.check(lib.mem_graph_compact_device(ctx.handle, M.ptr, nS, compact.ptr, C.byref(cnt)))
        vals, n_vals = compact, int(cnt.value)
    logEps = np.arange(-150, 150.2, 0.2)                                        # :146
    # find_thres (fergusonE.py:25-31): ss = sum exp(-d2 / (2 max eps)), n = number of graph entries
    one = np.array([np.max(logEps)], dtype=np.float64)
    out1 = np.zeros(1)
    _lib.check(lib.mem_ferguson_device(ctx.handle, vals.ptr, n_vals, one.ctypes.data, 1, float('inf'), out1.ctypes.data))
    ss = float(np.exp(out1[0]))
    if compact is not None:
        n_entries = float(n_vals)
    else:
        huge = np.array([700.0], dtype=np.float64)
        _lib.check(lib.mem_ferguson_device(ctx.handle, vals.ptr, n_vals, huge.ctypes.data, 1, float('inf'), out1.ctypes.data))
        n_entries = float(np.rint(np.exp(out1[0])))
    thr = max(-np.log(0.01 * ss / n_entries), 10)
    logSumWij = np.zeros(len(logEps))
    _lib.check(lib.mem_ferguson_device(ctx.handle, vals.ptr, n_vals, logEps.ctypes.data, len(logEps), float(thr),
                                       logSumWij.ctypes.data))
    idx, val = (idx_d.download(), val_d.download()) if want_lists else (None, None)
    for a in (Dd, idx_d, val_d, compact):
        if a is not None:
            a.free()
    return M, logEps, logSumWij, idx, val


def laplacian(M, nS, sigma, ctx=None, resident=False):
    """slaplacianonFly.op on the device: dense (nS,nS) float64 L, downloaded to the host or (resident=True)
    left on the device as a `_lib.DeviceArray`."""
    lib = _lib.load()
    ctx = ctx or _ctx()
    L = _lib.DeviceArray(ctx, (nS, nS), np.float64)
    _lib.check(lib.mem_laplacian_dense_device(ctx.handle, M.ptr, nS, float(sigma), L.ptr, None))
    if resident:
        return L
    out = L.download()
    L.free()
    return out


def device_operator(L_dev, nS, ctx=None):
    """scipy LinearOperator whose matvec runs on the device (L never leaves it): ARPACK stays on the host and
    follows the reference's own eigsh call, only the O(nS^2) operator application moves."""
    from scipy.sparse.linalg import LinearOperator
    lib = _lib.load()
    ctx = ctx or _ctx()
    y = np.empty(nS, dtype=np.float64)

    def matvec(x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        _lib.check(lib.mem_symv_host(ctx.handle, L_dev.ptr, nS, x.ctypes.data, y.ctypes.data))
        return y.copy()
    return LinearOperator((nS, nS), matvec=matvec, dtype=np.float64)


def upload(D, ctx=None):
    """D on the device for embed() / take(): float32 stays float32 (the dtype the distance stage computed it in and the
    sidecar record stores — half the bytes, identical lists), anything else goes up as float64."""
    D = np.asarray(D)
    dt = np.float32 if D.dtype == np.float32 else np.float64
    return _lib.DeviceArray(ctx or _ctx(), D.shape, dt, np.ascontiguousarray(D, dtype=dt))


def take(D_dev, sel, ctx=None):
    """D[sel][:, sel] on the device (manifoldTrimmingAuto.py:50,63): a new (m,m) device array of the same dtype."""
    lib = _lib.load()
    ctx = ctx or _ctx()
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    m = sel.shape[0]
    out = _lib.DeviceArray(ctx, (m, m), D_dev.dtype)
    _lib.check(lib.mem_gather_square_device(ctx.handle, D_dev.ptr, D_dev.dtype.itemsize, D_dev.shape[0], sel.ctypes.data, m,
                                            out.ptr, None))
    return out


def op(D, k, tune, prefsigma):
    out = embed(D, int(k), tune)
    nS = D.shape[0]
    D[np.arange(nS), np.arange(nS)] = -np.inf                                    # :48, in place like the reference
    return out


def embed(D, k, tune):
    """op() without the side effect on the caller's matrix; D may also be a device array from upload() / take() or
    pd_stage.run_pd_resident — the trimming loop re-embeds subsets of one resident D."""
    p = _cfg()
    nS = D.shape[0]
    M, logEps, logSumWij, _, _ = graph_and_sweep(D, k, want_lists=False)
    a0 = (np.random.rand(4, 1) - .5)                                             # :142 (unseeded in the reference)
    popt, resnorm, R_squared = _fit(logEps, logSumWij, a0)
    nEigs = min(getattr(p, 'num_eigs', 15), nS - 3)                              # :149
    sigma = tune * np.sqrt(2 * np.exp(-popt[1] / popt[0]))                       # :158
    L = laplacian(M, nS, sigma, resident=True)
    M.free()
    try:
        vals, vecs = eigsh(device_operator(L, nS), k=nEigs + 1, maxiter=300)     # sembeddingonFly.py:27
    except ArpackNoConvergence as e:
        vals, vecs = e.eigenvalues, e.eigenvectors
        print("eigsh not converging in 300 iterations...")
    finally:
        L.free()
    ix = np.argsort(vals)[::-1]
    lamb = np.sort(vals)[::-1]
    v = vecs[:, ix]
    true_shape = v.shape[1] - 1
    psi = np.zeros((v.shape[0], nEigs))
    psi[:, :true_shape] = v[:, 1:] / np.tile(v[:, 0].reshape((-1, 1)), (1, true_shape))    # :175-177
    mu = v[:, 0]
    mu = mu * mu                                                                 # :181-182
    return (lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared)
