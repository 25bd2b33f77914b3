"""Drop-in for the diffusion-map front end (modules/DMembeddingII.py:86-185).

op(D, k, tune, prefsigma) -> (lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared)

On the GPU (C ABI): kNN lists (a15), OR-symmetrised graph (a16), the 1,501-point Ferguson sweep (a17 — 97 %
of the reference's time), Gaussian kernel + alpha = 1 + symmetric normalisations (a18), and the eigen-solve (a19,
sembeddingonFly.py:27-35) as Lanczos with full re-orthogonalisation on the resident Laplacian (eigsh_device; the
reference's ARPACK call with the operator applied on the device stays selectable, p.eig_solver = 'arpack').  On the
host, with the same SciPy call as the reference: the 4-parameter tanh fit (curve_fit, fergusonE.py:45-57).
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


def lanczos_ritz_selection(alpha, beta, k, nS, tol):
    """Host part of eigsh_device between two blocks of steps: alpha (j,), beta (j + 1,) with beta[i] = ||w|| after step
    i - 1 (beta[0] unused).  Returns (theta (k',), S (j_eff, k') Ritz coefficients, j_eff, converged) for the k' <= k
    Ritz pairs of largest magnitude (ARPACK's which='LM').  A beta that vanishes (invariant subspace, or j = nS) truncates
    the recurrence there and makes the pairs exact."""
    from scipy.linalg import eigh_tridiagonal
    j = alpha.shape[0]
    scale = max(np.abs(alpha).max(), np.abs(beta[1:j + 1]).max(), np.finfo(float).tiny)
    small = np.nonzero(beta[1:j + 1] <= 1e-13 * scale)[0]
    exact = small.size > 0 or j >= nS
    j_eff = min(j, nS, int(small[0]) + 1 if small.size else j)
    if j_eff == 1:
        theta, S = alpha[:1].copy(), np.ones((1, 1))
    else:
        theta, S = eigh_tridiagonal(alpha[:j_eff], beta[1:j_eff])
    order = np.argsort(-np.abs(theta), kind='stable')[:k]
    theta, S = theta[order], S[:, order]
    if exact:
        return theta, S, j_eff, True
    res = np.abs(beta[j_eff]) * np.abs(S[-1, :])
    ok = res <= tol * np.maximum(np.finfo(float).eps ** (2.0 / 3.0), np.abs(theta))
    return theta, S, j_eff, bool(ok.all()) and theta.shape[0] == k


def eigsh_device(L_dev, nS, k, tol=1e-12, block=16, m_max=None, ctx=None):
    """The k eigenpairs of largest magnitude of the symmetric (nS,nS) float64 device matrix L_dev — what
    sembeddingonFly.py:27 asks ARPACK for — by Lanczos with full re-orthogonalisation on the device (eig.cu): the host
    enqueues `block` steps at a time and only reads the 2 x j recurrence coefficients in between.
    Returns (vals (k,), vecs (nS,k), info) in ARPACK's layout (unit-norm columns, no particular order);
    info = dict(steps, converged).  Convergence: |beta_j s_ji| <= tol * max(eps^(2/3), |theta_i|) for every wanted pair."""
    lib = _lib.load()
    ctx = ctx or _ctx()
    k = int(min(k, nS - 1))
    m_max = int(min(nS, m_max or max(20 * k, 640)))
    V = _lib.DeviceArray(ctx, (m_max + 1, nS), np.float64)
    ab = _lib.DeviceArray(ctx, (2, m_max + 1), np.float64)
    j = 0
    try:
        while True:
            j1 = min(m_max, max(j + block, 2 * k + 1) if j == 0 else j + block)
            _lib.check(lib.mem_lanczos_steps_device(ctx.handle, L_dev.ptr, nS, V.ptr, ab.ptr, m_max + 1, j, j1, None))
            j = j1
            c = ab.download()
            theta, S, j_eff, done = lanczos_ritz_selection(c[0, :j], c[1, :j + 1], k, nS, tol)
            if done or j >= m_max:
                break
        kk = theta.shape[0]
        X = _lib.DeviceArray(ctx, (kk, nS), np.float64)
        Sc = np.ascontiguousarray(S, dtype=np.float64)
        _lib.check(lib.mem_lanczos_ritz_device(ctx.handle, V.ptr, nS, j_eff, Sc.ctypes.data, kk, X.ptr, None))
        vecs = X.download().T.copy()
        X.free()
    finally:
        V.free()
        ab.free()
    return theta, vecs, dict(steps=j, converged=done)


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
        if getattr(p, 'eig_solver', 'lanczos') == 'arpack':
            # the reference's own call (sembeddingonFly.py:27) with the operator applied on the device: kept for cross-checks
            try:
                vals, vecs = eigsh(device_operator(L, nS), k=nEigs + 1, maxiter=300)
            except ArpackNoConvergence as e:
                vals, vecs = e.eigenvalues, e.eigenvectors
                print("eigsh not converging in 300 iterations...")
        else:
            vals, vecs, info = eigsh_device(L, nS, nEigs + 1)
            if not info['converged']:
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
