"""Oracle: kNN graph, Ferguson bandwidth, Gaussian-kernel Laplacian, embedding.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates

    modules/DMembeddingII.py:43-185      (initialize, op)
    modules/fergusonE.py:21-59           (find_thres, op)
    modules/slaplacianonFly.py:42-81     (op)
    modules/sembeddingonFly.py:17-39     (op)

Third-party arithmetic at the same call sites as the reference:
scipy.optimize.curve_fit, scipy.sparse, scipy.sparse.linalg.eigsh (ARPACK).
"""
import numpy as np
from scipy.optimize import curve_fit
from scipy.sparse import csc_matrix
from scipy.sparse.linalg import eigsh, ArpackNoConvergence

LOG_EPS = np.arange(-150, 150.2, 0.2)       # DMembeddingII.py:146


# --------------------------------------------------------------------- a15
def knn_lists(D, k):
    """DMembeddingII.initialize :43-57.  MUTATES D (diag <- -inf) like the
    reference.  Returns (idx (k,nS) int32, val (k,nS) f64), column i = the k
    smallest entries of D[:, i] in ascending order, self first with value 0."""
    nS = D.shape[0]
    D[np.arange(nS), np.arange(nS)] = -np.inf
    IX = np.argsort(D, axis=0)[:k]
    val = np.take_along_axis(D, IX, axis=0)
    val[0, :] = 0
    return IX.astype(np.int32), val


# --------------------------------------------------------------------- a16
def symmetrise(idx, val, nS):
    """DMembeddingII.op :97-140 — OR-symmetrised graph in COO form.

    Entries with d^2 < 1e-6 are the 'zeros' (kept, value 0); the rest are
    scattered as d into a dense matrix y and symmetrised as
    y^2 + (y^2)^T - y*y^T, i.e. the union kNN graph carrying d^2.
    Output order: zeros (row-major) ++ non-zeros (row-major), as :121-140.
    """
    k = idx.shape[0]
    yVal = val.flatten('F')                   # :56 / :24-30, row i of the graph = column i of D
    yCol = idx.flatten('F').astype(np.int64)
    yRow = np.repeat(np.arange(nS), k)        # :113-114
    is_zero = yVal < 1e-6                     # :115
    y = np.zeros((nS, nS))
    np.add.at(y, (yRow[~is_zero], yCol[~is_zero]), np.sqrt(yVal[~is_zero]))   # csr_matrix(...).toarray() sums duplicates
    y2 = y * y.T
    y = y ** 2
    y = y + y.T - y2                          # :73-78
    rNZ, cNZ = y.nonzero()
    vNZ = y[rNZ, cNZ]
    z = np.zeros((nS, nS))
    np.add.at(z, (yRow[is_zero], yCol[is_zero]), 1.0)
    rZ, cZ = z.nonzero()
    yRow = np.hstack((rZ, rNZ)).astype(int)
    yCol = np.hstack((cZ, cNZ)).astype(int)
    yVal = np.hstack((np.zeros(len(rZ)), vNZ))
    return yRow, yCol, yVal


# --------------------------------------------------------------------- a17
def ferguson_threshold(logEps, D2):
    """fergusonE.find_thres :25-31."""
    d = 1. / (2. * np.max(np.exp(logEps))) * D2
    ss = np.sum(np.exp(-np.sort(d)))
    return max(-np.log(0.01 * ss / len(D2)), 10)


def ferguson_logsum(dist, logEps=LOG_EPS):
    """fergusonE.op :36-43 — log sum_{d2/2eps < thr} exp(-d2/2eps) per eps."""
    D2 = dist * dist
    thr = ferguson_threshold(logEps, D2)
    out = np.zeros(len(logEps))
    for i, le in enumerate(logEps):
        d = 1. / (2. * np.exp(le)) * D2
        out[i] = np.log(np.sum(np.exp(-d[d < thr])))
    return out, thr


def tanh_model(xx, aa0, aa1, aa2, aa3):
    """fergusonE.fun :21-23."""
    return aa3 + aa2 * np.tanh(aa0 * xx + aa1)


def ferguson_fit(logEps, logSumWij, a0, rng=None):
    """fergusonE.op :45-57 — retry from random starts while sum sqrt|diag pcov| > 100."""
    rng = np.random if rng is None else rng
    resnorm = np.inf
    while resnorm > 100:
        popt, pcov = curve_fit(tanh_model, logEps, logSumWij, p0=np.ravel(a0))
        resnorm = np.sum(np.sqrt(np.fabs(np.diag(pcov))))
        a0 = rng.rand(4, 1) - .5
        res = logSumWij - tanh_model(logEps, *popt)
        R2 = 1 - np.sum(res ** 2) / np.sum((logSumWij - np.mean(logSumWij)) ** 2)
    return popt, resnorm, R2


# --------------------------------------------------------------------- a18
def laplacian(yVal, yCol, yRow, nS, sigma, alpha=1.0):
    """slaplacianonFly.op :42-81 — Gaussian kernel, Coifman-Lafon alpha
    normalisation, symmetric normalisation; returns sparse CSC L."""
    w = np.exp(-yVal / sigma ** 2)
    l = csc_matrix((w, (yRow, yCol)), shape=(nS, nS))
    d = np.asarray(l.sum(axis=0)).ravel()
    if alpha != 1:
        d = d ** alpha
    w = w / (d[yRow] * d[yCol])
    l = csc_matrix((w, (yRow, yCol)), shape=(nS, nS))
    d = np.sqrt(np.asarray(l.sum(axis=0)).ravel())
    w = w / (d[yRow] * d[yCol])
    l = csc_matrix((w, (yRow, yCol)), shape=(nS, nS))
    return abs(l + l.T) / 2.0


# --------------------------------------------------------------------- a19
def embed(L, nEigs):
    """sembeddingonFly.op :27-35."""
    try:
        vals, vecs = eigsh(L, k=nEigs + 1, maxiter=300)
    except ArpackNoConvergence as e:          # pragma: no cover
        vals, vecs = e.eigenvalues, e.eigenvectors
    ix = np.argsort(vals)[::-1]
    return np.sort(vals)[::-1], vecs[:, ix]


def dm_embedding(D, k, tune, prefsigma=None, num_eigs=15, a0=None, rng=None):
    """DMembeddingII.op :86-185.  D is mutated (diag <- -inf) like the reference.
    a0 / rng pin the otherwise unseeded random start (:142)."""
    nS = D.shape[0]
    idx, val = knn_lists(D, k)
    yRow, yCol, yVal = symmetrise(idx, val, nS)
    if a0 is None:
        a0 = (np.random if rng is None else rng).rand(4, 1) - .5
    logEps = LOG_EPS
    logSumWij, _ = ferguson_logsum(np.sqrt(yVal), logEps)
    popt, _, R2 = ferguson_fit(logEps, logSumWij, a0, rng)
    nEigs = min(num_eigs, nS - 3)
    sigma = tune * np.sqrt(2 * np.exp(-popt[1] / popt[0]))       # :158
    L = laplacian(yVal, yCol, yRow, nS, sigma)
    lamb, v = embed(L, nEigs)
    psi = np.zeros((nS, nEigs))
    psi[:, :v.shape[1] - 1] = v[:, 1:] / v[:, [0]]               # :175-177
    mu = v[:, 0] ** 2
    return lamb, psi, sigma, mu, logEps, logSumWij, popt, R2
