"""Stage timings of the DMembeddingII.op drop-in at reference-sized PDs (k = nS, the only way the reference calls it).
python scripts/dm_timing.py [nS ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import DMembeddingII, p   # noqa: E402
from scipy.sparse.linalg import eigsh                  # noqa: E402

p.init()
for nS in [int(a) for a in sys.argv[1:]] or [500, 2000]:
    rng = np.random.default_rng(nS)
    tau = rng.random(nS)
    X = np.stack([np.cos(3 * tau), np.sin(3 * tau), 0.3 * rng.standard_normal(nS)], 1) + 0.05 * rng.standard_normal((nS, 3))
    D = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) * 1e6
    D = (D + D.T) / 2
    for rep in range(2):
        t0 = time.time()
        M, logEps, logSumWij, idx, val = DMembeddingII.graph_and_sweep(D, nS)
        t1 = time.time()
        np.random.seed(0)
        popt, _, _ = DMembeddingII._fit(logEps, logSumWij, np.random.rand(4, 1) - .5)
        sigma = 3.0 * np.sqrt(2 * np.exp(-popt[1] / popt[0]))
        t2 = time.time()
        L = DMembeddingII.laplacian(M, nS, sigma, resident=True)
        M.free()
        t3 = time.time()
        vals, vecs = eigsh(DMembeddingII.device_operator(L, nS), k=16, maxiter=300)
        t4 = time.time()
        Lh = L.download()
        L.free()
        t5 = time.time()
        vals_h, vecs_h = eigsh(Lh, k=16, maxiter=300)
        t6 = time.time()
    print('nS=%d: upload+kNN+graph+Ferguson %.1f ms | curve_fit %.1f ms | Laplacian %.1f ms | ARPACK with device matvec %.1f ms '
          '(host matvec: %.1f ms; max |dlambda| %.1e)'
          % (nS, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t6 - t5) * 1e3,
             np.abs(np.sort(vals) - np.sort(vals_h)).max()))
