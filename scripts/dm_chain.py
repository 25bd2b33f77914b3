"""The device part of DMembeddingII.op (kNN, graph, compaction, Ferguson sweep, Laplacian) on a synthetic D —
the command ncu wraps for the launch list of rows a15-a18.
    python scripts/dm_chain.py [nS] [k] [reps]          k = 0 -> k = nS (how manifoldTrimmingAuto calls it)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import DMembeddingII, _lib, p   # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
k = k or nS
p.init()
ctx = DMembeddingII._ctx()
rng = np.random.default_rng(nS)
tau = rng.random(nS)
X = np.stack([np.cos(3 * tau), np.sin(3 * tau), 0.3 * rng.standard_normal(nS)], 1) + 0.05 * rng.standard_normal((nS, 3))
D = (((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) * 1e6).astype(np.float32)
D = np.maximum(D, D.T)
Dd = _lib.DeviceArray(ctx, (nS, nS), np.float32, D)
for r in range(reps):
    t0 = time.time()
    M, logEps, ls, idx, val = DMembeddingII.graph_and_sweep(Dd, k)
    t1 = time.time()
    L = DMembeddingII.laplacian(M, nS, 3.0 * np.sqrt(np.median(val[:, 1:])), resident=True)
    ctx.sync()
    t2 = time.time()
    M.free()
    L.free()
    print('rep %d nS=%d k=%d: kNN+graph+sweep %.2f ms, Laplacian %.2f ms, checksum %.17g' %
          (r, nS, k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, float(ls.sum())))
