"""a19 timing: device Lanczos (eigsh_device) vs the reference's ARPACK call with the operator on the device vs ARPACK on
a host matrix.   python scripts/eig_timing.py [nS ...]"""
import os
import sys
import time

import numpy as np
from scipy.sparse.linalg import eigsh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from manifoldem_python_b200 import DMembeddingII, _lib            # noqa: E402
from test_gpu_dm_dropin import _diffusion_like_matrix             # noqa: E402

ctx = _lib.Context(0)
for nS in [int(a) for a in sys.argv[1:]] or [500, 2000, 5000]:
    L = _diffusion_like_matrix(nS, 1)
    Ld = _lib.DeviceArray(ctx, (nS, nS), np.float64, L)
    k = 16
    for rep in range(3):
        t0 = time.perf_counter()
        vals, vecs, info = DMembeddingII.eigsh_device(Ld, nS, k, ctx=ctx)
        t1 = time.perf_counter()
    op = DMembeddingII.device_operator(Ld, nS, ctx=ctx)
    t2 = time.perf_counter()
    va, ve = eigsh(op, k=k, maxiter=300)
    t3 = time.perf_counter()
    vh, _ = eigsh(L, k=k, maxiter=300)
    t4 = time.perf_counter()
    err = np.abs(np.sort(vals) - np.sort(va)).max()
    print('nS=%5d  device Lanczos %7.2f ms (%d steps, converged=%s)   ARPACK + device operator %7.2f ms   ARPACK host matrix %7.2f ms   '
          'max |dlambda| = %.1e' % (nS, (t1 - t0) * 1e3, info['steps'], info['converged'], (t3 - t2) * 1e3, (t4 - t3) * 1e3, err))
    Ld.free()
