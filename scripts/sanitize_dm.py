"""Small shapes through every kernel added after the distance stage proper — kNN selection (resident D and split-K
partial tiles), graph, sorted-chunk Ferguson sweep, column sums / Laplacian, S2 assignment, stand-alone CTF — for
compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_dm.py
    compute-sanitizer --tool racecheck python scripts/sanitize_dm.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import DMembeddingII, S2tessellation, _lib, myio, p, pd_stage, synthetic   # noqa: E402

p.init()
lib = _lib.load()
ctx = DMembeddingII._ctx()
rng = np.random.default_rng(0)
for dtype, nS, k in ((np.float32, 300, 40), (np.float64, 257, 33), (np.float32, 64, 64)):
    D = rng.integers(0, 30, (nS, nS)).astype(dtype)
    D = np.maximum(D, D.T)
    for mode in (1, 2):
        _lib.check(lib.mem_knn_mode(mode))
        M, logEps, ls, idx, val = DMembeddingII.graph_and_sweep(D.astype(np.float64) if dtype == np.float64 else
                                                                _lib.DeviceArray(ctx, (nS, nS), np.float32, D), k)
        L = DMembeddingII.laplacian(M, nS, 3.0)
        M.free()
        assert np.isfinite(ls).all() and np.isfinite(L).all()
    _lib.check(lib.mem_knn_mode(0))
pd = synthetic.make_pd(150, 64, seed=3, snr=0.5)
em = pd['em']
for kw in (dict(), dict(split_k=3), dict(contraction=2)):
    res = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], 64, em['pix_size'], em['Cs'], em['EkV'],
                          em['AmpContrast'], fields=(), knn_k=20, float64=False, **kw)
    assert (res['knn_idx'][:, 0] == np.arange(150)).all()
X, _ = S2tessellation.sphere_points(314)
Q = rng.standard_normal((5000, 3))
IND, NC = S2tessellation.classS2(X, Q / np.linalg.norm(Q, axis=1, keepdims=True))
assert NC.sum() == 5000
ctf = myio._ctf_field(np.linspace(1e4, 3e4, 10), dict(N=32, pix_size=1.2, Cs=2.2, EkV=300.0, gaussEnv=np.inf,
                                                       AmpContrast=0.1, shape=(10, 32 * 32)))
assert np.isfinite(ctf).all()
print('sanitize_dm: done')
