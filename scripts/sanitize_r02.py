"""Small shapes through every kernel added in round 2 — device Lanczos, the NLSA stage (ConD, Wiener supervector sums, Gram,
projection, reconstruction, L2), the manifold-fit kernel, S2 pairwise distances, the batched distance stage with its grouped
tcgen05 launch, the new rotation kernels — for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_r02.py
    compute-sanitizer --tool racecheck python scripts/sanitize_r02.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import DMembeddingII, FindCCGraph, NLSA, _lib, fit_1D_open_manifold_3D, p, pd_stage, synthetic   # noqa: E402

p.init()
ctx = DMembeddingII._ctx()
rng = np.random.default_rng(0)
# ---- Lanczos
n = 90
X = rng.standard_normal((n, 3))
W = np.exp(-((X[:, None] - X[None]) ** 2).sum(-1))
d = np.sqrt(W.sum(1))
L = W / np.outer(d, d)
Ld = _lib.DeviceArray(ctx, (n, n), np.float64, L)
vals, vecs, info = DMembeddingII.eigsh_device(Ld, n, 8)
Ld.free()
assert info['converged'] and np.allclose(np.sort(vals), np.sort(np.linalg.eigvalsh(L))[-8:], atol=1e-9)
# ---- NLSA on the reference-generated golden inputs (one psi, disc mask)
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'nlsa_nS80_N24.npz'))
nS = int(g['nS'])
ConOrder, psiTrunc = int(g['params'][3]), int(g['params'][1])
par = dict(num=nS, ConOrder=ConOrder, k=nS - ConOrder, tune=3.0, nS=nS, save=False, psiTrunc=psiTrunc)
sel = np.argsort(g['psi'][:, 0])
np.random.seed(1)
out = NLSA.op(par, g['D'][sel][:, sel], g['posPath'], sel, g['imgAll'], g['disc'], g['CTF'], dict(prD=0))
assert np.isfinite(out[0]).all()
a, b, tau = fit_1D_open_manifold_3D.op(out[2])
assert np.isfinite(tau).all()
# ---- S2 pairwise
Xs = rng.standard_normal((3, 77))
Xs /= np.linalg.norm(Xs, axis=0)
dot, dist = FindCCGraph.CalcPairwiseDistS2(Xs)
assert np.allclose(np.diag(dist), 0)
# ---- batched distance stage (grouped tcgen05 launch), N = 64: three PDs of one stack
big = synthetic.make_pd_fast(150, 64, seed=4, snr=0.5)
perm = rng.permutation(150)
jobs = [(big['ind'][np.sort(perm[a:b])], big['q'][:, np.sort(perm[a:b])], big['df'][np.sort(perm[a:b])]) for a, b in ((0, 40), (40, 110), (110, 150))]
em = big['em']
res = pd_stage.run_pd_batch(jobs, big['stack'], big['nStot'], 64, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast'])
assert all(np.isfinite(r['D']).all() for r in res)
print('sanitize_r02: done')
