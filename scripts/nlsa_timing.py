"""NLSA stage timing on one PD (device path): PdState upload + per-psi analyse.   python scripts/nlsa_timing.py [nS] [N] [psis]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import NLSA, DMembeddingII, _lib, synthetic, p   # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
npsi = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p.init()
rng = np.random.default_rng(0)
t = np.sort(rng.uniform(0, 1, nS))
g = (np.arange(N) - N / 2) / N
yy, xx = np.meshgrid(g, g, indexing='ij')
imgAll = np.stack([np.exp(-((xx - 0.2 * (ti - 0.5)) ** 2 + yy ** 2) / 0.02) for ti in t]) + 0.3 * rng.standard_normal((nS, N, N))
CTF = np.stack([synthetic.ctf_2d(N, df) for df in rng.uniform(10000, 30000, nS)])
flat = imgAll.reshape(nS, -1)
sq = (flat ** 2).sum(1)
D = np.maximum(sq[:, None] + sq[None, :] - 2 * flat @ flat.T, 0)
np.random.seed(1)
psi = DMembeddingII.embed(D.copy(), nS, 3.0)[1]
ConOrder = nS // 50
par = dict(num=nS, ConOrder=ConOrder, k=nS - ConOrder, tune=3.0, nS=nS, save=False, psiTrunc=8)
t0 = time.perf_counter()
state = NLSA.PdState(D, imgAll, CTF)
t1 = time.perf_counter()
print('nS=%d N=%d ConOrder=%d   PdState (upload + %d forward transforms): %.1f ms' % (nS, N, ConOrder, nS, (t1 - t0) * 1e3))
for rep in range(2):
    for psinum in range(npsi):
        sel = np.argsort(psi[:, psinum])
        t2 = time.perf_counter()
        tm = {}
        out = NLSA.analyse(state, sel, sel, par, 1, keep_IMGT_on_device=True, timings=tm if rep else None)
        state.ctx.sync()
        t3 = time.perf_counter()
        out[0].free()
        print('rep %d psi %d: analyse %.1f ms   (reference: %d fft2/ifft2 pairs of %d^2 + a %d x %d x %d float64 Gram)'
              % (rep, psinum, (t3 - t2) * 1e3, ConOrder * (nS - ConOrder), N, nS - 2 * ConOrder, nS - 2 * ConOrder, N * N))
        if tm:
            print('    ', {k: round(v, 1) for k, v in tm.items()})
state.free()
