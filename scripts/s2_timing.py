"""classS2 (S2tessellation.py:59-63) for n directions against nG bin centres: the C-ABI call (host buffers in, indices
out) against the reference's own ball-tree query on the host cores.   python scripts/s2_timing.py [n] [nG]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import S2tessellation as S2t   # noqa: E402
from manifoldem_python_b200 import _lib                    # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nG = int(sys.argv[2]) if len(sys.argv) > 2 else 4071
rng = np.random.default_rng(0)
X, _ = S2t.sphere_points(nG)
Q = rng.standard_normal((n, 3))
Q /= np.linalg.norm(Q, axis=1, keepdims=True)
ctx = S2t._ctx()
for r in range(3):
    t0 = time.time()
    IND, NC = S2t.classS2(X, Q, ctx)
    t_gpu = time.time() - t0
launches0 = ctx.launches()
print('mem_s2_assign_host: %d directions x %d centres in %.2f ms (host buffers in and out, %.1f G distance evaluations/s)'
      % (n, nG, t_gpu * 1e3, n * nG / t_gpu / 1e9))
try:
    from sklearn.neighbors import NearestNeighbors
    t0 = time.time()
    nbrs = NearestNeighbors(n_neighbors=1, algorithm='ball_tree').fit(X)
    _, ref = nbrs.kneighbors(Q)
    t_cpu = time.time() - t0
    same = np.array_equal(ref[:, 0], IND[:, 0])
    print('reference ball tree on the host (%d cores): %.2f s -> %.0fx; identical indices: %s (%d differ)'
          % (os.cpu_count(), t_cpu, t_cpu / t_gpu, same, int((ref[:, 0] != IND[:, 0]).sum())))
except ImportError:
    print('sklearn not installed: no host timing')
