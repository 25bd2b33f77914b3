"""Golden vectors of FindCCGraph.CalcPairwiseDistS2 by running the UNMODIFIED reference (modules/FindCCGraph.py:227-273):

    python tests/golden/make_golden_s2_pairwise.py     -> tests/golden/s2_pairwise.npz

Inputs: the thresholded bin centres S20_th of the two tessellation cases (already in s2_tessellation.npz) — what
FindCCGraph.op passes at :296 — plus one index-pair call and a non-unit-vector case that exposes the broadcast of :267."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden                                           # noqa: E402

make_golden.load_reference()
import FindCCGraph                                           # noqa: E402  (the reference's)

g = np.load(os.path.join(HERE, 's2_tessellation.npz'))
out = {}
for tag in ('a', 'b'):
    X = g[tag + '_S20_th']
    dot, dist = FindCCGraph.CalcPairwiseDistS2(X)
    out[tag + '_dot'], out[tag + '_dist'] = dot, dist
    print(tag, X.shape, 'min nonzero dist', dist[dist > 0].min())
X = g['a_S20_th']
iu, iv = np.arange(0, 20), np.arange(5, 25)
out['idx_u'], out['idx_v'] = iu, iv
out['idx_dot'], out['idx_dist'] = FindCCGraph.CalcPairwiseDistS2(X, iu, iv)
rng = np.random.default_rng(5)
Y = rng.standard_normal((3, 40)) * rng.uniform(0.5, 2.0, 40)
out['nonunit_X'] = Y
out['nonunit_dot'], out['nonunit_dist'] = FindCCGraph.CalcPairwiseDistS2(Y, np.arange(0, 30), np.arange(10, 40))
np.savez_compressed(os.path.join(HERE, 's2_pairwise.npz'), **out)
print('written', os.path.getsize(os.path.join(HERE, 's2_pairwise.npz')), 'bytes')
