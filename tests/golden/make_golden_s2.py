"""Golden vectors of the S2 tessellation by running the UNMODIFIED reference
(modules/S2tessellation.py, modules/distribute3Sphere.py) in the build container:

    python tests/golden/make_golden_s2.py          -> tests/golden/s2_tessellation.npz

Same shims as make_golden.py (plotting stubs only; sklearn's NearestNeighbors is installed here)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden                                           # noqa: E402

make_golden.load_reference()
import S2tessellation                                        # noqa: E402  (the reference's)
import distribute3Sphere                                     # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'tests', 'tools'))
from s2_inputs import CASES, quats                          # noqa: E402

out = {}
for tag, n, width, lo, hi, seed in CASES:
    q = quats(n, seed)
    CG1, CG, nG, S2, S20_th, S20, NC = S2tessellation.op(q, width, lo, False, hi)
    IND = np.full(n, -1, dtype=np.int32)
    for i, a in enumerate(CG1):
        IND[a] = i
    assert (IND >= 0).all()
    out.update({tag + '_q_head': q[:, :32], tag + '_q_sum': q.sum(1), tag + '_args': np.array([width, lo, hi]), tag + '_nG': np.int64(nG),
                tag + '_IND': IND.astype(np.int16 if nG < 32768 else np.int32), tag + '_NC': NC,
                tag + '_S20': S20, tag + '_S20_th': S20_th, tag + '_S2_head': S2[:, :64],
                tag + '_CG_len': np.array([len(a) for a in CG]), tag + '_CG_first': np.array([a[0] for a in CG]),
                tag + '_CG_last': np.array([a[-1] for a in CG]),
                tag + '_CG_sum': np.array([int(np.sum(a)) for a in CG])})
    print(tag, 'nG', nG, 'PDs kept', len(CG), 'occupancy', min(map(len, CG)) if CG else None, max(map(len, CG)) if CG else None,
          'NC len', len(NC))
for K in (7, 100, 1000):
    pts, it = distribute3Sphere.op(K)
    out['sphere_%d' % K] = pts
    out['sphere_%d_iter' % K] = np.int64(it)
np.savez_compressed(os.path.join(HERE, 's2_tessellation.npz'), **out)
print('written', os.path.getsize(os.path.join(HERE, 's2_tessellation.npz')), 'bytes')
