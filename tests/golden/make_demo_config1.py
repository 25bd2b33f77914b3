"""BASELINE config 1 bookkeeping as a fixture (the demo star itself does not travel to the GPU box):

    python tests/golden/make_demo_config1.py      -> tests/golden/demo_config1.npz

Reads /root/reference/demo/RyR1GCs_clustRem.star with the UNMODIFIED reference reader (read_alignfile.get_from_relion,
util.augment, Data.py's defocus averaging) and tessellates with the reference's S2tessellation.op at the manual's
settings (aperture index 4 at 5 A / 360 A, thresholds 100 / 2000): 53 projection directions, 117..450 particles.
Stored: for every PD the member indices into the augmented set, their quaternions and defoci, and nStot — what
GetDistancesS2.divide (GetDistancesS2.py:37-47) hands to the per-PD worker.  The particle images of the demo
(.mrcs) are not in the repository; bench.py / tests pair this bookkeeping with synthetic images."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden                                           # noqa: E402

import builtins                                              # noqa: E402

make_golden.load_reference()
_open = builtins.open                                        # star.py opens with mode 'rU' (gone in Python >= 3.11)
builtins.open = lambda f, mode='r', *a, **k: _open(f, mode.replace('U', ''), *a, **k)
import read_alignfile                                        # noqa: E402  (the reference's)
import util                                                  # noqa: E402
import S2tessellation                                        # noqa: E402

STAR = '/root/reference/demo/RyR1GCs_clustRem.star'
sh, q, U, V = read_alignfile.get_from_relion(STAR, flip=True)
df = (U + V) / 2.0                                           # Data.py:90
q = util.augment(q)                                          # Data.py:92 (conjugates appended)
df = np.concatenate((df, df))                                # Data.py:93
CG1, CG, nG, S2, S20_th, S20, NC = S2tessellation.op(q, 4 * 5.0 / 360, 100, False, 2000)
occ = np.array([len(a) for a in CG])
assert len(CG) == 53 and occ.min() == 117 and occ.max() == 450 and int((occ ** 2).sum()) == 3130240
ind = np.concatenate([np.asarray(a, dtype=np.int32) for a in CG])
off = np.concatenate(([0], np.cumsum(occ))).astype(np.int32)
np.savez_compressed(os.path.join(HERE, 'demo_config1.npz'), ind=ind, offsets=off, q=q[:, ind].astype(np.float64),
                    df=df[ind].astype(np.float64), nStot=np.int64(q.shape[1]), nG=np.int64(nG),
                    em=np.array([300.0, 2.26, 0.1, 1.255]))     # kV, Cs, ampC (star optics group), pixel size (manual)
print('PDs', len(CG), 'particles', int(occ.sum()), 'nStot', q.shape[1], 'bytes',
      os.path.getsize(os.path.join(HERE, 'demo_config1.npz')))
