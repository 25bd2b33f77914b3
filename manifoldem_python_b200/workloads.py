"""Workloads of BASELINE.json's configs as inputs for bench.py, scripts/ and the tests.

config 1 — the repo demo: orientations, defoci and PD membership of demo/RyR1GCs_clustRem.star through the reference's
reader and tessellation (53 projection directions, 117..450 particles; fixture tests/golden/demo_config1.npz written by
tests/golden/make_demo_config1.py in the build container).  The demo's particle images are not in the repository:
they are replaced by synthetic images at a stated box size."""
import os

import numpy as np

from . import pd_stage

_HERE = os.path.dirname(os.path.abspath(__file__))
DEMO_FIXTURE = os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'demo_config1.npz')


def demo_config1():
    """-> list of dict(ind, q, df, nStot, psi_deg, psi_p, flip, PD) for the 53 PDs of the demo, plus em dict."""
    g = np.load(DEMO_FIXTURE)
    off = g['offsets']
    nStot = int(g['nStot'])
    kV, Cs, ampC, pix = (float(x) for x in g['em'])
    pds = []
    for p in range(len(off) - 1):
        sl = slice(off[p], off[p + 1])
        ind, q, df = g['ind'][sl].astype(np.int64), np.ascontiguousarray(g['q'][:, sl]), np.ascontiguousarray(g['df'][sl])
        PDs, PD, psi_p, Psi, s, c = pd_stage.host_angles(q)
        pds.append(dict(ind=ind, q=q, df=df, nStot=nStot, psi_deg=np.ascontiguousarray(-(180 / np.pi) * Psi),
                        psi_p=float(psi_p), flip=(ind >= nStot / 2).astype(np.uint8), PD=PD))
    return pds, dict(EkV=kV, Cs=Cs, AmpContrast=ampC, pix_size=pix)
