"""Cross-check of bench.py's CPU arm (VERDICT r1 "What's weak" 8): the UNMODIFIED reference
getDistanceCTF_local_Conj9combinedS2.op (shimmed import exactly as tests/golden/make_golden.py does) timed beside the
oracle port on the same PDs, in the build container (the reference does not travel to the GPU box, so bench.py's
reference arm runs the port there and quotes the ratio measured here).

    python scripts/ref_vs_port_cpu.py  > profiles/r02_reference_vs_port_cpu.txt
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden                                      # noqa: E402
from manifoldem_python_b200 import synthetic            # noqa: E402
from oracle import pd_distance as opd                   # noqa: E402

ref = make_golden.load_reference()
print('host: %d cores; reference = /root/reference (unmodified), port = oracle/pd_distance.py, rotate_impl="tile"' % os.cpu_count())
for nS, N in ((48, 128), (32, 256), (96, 128)):
    pd = synthetic.make_pd(nS, N, seed=5, snr=0.3)
    em = pd['em']
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        res = make_golden.run_reference_pd(ref, pd, N, tmp=tmp)
        t_ref = time.perf_counter() - t0
    tm = {}
    t0 = time.perf_counter()
    out = opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'], em['EkV'],
                          em['AmpContrast'], rotate_impl='tile', timings=tm)
    t_port = time.perf_counter() - t0
    err = np.abs(out['D'] - res['D']).max() / np.abs(res['D']).max()
    print('PD %3d x %d^2: reference op() %.2f s (%.1f ms per particle, incl. its Python mask loop and the pickle dump), '
          'port %.2f s -> reference / port = %.2f;  max |D_port - D_ref| / max D = %.1e'
          % (nS, N, t_ref, 1e3 * t_ref / nS, t_port, t_ref / t_port, err))
    print('      port stage split [s]: ' + ', '.join('%s %.2f' % kv for kv in tm.items()))
