"""The stage as a user of the reference runs it — GetDistancesS2.op (drop-in) from a SPIDER stack on disk to the per-PD
records on disk, then manifoldTrimmingAuto.op on every record — with the reference's record layout, with the
'sidecar' layout holding every array and with its default (CTF field virtual, imgAllFlip skipped).  Wall clock per PD; white-noise particles (timing only).
    python scripts/dropin_e2e.py [n_pd] [nS] [N] [dir]           default 4 PDs x 2000 x 256^2
"""
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import GetDistancesS2, manifoldTrimmingAuto, myio, p, synthetic   # noqa: E402

n_pd = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nS = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
work = sys.argv[4] if len(sys.argv) > 4 else '/tmp/dropin_e2e'
trim = '--trim' in sys.argv

rng = np.random.default_rng(0)
n_half = n_pd * nS
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
stack_file = os.path.join(work, 'stack.dat')
with open(stack_file, 'wb') as f:
    for _ in range(n_pd):
        rng.standard_normal((nS, N * N), dtype=np.float32).tofile(f)
q = np.zeros((4, 2 * n_half))
q[:, :n_half] = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(n_half), 1.1 + 0.03 * rng.standard_normal(n_half),
                                        rng.uniform(0, 2 * np.pi, n_half))
q[:, n_half:] = q[:, :n_half]
df = np.tile(rng.uniform(10000.0, 30000.0, n_half), 2)
CG = [np.arange(i * nS, (i + 1) * nS) for i in range(n_pd)]

layouts = [a.split('=')[1].split(',') for a in sys.argv if a.startswith('--layouts=')]
for layout in (layouts[0] if layouts else ('pickle', 'sidecar-full', 'sidecar')):
    p.init()
    p.user_dir, p.proj_name = work, 'run_' + layout
    p.create_dir()
    p.pix_size, p.Cs, p.EkV, p.AmpContrast = 1.255, 2.26, 300.0, 0.1
    p.relion_data, p.ncpu, p.num_part, p.numberofJobs = False, 1, n_half, n_pd
    p.img_stack_file = stack_file
    p.record_layout = layout.split('-')[0]
    p.record_virtual_ctf = (layout == 'sidecar')           # 'sidecar-full': every array of the reference's record stored
    p.record_skip = ('imgAllFlip',) if layout == 'sidecar' else ()
    myio.fout1(p.tess_file, ['CG', 'df', 'q', 'sh'], [CG, df, q, (np.zeros(n_half), np.zeros(n_half))], layout='pickle')
    GetDistancesS2.op()                                  # warm-up pass over every PD (plans, workspaces, page cache)
    for f in os.listdir(p.dist_prog):
        os.remove(os.path.join(p.dist_prog, f))
    t0 = time.time()
    GetDistancesS2.op()
    t_dist = (time.time() - t0) / n_pd
    size = sum(os.path.getsize(os.path.join(os.path.dirname(p.dist_file), f))
               for f in os.listdir(os.path.dirname(p.dist_file)) if 'prD_' in f) / n_pd
    t0 = time.time()
    for prD in range(n_pd):                              # what the embedding stage does first (manifoldTrimmingAuto.py:44-46)
        data = myio.fin1('{}prD_{}'.format(p.dist_file, prD))
        D, ind = data['D'], data['ind']
        assert D.shape == (nS, nS) and D.dtype == np.float64
    t_read = (time.time() - t0) / n_pd
    t0 = time.time()
    ctf = myio.fin1('{}prD_{}'.format(p.dist_file, 0))['CTF']      # what psiAnalysisParS2.py:61-65 reads besides imgAll
    assert ctf.shape == (nS, N * N) and ctf.dtype == np.float64
    t_ctf = time.time() - t0
    del ctf
    msg = '%-12s distance stage %.3f s / PD (record %.2f GB), consumer reads D + ind in %.3f s / PD, the CTF field in %.3f s' % (layout, t_dist, size / 1e9, t_read, t_ctf)
    if trim:
        t0 = time.time()
        for prD in range(n_pd):
            manifoldTrimmingAuto.op(['{}prD_{}'.format(p.dist_file, prD), '{}prD_{}'.format(p.psi_file, prD),
                                     os.path.join(p.psi_dir, 'eig_prD_%d' % prD), prD], 0, 3.0, 5.0, False, dict(outputFile='', Is=True))
        msg += ', manifoldTrimmingAuto.op %.3f s / PD' % ((time.time() - t0) / n_pd)
    print(msg, flush=True)
    shutil.rmtree(os.path.join(work, 'outputs_run_' + layout), ignore_errors=True)
shutil.rmtree(work, ignore_errors=True)
