"""GetDistancesS2.op on a box with several GPUs: p.ncpu caps the worker count, one spawned process per GPU,
one job queue for the box (distance stage) / static LPT partition (embedding, psi analysis), markers as the only gather.  python scripts/multigpu_driver_check.py [n_gpus]"""
import os
import pickle
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from manifoldem_python_b200 import GetDistancesS2, myio, p, synthetic
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    N = 64
    sizes = (150, 90, 200, 120, 60, 180, 75)
    pds = [synthetic.make_pd(n, N, seed=30 + i, snr=0.5) for i, n in enumerate(sizes)]
    n_half = sum(pd['nStot'] // 2 for pd in pds)
    stack = np.concatenate([pd['stack'] for pd in pds])
    q = np.zeros((4, 2 * n_half))
    df = np.zeros(2 * n_half)
    CG, off = [], 0
    for pd in pds:
        h = pd['nStot'] // 2
        base = np.where(pd['ind'] >= h, pd['ind'] - h, pd['ind']) + off
        ind = np.where(pd['ind'] >= h, base + n_half, base)
        q[:, ind] = pd['q']
        df[ind] = pd['df']
        CG.append(ind)
        off += h
    with tempfile.TemporaryDirectory() as tmp:
        p.init()
        p.user_dir, p.proj_name = tmp, 'multi'
        p.create_dir()
        em = pds[0]['em']
        p.pix_size, p.Cs, p.EkV, p.AmpContrast = em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
        p.relion_data, p.ncpu, p.num_part = False, n_gpus, n_half
        p.img_stack_file = os.path.join(tmp, 'stack.dat')
        stack.tofile(p.img_stack_file)
        p.numberofJobs = len(CG)
        myio.fout1(p.tess_file, ['CG', 'df', 'q', 'sh'], [CG, df, q, (np.zeros(n_half), np.zeros(n_half))])

        class Sig:
            vals = []

            def emit(self, v):
                self.vals.append(v)
        t0 = time.time()
        GetDistancesS2.op(Sig())
        dt = time.time() - t0
        assert sorted(int(f) for f in os.listdir(p.dist_prog)) == list(range(len(CG)))
        assert Sig.vals[-1] == 100
        for prD in range(len(CG)):
            with open('{}prD_{}'.format(p.dist_file, prD), 'rb') as f:
                rec = pickle.load(f)
            nS = len(CG[prD])
            assert rec['D'].shape == (nS, nS) and np.array_equal(rec['D'], rec['D'].T) and np.isfinite(rec['D']).all()
        print('GetDistancesS2.op with p.ncpu=%d on %d visible GPUs: %d PDs in %.1f s (incl. worker spawn + CUDA init), '
              'all markers present, progress %s' % (n_gpus, GetDistancesS2._n_gpus(), len(CG), dt, Sig.vals[-3:]))
        # the next two stages through their drivers, both sharded over the GPUs
        from manifoldem_python_b200 import manifoldAnalysis, psiAnalysis
        p.tune, p.rad = 3.0, 5.0
        Sig.vals = []
        t0 = time.time()
        manifoldAnalysis.op(Sig())
        dt = time.time() - t0
        assert sorted(int(f) for f in os.listdir(p.psi_prog)) == list(range(len(CG))) and Sig.vals[-1] == 100
        print('manifoldAnalysis.op on %d visible GPUs: %d PDs in %.1f s, all markers present' % (GetDistancesS2._n_gpus(), len(CG), dt))
        out = os.path.join(tmp, 'outputs_multi')
        p.psi2_dir, p.EL_dir = os.path.join(out, 'psi_analysis/'), os.path.join(out, 'ELConc10/')
        p.psi2_prog, p.EL_prog = os.path.join(p.psi2_dir, 'progress/'), os.path.join(p.EL_dir, 'progress/')
        for d in (p.psi2_prog, p.EL_prog):
            os.makedirs(d)
        p.psi2_file, p.EL_file = os.path.join(p.psi2_dir, 'S2_'), os.path.join(p.EL_dir, 'S2_')
        p.num_psis, p.conOrderRange, p.num_psiTrunc, p.nClass, p.trajName, p.tune = 2, 10, 5, 50, '1', 3
        Sig.vals = []
        t0 = time.time()
        psiAnalysis.op(Sig())
        dt = time.time() - t0
        done = sorted(os.listdir(p.psi2_prog))
        assert len(done) == 2 * len(CG) and Sig.vals[-1] == 100, (done, Sig.vals[-3:])
        print('psiAnalysis.op on %d visible GPUs: %d (PD, psi) pairs in %.1f s, all markers present' % (GetDistancesS2._n_gpus(), len(done), dt))


if __name__ == '__main__':
    main()
