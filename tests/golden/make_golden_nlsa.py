"""Golden vectors of the NLSA / psi-analysis stage by running the UNMODIFIED reference
(modules/NLSA.py:23-158 with get_wiener, svdRF, L2_distance, fit_1D_open_manifold_3D; modules/psiAnalysisParS2.py:45-170):

    python tests/golden/make_golden_nlsa.py        -> tests/golden/nlsa_nS80_N24.npz (+ _mask variant inside)

A structured synthetic PD (1-D conformational coordinate) goes through the reference's distance stage and
DMembeddingII; NLSA.op then runs for the first two diffusion coordinates, with msk2 = 1 and with a disc mask.
np.random is seeded before every DMembeddingII-containing call (its curve_fit start is np.random.rand)."""
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden                                           # noqa: E402

SEED = 4321


def main():
    from manifoldem_python_b200 import synthetic
    ref = make_golden.load_reference()
    import NLSA                                              # noqa: E402  (the reference's)
    import psiAnalysisParS2                                  # noqa: E402
    import myio                                              # noqa: E402
    p = ref['p']
    N, nS = 24, 80
    with tempfile.TemporaryDirectory() as tmp:
        pd = synthetic.make_pd(nS, N, seed=9, snr=3.0)
        rec = make_golden.run_reference_pd(ref, pd, N, tmp=tmp)
        D, imgAll, CTF = np.array(rec['D']), np.array(rec['imgAll']), np.array(rec['CTF']).reshape(nS, N, N)
        p.num_eigs = 15
        np.random.seed(SEED)
        lamb, psi, sigma, mu, logEps, logSumWij, popt, R2 = ref['dm'].op(D.copy(), nS, 3.0, 60000)
        posPath = np.arange(nS)
        out = dict(D=D, imgAll=imgAll, CTF=CTF, psi=psi, posPath=posPath, N=N, nS=nS, tau_true=pd['tau'])
        conOrderRange, psiTrunc, tune = 10, 5, 3.0
        ConOrder = nS // conOrderRange
        g = np.arange(N) - N / 2 + 0.5
        disc = ((g[:, None] ** 2 + g[None, :] ** 2) < (0.4 * N) ** 2)
        for tag, msk2 in (('m1', 1), ('md', disc)):
            for psinum in (0, 1):
                PosPsi1 = np.argsort(psi[:, psinum])
                DD = D[posPath][:, posPath][PosPsi1][:, PosPsi1]
                par = dict(num=nS, ConOrder=ConOrder, k=nS - ConOrder, tune=tune, nS=nS, save=False, psiTrunc=psiTrunc)
                np.random.seed(SEED + psinum)
                IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu2, tau = NLSA.op(par, DD.copy(), posPath, PosPsi1, imgAll, msk2, CTF,
                                                                             dict(outDir='', prD=0))
                pre = '%s_psi%d_' % (tag, psinum)
                out.update({pre + 'IMGT': IMGT, pre + 'Topo_mean': Topo_mean, pre + 'psirec': psirec, pre + 'psiC1': psiC1,
                            pre + 'sdiag': sdiag, pre + 'VX': VX, pre + 'mu': mu2, pre + 'tau': tau, pre + 'PosPsi1': PosPsi1})
                print(tag, psinum, 'IMGT', IMGT.shape, 'sdiag', np.diag(sdiag)[:3], 'tau range', float(tau.min()), float(tau.max()))
        out['disc'] = disc
        out['params'] = np.array([conOrderRange, psiTrunc, tune, ConOrder])
        # ---- psiAnalysisParS2.op first pass (isFull = 0) on the same record, through the reference's own file protocol
        dist_file, psi_file, psi2_file = os.path.join(tmp, 'dist_0'), os.path.join(tmp, 'psi_0'), os.path.join(tmp, 'psi2_0')
        rec2 = dict(rec)
        rec2['CTF'] = CTF.reshape(nS, N * N)
        myio.fout1(dist_file, list(rec2.keys()), list(rec2.values()))
        myio.fout1(psi_file, ['psi', 'posPath'], [psi, posPath])
        p.psi2_prog = os.path.join(tmp, 'psi2_prog')
        os.makedirs(p.psi2_prog, exist_ok=True)
        p.tune, p.nClass, p.num_psis, p.numberofJobs = tune, 50, 2, 1
        np.random.seed(SEED)
        res = psiAnalysisParS2.op([dist_file, psi_file, psi2_file, os.path.join(tmp, 'EL_0'), np.array([0, 1]), np.array([1, 1]), 0],
                                  conOrderRange, 'traj', 0, psiTrunc)
        assert res == 'ok' and sorted(os.listdir(p.psi2_prog)) == ['0_0', '0_1']
        for psinum in (0, 1):
            d = myio.fin1('%s_psi_%d' % (psi2_file, psinum))
            out.update({'pa_psi%d_%s' % (psinum, k): np.asarray(d[k]) for k in ('IMG1', 'psirec', 'tau', 'psiC1', 'mu', 'VX', 'sdiag',
                                                                                  'Topo_mean', 'tauinds')})
    np.savez_compressed(os.path.join(HERE, 'nlsa_nS80_N24.npz'), **out)
    print('written', os.path.getsize(os.path.join(HERE, 'nlsa_nS80_N24.npz')), 'bytes')


if __name__ == '__main__':
    main()
