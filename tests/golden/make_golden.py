"""Generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so the pins
are its own outputs on seeded synthetic PDs.  The shims below exist because of
missing packages / newer NumPy-Python in this image, not because of the
algorithm: stub modules for plotting / MRC / MPI,
np.complex and np.Inf aliases, modules/ and modules/CC/ on sys.path.
Outputs: tests/golden/*.npz (committed; small).
"""
import os
import sys
import tempfile
import types
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/modules'
sys.path.insert(0, ROOT)


class _FakeMrc:
    """Stands in for mrcfile.mmap()/open(): only .data and is_image_stack() are used
    (getDistanceCTF_local_Conj9combinedS2.py:260-262, :305-306)."""
    registry = {}

    def __init__(self, name, *a, **k):
        self.data = _FakeMrc.registry[name]

    def is_image_stack(self):
        return True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mrcfile", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpi4py", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["mrcfile"].mmap = _FakeMrc
    sys.modules["mrcfile"].open = _FakeMrc
    np.complex, np.Inf = complex, np.inf
    sys.path[:0] = [REF]
    sys.path.append(os.path.join(REF, 'CC'))
    import p
    import getDistanceCTF_local_Conj9combinedS2 as gd
    import DMembeddingII
    import rotatefill
    import annularMask
    import ctemh_cryoFrank
    import q2Spider
    return dict(p=p, gd=gd, dm=DMembeddingII, rotatefill=rotatefill, annularMask=annularMask,
                ctemh=ctemh_cryoFrank, q2Spider=q2Spider)


def run_reference_pd(ref, pd, N, relion=False, stack3d=None, sh=None, parallel=False, mask3d=None, tmp=None):
    p, gd = ref['p'], ref['gd']
    p.init()
    em = pd['em']
    p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast = N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
    p.mask_vol_file, p.ncpu, p.relion_data = '', 1, relion
    p.dist_prog = os.path.join(tmp, 'progress')
    os.makedirs(p.dist_prog, exist_ok=True)
    if mask3d is not None:
        p.mask_vol_file = 'mask3d'
        _FakeMrc.registry['mask3d'] = mask3d
    if relion:
        img_file = 'relion_stack'
        _FakeMrc.registry[img_file] = stack3d
    else:
        img_file = os.path.join(tmp, 'stack.dat')
        pd['stack'].astype(np.float32).tofile(img_file)
    out = os.path.join(tmp, 'IMGs_prD_0')
    opts = dict(verbose=False, avgOnly=False, visual=False, parallel=parallel, relion_data=relion, thres=2000)
    gd.op([pd['ind'], pd['q'], pd['df'], out, 0], dict(type='Butter', Qc=0.5, N=8), img_file,
          sh if sh is not None else pd['sh'], pd['nStot'], opts)
    assert os.path.exists(os.path.join(p.dist_prog, '0'))
    with open(out, 'rb') as f:
        return pickle.load(f)


_KEYS = ['D', 'CTF', 'imgAll', 'PD', 'PDs', 'Psis', 'imgAvg', 'imgAvgFlip', 'imgAllFlip', 'imgLabels',
         'Dnom', 'Nom', 'imgAllIntensity']


def pack(pd, N, res, extra=None):
    d = dict(stack=pd['stack'], ind=pd['ind'], q=pd['q'], df=pd['df'], shx=pd['sh'][0], shy=pd['sh'][1],
             nStot=pd['nStot'], N=N, **{'em_' + k: v for k, v in pd['em'].items()})
    for k in _KEYS:
        d['ref_' + k] = np.asarray(res[k])
    d['ref_msk2'] = np.asarray(res['msk2'])
    if extra:
        d.update(extra)
    return d


def main():
    from manifoldem_python_b200 import synthetic
    ref = load_reference()
    with tempfile.TemporaryDirectory() as tmp:
        # ---- case A: SPIDER path, matrix D and definitional D ---------------------------------
        N, nS = 32, 16
        pd = synthetic.make_pd(nS, N, seed=0, snr=0.1)
        resA = run_reference_pd(ref, pd, N, tmp=tmp)
        resAd = run_reference_pd(ref, pd, N, parallel=True, tmp=tmp)
        np.savez_compressed(os.path.join(HERE, 'pd_spider_N32.npz'),
                            **pack(pd, N, resA, dict(ref_D_direct=resAd['D'])))
        print('A  max|D-Ddirect|/max D =', np.abs(resA['D'] - resAd['D']).max() / resA['D'].max())

        # ---- case A2: odd box, low noise (worst cancellation) ---------------------------------
        N, nS = 25, 10
        pd = synthetic.make_pd(nS, N, seed=1, snr=10.0)
        res = run_reference_pd(ref, pd, N, tmp=tmp)
        np.savez_compressed(os.path.join(HERE, 'pd_spider_N25_lownoise.npz'), **pack(pd, N, res))

        # ---- case B: RELION path (cubic 'wrap' shift by (shy-0.5, shx-0.5)) ---------------------
        N, nS = 24, 12
        pd = synthetic.make_pd(nS, N, seed=2, snr=0.5)
        rng = np.random.default_rng(5)
        n_half = pd['nStot'] // 2
        sh = (rng.uniform(-3, 3, n_half), rng.uniform(-3, 3, n_half))
        # mrcfile's .data is (n, N, N) in picture orientation (no transpose on that path, :260-264)
        stack3d = pd['stack'].reshape(n_half, N, N).copy()
        res = run_reference_pd(ref, pd, N, relion=True, stack3d=stack3d, sh=sh, tmp=tmp)
        pd_b = dict(pd)
        pd_b['sh'] = sh
        np.savez_compressed(os.path.join(HERE, 'pd_relion_N24.npz'), **pack(pd_b, N, res))

        # ---- case C: volumetric mask (projectMask.op) ---------------------------------------------
        N, nS = 24, 10
        pd = synthetic.make_pd(nS, N, seed=3, snr=0.5)
        g = np.arange(N) - N / 2
        zz, yy, xx = np.meshgrid(g, g, g, indexing='ij')
        mask3d = ((xx / 9.0) ** 2 + (yy / 7.0) ** 2 + (zz / 5.0) ** 2 < 1).astype(np.float32)
        res = run_reference_pd(ref, pd, N, mask3d=mask3d, tmp=tmp)
        np.savez_compressed(os.path.join(HERE, 'pd_volmask_N24.npz'), **pack(pd, N, res, dict(mask3d=mask3d)))

        # ---- case D: diffusion-map front end on a structured PD -------------------------------
        N, nS = 32, 72
        pd = synthetic.make_pd(nS, N, seed=4, snr=2.0)
        res = run_reference_pd(ref, pd, N, tmp=tmp)
        D = res['D']
        out = dict(D=D, tau=pd['tau'])
        ref['p'].num_eigs = 15
        for k in (nS, 20):
            np.random.seed(1234)
            a0 = np.random.rand(4, 1) - .5            # same draw the reference makes first (DMembeddingII.py:142)
            np.random.seed(1234)
            lamb, psi, sigma, mu, logEps, logSumWij, popt, R2 = ref['dm'].op(D.copy(), k, 3.0, 60000)
            out.update({f'k{k}_lamb': lamb, f'k{k}_psi': psi, f'k{k}_sigma': sigma, f'k{k}_mu': mu,
                        f'k{k}_logSumWij': logSumWij, f'k{k}_popt': popt, f'k{k}_R2': R2, f'k{k}_a0': a0})
        out['logEps'] = logEps
        np.savez_compressed(os.path.join(HERE, 'dm_nS72.npz'), **out)

        # ---- micro known answers (helpers) ---------------------------------------------------------
        rng = np.random.default_rng(7)
        imgs = rng.standard_normal((3, 20, 20))
        angs = np.array([-37.3, 123.456, 301.0])
        rot = np.stack([ref['rotatefill'].op(imgs[i], angs[i]) for i in range(3)])
        msk = ref['annularMask'].op(0, 10.0, 20, 20)
        msk_odd = ref['annularMask'].op(0, 12.5, 25, 25)
        Q = ref['gd'].create_grid(20, 10.0)
        ctf = ref['ctemh'].op(Q / (2 * 1.255), [2.26, 21234.5, 300.0, np.inf, 0.1])
        qs = rng.standard_normal((4, 5))
        qs /= np.linalg.norm(qs, axis=0)
        eul = np.array([ref['q2Spider'].op(qs[:, i]) for i in range(5)])
        np.savez_compressed(os.path.join(HERE, 'helpers.npz'), imgs=imgs, angs=angs, rot=rot, msk=msk,
                            msk_odd=msk_odd, Q=Q, ctf=ctf, qs=qs, eul=eul)
    print('golden vectors written to', HERE)


if __name__ == '__main__':
    main()
