"""CPU: the oracle restatement against vectors produced by the reference itself
(tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests then
compare the CUDA path with the oracle."""
import os

import numpy as np
import pytest

from oracle import pd_distance as opd
from oracle import dm_embedding as odm
from oracle.project_mask import project_mask


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _run(g, **kw):
    N = int(g['N'])
    return opd.pd_distance(g['ind'], g['q'], g['df'], g['stack'], int(g['nStot']), N, float(g['em_pix_size']),
                           float(g['em_Cs']), float(g['em_EkV']), float(g['em_AmpContrast']), **kw)


_FIELDS = ['D', 'CTF', 'imgAll', 'PD', 'PDs', 'Psis', 'imgAvg', 'imgAvgFlip', 'imgAllFlip', 'imgLabels',
           'Dnom', 'Nom', 'imgAllIntensity']


def _check(out, g, rtol=1e-10):
    for k in _FIELDS:
        ref = g['ref_' + k]
        got = np.asarray(out[k])
        assert got.shape == ref.shape, k
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(got - ref).max() <= rtol * scale, (k, np.abs(got - ref).max() / scale)


@pytest.mark.parametrize('name', ['pd_spider_N32.npz', 'pd_spider_N25_lownoise.npz'])
def test_pd_spider(golden_dir, name):
    g = _load(golden_dir, name)
    _check(_run(g), g)


def test_pd_direct_and_periodic(golden_dir):
    g = _load(golden_dir, 'pd_spider_N32.npz')
    out = _run(g, direct=True)
    assert np.abs(out['D'] - g['ref_D_direct']).max() <= 1e-12 * g['ref_D'].max()
    # the closed form the CUDA kernels implement (periodic cubic B-spline) == the reference's tile+rotate
    out = _run(g, rotate_impl='periodic')
    _check(out, g, rtol=1e-9)


def test_pd_relion(golden_dir):
    g = _load(golden_dir, 'pd_relion_N24.npz')
    N = int(g['N'])
    stack3d = g['stack'].reshape(-1, N, N)
    out = _run(g, stack=None) if False else opd.pd_distance(
        g['ind'], g['q'], g['df'], stack3d, int(g['nStot']), N, float(g['em_pix_size']), float(g['em_Cs']),
        float(g['em_EkV']), float(g['em_AmpContrast']), relion=True, sh=(g['shx'], g['shy']))
    _check(out, g)


def test_pd_volmask(golden_dir):
    g = _load(golden_dir, 'pd_volmask_N24.npz')
    PD = g['ref_PD']
    msk2 = project_mask(g['mask3d'], PD)
    assert msk2.dtype == bool and np.array_equal(msk2, g['ref_msk2'])
    assert 0 < msk2.sum() < msk2.size
    _check(_run(g, msk2=msk2), g)


def test_helpers(golden_dir):
    h = _load(golden_dir, 'helpers.npz')
    for impl, tol in (('tile', 1e-14), ('periodic', 1e-9)):   # N=20: 0.268^(0.79 N) edge term
        rot = np.stack([opd.rotatefill(h['imgs'][i], h['angs'][i], impl) for i in range(3)])
        assert np.abs(rot - h['rot']).max() < tol
    assert np.array_equal(opd.annular_mask(0, 10.0, 20, 20), h['msk'])
    assert np.array_equal(opd.annular_mask(0, 12.5, 25, 25), h['msk_odd'])
    assert np.array_equal(opd.create_grid(20), h['Q'])
    ctf = opd.ctemh_cryo_frank(h['Q'] / (2 * 1.255), 2.26, 21234.5, 300.0, np.inf, 0.1)
    assert np.abs(ctf - h['ctf']).max() < 1e-13
    eul = np.array([opd.q2spider(h['qs'][:, i]) for i in range(5)])
    assert np.abs(eul - h['eul']).max() < 1e-12


@pytest.mark.parametrize('k', [72, 20])
def test_dm_embedding(golden_dir, k):
    g = _load(golden_dir, 'dm_nS72.npz')
    D = g['D'].copy()
    rng = np.random.RandomState(1234)
    a0 = rng.rand(4, 1) - .5
    assert np.array_equal(a0, g[f'k{k}_a0'])
    lamb, psi, sigma, mu, logEps, logSumWij, popt, R2 = odm.dm_embedding(D, k, 3.0, a0=a0, rng=rng)
    assert np.isneginf(D[0, 0])                                   # mutated like the reference
    assert np.allclose(logSumWij, g[f'k{k}_logSumWij'], rtol=1e-12, atol=1e-12)
    assert np.allclose(popt, g[f'k{k}_popt'], rtol=1e-6)
    assert abs(sigma - float(g[f'k{k}_sigma'])) <= 1e-8 * sigma
    assert np.allclose(lamb, g[f'k{k}_lamb'], rtol=1e-8, atol=1e-10)
    assert np.allclose(mu, g[f'k{k}_mu'], rtol=1e-6, atol=1e-12)
    ref_psi = g[f'k{k}_psi']
    for j in range(4):                                            # leading eigenvectors, sign-free
        c = abs(np.corrcoef(psi[:, j], ref_psi[:, j])[0, 1])
        assert c > 1 - 1e-8, (j, c)


# ---- NLSA / psi analysis: oracle vs the unmodified reference's outputs -------------------------------
NLSA_SEED = 4321


def _cols_match(a, b, tol):
    """Columns equal up to a sign each (eigenvectors: ARPACK / LAPACK leave the sign free)."""
    scale = np.abs(b).max()
    for j in range(b.shape[1]):
        d = min(np.abs(a[:, j] - b[:, j]).max(), np.abs(a[:, j] + b[:, j]).max())
        assert d <= tol * scale, (j, d, scale)


def nlsa_outputs_match(out, g, pre, tol_lin, tol_eig):
    """NLSA.op outputs against the reference's: IMGT and sdiag are sign-free quantities; Topo_mean, psiC1, VX^T and psirec
    are eigen/singular vectors (sign per column free); tau is defined up to tau <-> 1 - tau (sign of psirec[:, 0])."""
    IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau = out
    assert np.allclose(IMGT, g[pre + 'IMGT'], rtol=0, atol=tol_lin)
    assert np.allclose(sdiag, g[pre + 'sdiag'], rtol=10 * tol_lin, atol=1e-12)
    _cols_match(Topo_mean, g[pre + 'Topo_mean'], 10 * tol_lin)
    _cols_match(psiC1[:, :6], g[pre + 'psiC1'][:, :6], tol_eig)
    _cols_match(VX.T, g[pre + 'VX'].T, tol_eig)
    _cols_match(psirec[:, :4], g[pre + 'psirec'][:, :4], 10 * tol_eig)
    assert np.allclose(mu, g[pre + 'mu'], rtol=10 * tol_eig, atol=1e-12)
    rt = g[pre + 'tau']
    assert min(np.abs(tau - rt).max(), np.abs(tau - (1 - rt)).max()) <= 1e-5


def _nlsa_case(g, tag, psinum):
    from oracle import nlsa as onl
    nS, N = int(g['nS']), int(g['N'])
    conOrderRange, psiTrunc, tune, ConOrder = g['params']
    ConOrder, psiTrunc = int(ConOrder), int(psiTrunc)
    posPath = g['posPath']
    PosPsi1 = np.argsort(g['psi'][:, psinum])
    assert np.array_equal(PosPsi1, g['%s_psi%d_PosPsi1' % (tag, psinum)])
    DD = g['D'][posPath][:, posPath][PosPsi1][:, PosPsi1]
    par = dict(num=nS, ConOrder=ConOrder, k=nS - ConOrder, tune=float(tune), nS=nS, save=False, psiTrunc=psiTrunc)
    msk2 = 1 if tag == 'm1' else g['disc']
    np.random.seed(NLSA_SEED + psinum)
    return onl.nlsa(par, DD, posPath, PosPsi1, g['imgAll'], msk2, g['CTF']), '%s_psi%d_' % (tag, psinum)


@pytest.mark.parametrize('tag,psinum', [('m1', 0), ('m1', 1), ('md', 0), ('md', 1)])
def test_nlsa_oracle_vs_reference(golden_dir, tag, psinum):
    g = _load(golden_dir, 'nlsa_nS80_N24.npz')
    (IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau), pre = _nlsa_case(g, tag, psinum)
    nlsa_outputs_match((IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau), g, pre, 1e-9, 1e-7)


def test_psi_analysis_class_representatives_vs_reference(golden_dir):
    from oracle import nlsa as onl
    g = _load(golden_dir, 'nlsa_nS80_N24.npz')
    for psinum in (0, 1):
        pre = 'm1_psi%d_' % psinum
        # the psiAnalysisParS2 run drew other random numbers than the direct NLSA.op run (tau may come out as 1 - tau):
        # feed its own (already rescaled: idempotent) tau and the sign-free IMGT of the direct run
        IMG1, tau, tauinds = onl.class_representatives(g[pre + 'IMGT'], g['pa_psi%d_tau' % psinum], 50)
        assert np.array_equal(tauinds, g['pa_psi%d_tauinds' % psinum]) and np.array_equal(tau, g['pa_psi%d_tau' % psinum])
        # sigma of the Ferguson fit depends on curve_fit's random start at the 1e-6 level, and IMGT with it
        assert np.allclose(IMG1, g['pa_psi%d_IMG1' % psinum], rtol=0, atol=1e-4)


from _box_golden import _box_case, check_box_outputs   # noqa: E402


@pytest.mark.parametrize('N', [128, 256, 320])
def test_pd_at_baseline_box_sizes(golden_dir, N):
    """The oracle (reference's tile + rotate, and the periodic closed form) at the box sizes BASELINE names."""
    g, pd = _box_case(golden_dir, N)
    em = pd['em']
    for impl in ('tile', 'periodic'):
        out = opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'], em['EkV'],
                              em['AmpContrast'], rotate_impl=impl)
        check_box_outputs(out, g, d_rtol=1e-8, img_tol=2e-7)
