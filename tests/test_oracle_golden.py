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
