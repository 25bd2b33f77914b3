"""S2 tessellation drop-in against vectors produced by the unmodified reference
(tests/golden/make_golden_s2.py -> tests/golden/s2_tessellation.npz).

CPU: the oracle's nearest-bin assignment and the drop-in's host logic (bin centres, thresholds, grouping; the device
call replaced by the oracle inside the test only).  GPU: the CUDA assignment through the C ABI — integer work, exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tools'))
from s2_inputs import CASES, quats                      # noqa: E402


@pytest.fixture(scope='module')
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, 's2_tessellation.npz'), allow_pickle=False)


def _check_case(S2t, g, tag, n, width, lo, hi, seed):
    q = quats(n, seed)
    assert np.array_equal(q[:, :32], g[tag + '_q_head']) and np.allclose(q.sum(1), g[tag + '_q_sum'], rtol=0, atol=1e-9)
    CG1, CG, nG, S2, S20_th, S20, NC = S2t.op(q, width, lo, False, hi)
    assert int(nG) == int(g[tag + '_nG']) and np.array_equal(S20, g[tag + '_S20'])
    assert np.array_equal(S2[:, :64], g[tag + '_S2_head']) and S2.shape == (3, n)
    IND = np.full(n, -1, dtype=np.int64)
    for i, a in enumerate(CG1):
        assert (np.diff(a) > 0).all()
        IND[a] = i
    assert np.array_equal(IND, g[tag + '_IND'].astype(np.int64))
    assert np.array_equal(NC, g[tag + '_NC']) and np.array_equal(S20_th, g[tag + '_S20_th'])
    assert np.array_equal([len(a) for a in CG], g[tag + '_CG_len'])
    assert np.array_equal([a[0] for a in CG], g[tag + '_CG_first'])
    assert np.array_equal([a[-1] for a in CG], g[tag + '_CG_last'])
    assert np.array_equal([int(np.sum(a)) for a in CG], g[tag + '_CG_sum'])


def test_bin_centres_match_the_reference(golden):
    from manifoldem_python_b200 import S2tessellation as S2t
    for K in (7, 100, 1000):
        pts, passes = S2t.sphere_points(K)
        assert np.array_equal(pts, golden['sphere_%d' % K]) and passes == int(golden['sphere_%d_iter' % K])


def test_oracle_assignment_matches_the_reference(golden):
    from oracle import s2_tessellation as os2
    from manifoldem_python_b200 import S2tessellation as S2t
    for tag, n, width, lo, hi, seed in CASES:
        S2 = S2t.get_S2(quats(n, seed))
        ind = os2.nearest_bin(golden[tag + '_S20'].T, S2.T)
        assert np.array_equal(ind, golden[tag + '_IND'].astype(np.int64)), tag


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_host_logic_with_the_oracle_assignment(golden, case, monkeypatch):
    from oracle import s2_tessellation as os2
    from manifoldem_python_b200 import S2tessellation as S2t

    def cpu_class(X, Q, ctx=None):
        ind = os2.nearest_bin(X, Q).reshape(-1, 1)
        return ind, np.bincount(ind[:, 0])
    monkeypatch.setattr(S2t, 'classS2', cpu_class)
    _check_case(S2t, golden, *case)


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_gpu_tessellation_matches_the_reference(golden, case):
    from manifoldem_python_b200 import S2tessellation as S2t
    _check_case(S2t, golden, *case)


@pytest.mark.gpu
def test_gpu_assignment_large_and_edge_cases():
    """10^6 directions against 4,071 centres equal the oracle's argmin; ties go to the smallest index; one centre,
    zero directions and bad shapes behave."""
    from oracle import s2_tessellation as os2
    from manifoldem_python_b200 import S2tessellation as S2t
    rng = np.random.default_rng(9)
    X, _ = S2t.sphere_points(4071)
    Q = rng.standard_normal((1000000, 3))
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    IND, NC = S2t.classS2(X, Q)
    assert IND.shape == (1000000, 1) and NC.sum() == 1000000
    sel = rng.integers(0, 1000000, 20000)
    assert np.array_equal(IND[sel, 0], os2.nearest_bin(X, Q[sel]))
    Xt = np.array([[1.0, 0, 0], [0, 1.0, 0], [1.0, 0, 0], [0, 0, 1.0]])
    IND, NC = S2t.classS2(Xt, np.array([[1.0, 0, 0], [0.6, 0.6, 0.0], [0, 0, -1.0]]))
    assert list(IND[:, 0]) == [0, 0, 0] and list(NC) == [3]
    IND, _ = S2t.classS2(Xt[:1], Q[:5])
    assert list(IND[:, 0]) == [0] * 5
    IND, NC = S2t.classS2(Xt, np.zeros((0, 3)))
    assert IND.shape == (0, 1) and len(NC) == 0
    with pytest.raises(ValueError):
        S2t.classS2(Xt[:, :2], Q[:5])


REF_STAR = '/root/reference/demo/RyR1GCs_clustRem.star'


@pytest.mark.skipif(not os.path.exists(REF_STAR), reason='reference demo star only exists in the build container')
def test_demo_star_gives_the_53_projection_directions(monkeypatch):
    """BASELINE config 1 bookkeeping: the orientations of the repo demo (read with the reference's own
    star reader), augmented and tessellated with the manual's settings — aperture index 4 at 5 A / 360 A => width
    4*5/360, thresholds 100 / 2000 — give 53 projection directions with 117..450 particles, sum nS^2 = 3,130,240.
    Host logic of the drop-in with the oracle's nearest-bin assignment; build container only."""
    import builtins
    import types
    for name in ('matplotlib', 'matplotlib.pyplot', 'mrcfile', 'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpi4py'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    ref_mod = '/root/reference/modules'
    monkeypatch.syspath_prepend(ref_mod)
    _open = builtins.open
    monkeypatch.setattr(builtins, 'open', lambda f, mode='r', *a, **k: _open(f, mode.replace('U', ''), *a, **k))
    before = set(sys.modules)
    try:
        import read_alignfile
        import util
        sh, q, U, V = read_alignfile.get_from_relion(REF_STAR, flip=True)
        q = util.augment(q)
    finally:                                                # leave no reference module behind for other tests
        for name in set(sys.modules) - before:
            if getattr(sys.modules[name], '__file__', None) and str(sys.modules[name].__file__).startswith('/root/reference'):
                del sys.modules[name]
    from oracle import s2_tessellation as os2
    from manifoldem_python_b200 import S2tessellation as S2t

    def cpu_class(X, Q, ctx=None):
        ind = os2.nearest_bin(X, Q).reshape(-1, 1)
        return ind, np.bincount(ind[:, 0])
    monkeypatch.setattr(S2t, 'classS2', cpu_class)
    CG1, CG, nG, S2, S20_th, S20, NC = S2t.op(q, 4 * 5.0 / 360, 100, False, 2000)
    occ = np.array([len(a) for a in CG])
    assert int(nG) == 4071 and len(CG) == 53
    assert occ.min() == 117 and occ.max() == 450 and int(np.median(occ)) == 206 and int((occ ** 2).sum()) == 3130240


# ---- FindCCGraph.CalcPairwiseDistS2 (modules/FindCCGraph.py:227-273) -------------------------------------------------
@pytest.fixture(scope='module')
def golden_pw(golden_dir):
    return np.load(os.path.join(golden_dir, 's2_pairwise.npz'), allow_pickle=False)


def _pairwise_cases(golden, golden_pw):
    yield 'a', golden['a_S20_th'], None, None, golden_pw['a_dot'], golden_pw['a_dist']
    yield 'b', golden['b_S20_th'], None, None, golden_pw['b_dot'], golden_pw['b_dist']
    yield 'idx', golden['a_S20_th'], golden_pw['idx_u'], golden_pw['idx_v'], golden_pw['idx_dot'], golden_pw['idx_dist']
    yield 'nonunit', golden_pw['nonunit_X'], np.arange(0, 30), np.arange(10, 40), golden_pw['nonunit_dot'], golden_pw['nonunit_dist']


def test_oracle_pairwise_matches_the_reference(golden, golden_pw):
    from oracle import s2_tessellation as os2
    for tag, X, iu, iv, dot, dist in _pairwise_cases(golden, golden_pw):
        d, e = os2.pairwise_dist_s2(X, iu, iv)
        # BLAS may fuse the three products differently: 1 ulp on the dot product, sqrt of a difference near the 1e-6 cut
        assert np.abs(d - dot).max() <= 4e-16 * max(1.0, np.abs(dot).max()), tag
        assert np.array_equal(e == 0, dist == 0) and np.abs(e - dist).max() <= 1e-12 * max(1.0, dist.max()), tag


@pytest.mark.gpu
def test_pairwise_dist_s2_on_the_device(golden, golden_pw):
    from manifoldem_python_b200 import FindCCGraph as F
    for tag, X, iu, iv, dot, dist in _pairwise_cases(golden, golden_pw):
        d, e = F.CalcPairwiseDistS2(X) if iu is None else F.CalcPairwiseDistS2(X, iu, iv)
        assert d.shape == dot.shape and np.abs(d - dot).max() <= 4e-16 * max(1.0, np.abs(dot).max()), tag
        assert np.array_equal(e == 0, dist == 0) and np.abs(e - dist).max() <= 1e-12 * max(1.0, dist.max()), tag
    with pytest.raises(ValueError):
        F.CalcPairwiseDistS2(golden['a_S20_th'], np.arange(3), np.arange(5))
    with pytest.raises(AssertionError):
        F.CalcPairwiseDistS2(golden['a_S20_th'], np.arange(3))
