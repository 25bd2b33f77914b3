"""GPU tests of the kNN selection kernel (a15, DMembeddingII.py:43-57) and of the neighbour lists taken straight from
the contraction's split-K partial tiles (BASELINE config 3: "kNN epilogue only", D never assembled).
Index lists are integer work: bit-exact against a host lexsort of the same row and between the two kernels."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lexsort_lists(D, k, rows):
    nS = D.shape[0]
    out = {}
    for i in rows:
        row = D[i].astype(np.float64)
        row[i] = -np.inf
        order = np.lexsort((np.arange(nS), row))[:k]
        val = row[order]
        val[0] = 0.0
        out[i] = (order.astype(np.int32), val)
    return out


def _knn(D, k, mode):
    from manifoldem_python_b200 import _lib
    lib, ctx = _lib.load(), _lib.default_context()
    nS = D.shape[0]
    Dd = _lib.DeviceArray(ctx, (nS, nS), D.dtype, D)
    idx_d = _lib.DeviceArray(ctx, (nS, k), np.int32)
    val_d = _lib.DeviceArray(ctx, (nS, k), np.float64)
    fn = lib.mem_knn_device_f32 if D.dtype == np.float32 else lib.mem_knn_device
    _lib.check(lib.mem_knn_mode(mode))
    try:
        _lib.check(fn(ctx.handle, Dd.ptr, nS, k, idx_d.ptr, val_d.ptr, None))
        idx, val = idx_d.download(), val_d.download()
    finally:
        _lib.check(lib.mem_knn_mode(0))
        for a in (Dd, idx_d, val_d):
            a.free()
    return idx, val


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('nS,k,kind', [(40, 7, 'ties'), (333, 25, 'rand'), (1000, 100, 'int'), (777, 1, 'rand'),
                                       (513, 128, 'neg'), (64, 16, 'zeros'), (50, 50, 'int'), (2000, 257, 'int'),
                                       (5000, 100, 'rand'), (5000, 2048, 'int')])
def test_selection_equals_sort_equals_lexsort(dtype, nS, k, kind):
    rng = np.random.default_rng(nS + k)
    if kind == 'ties':
        D = np.full((nS, nS), 5.0)
    elif kind == 'rand':
        D = rng.random((nS, nS)) * 1e3
    elif kind == 'int':
        D = rng.integers(0, 50, (nS, nS)).astype(float)          # many exact ties, also at the threshold
    elif kind == 'neg':
        D = rng.standard_normal((nS, nS))
    else:
        D = np.where(rng.random((nS, nS)) < 0.5, 0.0, -0.0)
    D = np.maximum(D, D.T).astype(dtype)
    idx_sel, val_sel = _knn(D, k, 2)
    idx_srt, val_srt = _knn(D, k, 1)
    idx_auto, val_auto = _knn(D, k, 0)
    assert np.array_equal(idx_sel, idx_srt) and np.array_equal(val_sel, val_srt)
    assert np.array_equal(idx_auto, idx_srt) and np.array_equal(val_auto, val_srt)
    rows = sorted(set(list(range(0, nS, max(1, nS // 9))) + [nS - 1]))
    for i, (o, v) in _lexsort_lists(D, k, rows).items():
        assert np.array_equal(idx_sel[i], o), i
        assert np.array_equal(val_sel[i], v), i


def test_long_rows_selection_and_errors():
    """C5-sized rows (20,000 entries): the selection keeps the row in shared memory; k beyond its list capacity or
    more than a quarter of the row goes to the sort; bad k fails loudly."""
    from manifoldem_python_b200 import _lib
    lib, ctx = _lib.load(), _lib.default_context()
    nS, k = 20000, 100
    rng = np.random.default_rng(7)
    D = rng.integers(0, 4000, size=(nS, nS)).astype(np.float32)
    D = np.maximum(D, D.T)
    idx, val = _knn(D, k, 0)
    for i, (o, v) in _lexsort_lists(D, k, list(range(0, nS, 2221)) + [nS - 1]).items():
        assert np.array_equal(idx[i], o) and np.array_equal(val[i], v), i
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_knn_mode(3))
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_knn_device_f32(ctx.handle, None, 10, 0, None, None, None))


@pytest.mark.parametrize('nS,N,k,kw', [(333, 64, 25, {}), (600, 64, 100, dict(split_k=3)), (257, 64, 64, {}),
                                        (300, 64, 200, {}), (260, 64, 30, dict(contraction=2)),
                                        (150, 64, 20, dict(contraction=1))])
def test_lists_from_the_contraction_equal_lists_from_D(nS, N, k, kw):
    """pd_stage.run_pd(knn_k=k): (i) with D also requested, the lists equal a host lexsort of that D; (ii) without D
    (never assembled: lists selected from the split-K partial tiles) they are identical to (i).  Covers both tcgen05
    tilings, the SIMT checker and a list too long for the selection (sort of the assembled row)."""
    from manifoldem_python_b200 import pd_stage, synthetic
    pd = synthetic.make_pd(nS, N, seed=nS + k, snr=0.3)
    em = pd['em']
    run = lambda fields: pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'],
                                         em['Cs'], em['EkV'], em['AmpContrast'], fields=fields, knn_k=k,
                                         float64=False, **kw)
    both = run(('D',))
    only = run(())
    D = both['D']
    assert D.dtype == np.float32 and np.array_equal(D, D.T)
    ref = _lexsort_lists(D, k, range(nS))
    for i in range(nS):
        assert np.array_equal(both['knn_idx'][i], ref[i][0]), i
        assert np.array_equal(both['knn_val'][i], ref[i][1]), i
    assert np.array_equal(only['knn_idx'], both['knn_idx'])
    assert np.array_equal(only['knn_val'], both['knn_val'])
    assert not only['D'].any()                                   # placeholder: D was not produced


def test_resident_chain_without_D():
    """Config-3 shape of use end to end: distance stage -> kNN lists (no D) -> graph + Ferguson sweep, equal to the
    chain that goes through the resident D."""
    from manifoldem_python_b200 import DMembeddingII, pd_stage, synthetic
    nS, N, k = 400, 64, 30
    pd = synthetic.make_pd(nS, N, seed=21, snr=0.3)
    em = pd['em']
    args = (pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'], em['EkV'],
            em['AmpContrast'])
    none, idx_d, val_d = pd_stage.run_pd_resident(*args, knn_k=k, keep_D=False)
    assert none is None
    M1, logEps, ls1, idx1, val1 = DMembeddingII.graph_and_sweep(None, k, knn=(idx_d, val_d))
    Dd = pd_stage.run_pd_resident(*args)
    M2, _, ls2, idx2, val2 = DMembeddingII.graph_and_sweep(Dd, k)
    assert np.array_equal(idx1, idx2) and np.array_equal(val1, val2)
    assert np.array_equal(M1.download(), M2.download()) and np.array_equal(ls1, ls2)
    for a in (Dd, M1, M2):
        a.free()


def test_contract_knn_entry_point_and_bad_requests():
    from manifoldem_python_b200 import _lib
    lib, ctx = _lib.load(), _lib.default_context()
    rng = np.random.default_rng(5)
    nS, n1, n3, k = 700, 3, 20, 40
    K = 32 * (2 * n1 + n3)
    Z = rng.standard_normal((nS, K)).astype(np.float32)
    Z[:, :64 * n1] = np.abs(Z[:, :64 * n1])
    hi = (Z.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = ((Z - hi).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    zhi = _lib.DeviceArray(ctx, Z.shape, np.float32, hi)
    zlo = _lib.DeviceArray(ctx, Z.shape, np.float32, lo)
    Dd = _lib.DeviceArray(ctx, (nS, nS), np.float32)
    idx_d = _lib.DeviceArray(ctx, (nS, k), np.int32)
    val_d = _lib.DeviceArray(ctx, (nS, k), np.float64)
    shp = _lib.ContractShape(nS=nS, n1_blocks=n1, n3_blocks=n3, ldz=K)
    for split in (0, 1, 4):
        _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, 0, 0, split, None))
        ctx.sync()
        D = Dd.download()
        idx_d.upload(np.zeros((nS, k), np.int32))
        _lib.check(lib.mem_contract_knn_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, None, k, idx_d.ptr,
                                               val_d.ptr, 0, 0, split, None))
        ctx.sync()
        idx, val = idx_d.download(), val_d.download()
        for i, (o, v) in _lexsort_lists(D, k, range(0, nS, 37)).items():
            assert np.array_equal(idx[i], o) and np.array_equal(val[i], v), (split, i)
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_contract_knn_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, None, 0, idx_d.ptr,
                                               val_d.ptr, 0, 0, 0, None))
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_contract_knn_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, None, nS + 1, idx_d.ptr,
                                               val_d.ptr, 0, 0, 0, None))
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, None, 0, 0, 0, None))
    for a in (zhi, zlo, Dd, idx_d, val_d):
        a.free()


def test_full_size_config4_pd_lists_and_properties():
    """BASELINE config 4 PD at full size (2,000 particles x 256^2), no oracle: D is symmetric with a ~zero diagonal,
    a duplicated particle is at distance ~0 from its twin, and the k = 100 neighbour lists selected from the partial
    tiles (D not requested) are exactly the lexsort of the D the other call returns; each twin is the other's nearest
    neighbour."""
    from manifoldem_python_b200 import pd_stage, synthetic
    nS, N, k = 2000, 256, 100
    rng = np.random.default_rng(11)
    stack = rng.standard_normal((nS, N * N), dtype=np.float32)
    stack[1] = stack[0]
    em = synthetic.make_pd(4, 16, seed=0)['em']
    q = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS), 1.1 + 0.03 * rng.standard_normal(nS),
                                rng.uniform(0, 2 * np.pi, nS))
    df = rng.uniform(10000, 30000, nS)
    q[:, 1], df[1] = q[:, 0], df[0]
    run = lambda fields, kk: pd_stage.run_pd(np.arange(nS), q, df, stack.reshape(-1), 2 * nS, N, em['pix_size'], em['Cs'],
                                             em['EkV'], em['AmpContrast'], fields=fields, knn_k=kk, float64=False)
    D = run(('D',), 0)['D']
    assert np.array_equal(D, D.T) and np.abs(np.diag(D)).max() <= 1e-5 * D.max() and abs(D[0, 1]) <= 1e-5 * D.max()
    only = run((), k)
    ref = _lexsort_lists(D, k, range(0, nS, 7))
    for i, (o, v) in ref.items():
        assert np.array_equal(only['knn_idx'][i], o) and np.array_equal(only['knn_val'][i], v), i
    assert only['knn_idx'][0, 1] == 1 and only['knn_idx'][1, 1] == 0
