"""GPU tests: diffusion-map front end (kNN, graph, Ferguson sweep, Laplacian) against the oracle and the
reference-generated golden vectors; the three drop-in callables end to end (pickle layout, resume markers)."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load_dm(golden_dir):
    return np.load(os.path.join(golden_dir, 'dm_nS72.npz'))


@pytest.mark.parametrize('k', [72, 20])
def test_dm_against_reference_golden(golden_dir, k):
    """DMembeddingII.op drop-in vs the reference's own outputs on the same D (same np.random draws)."""
    from manifoldem_python_b200 import DMembeddingII, p
    p.init()
    g = _load_dm(golden_dir)
    D = g['D'].copy()
    np.random.seed(1234)
    lamb, psi, sigma, mu, logEps, logSumWij, popt, R2 = DMembeddingII.op(D, k, 3.0, 60000)
    assert np.isneginf(D[3, 3])                                              # mutated in place like the reference
    assert np.array_equal(logEps, g['logEps'])
    assert np.allclose(logSumWij, g[f'k{k}_logSumWij'], rtol=1e-11, atol=1e-11)
    assert np.allclose(popt, g[f'k{k}_popt'], rtol=1e-6)
    assert abs(sigma - float(g[f'k{k}_sigma'])) <= 1e-8 * sigma
    assert np.allclose(lamb, g[f'k{k}_lamb'], rtol=1e-8, atol=1e-10)
    assert np.allclose(mu, g[f'k{k}_mu'], rtol=1e-6, atol=1e-12)
    assert abs(R2 - float(g[f'k{k}_R2'])) < 1e-9
    ref_psi = g[f'k{k}_psi']
    assert psi.shape == ref_psi.shape
    for j in range(5):                                                       # |corr| >= 0.9999 (north_star)
        assert abs(np.corrcoef(psi[:, j], ref_psi[:, j])[0, 1]) > 0.9999, j


@pytest.mark.parametrize('nS,k', [(300, 300), (300, 40), (1000, 100), (77, 77)])
def test_knn_graph_laplacian_vs_oracle(nS, k):
    from manifoldem_python_b200 import DMembeddingII, _lib
    from oracle import dm_embedding as odm
    rng = np.random.default_rng(nS + k)
    X = rng.standard_normal((nS, 6)) * np.array([3, 2, 1, 1, 0.5, 0.5])
    D = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) * 1e3                 # squared distances, symmetric
    D = (D + D.T) / 2
    M, logEps, logSumWij, idx, val = DMembeddingII.graph_and_sweep(D, k)
    Do = D.copy()
    idx_o, val_o = odm.knn_lists(Do, k)                                      # (k, nS)
    assert np.array_equal(idx[:, 0], np.arange(nS)) and not val[:, 0].any()
    assert np.array_equal(val, val_o.T)                                      # distances bit-exact
    for i in range(nS):                                                      # identical kNN index sets (no ties here)
        assert set(idx[i]) == set(idx_o[:, i])
    assert np.array_equal(idx, idx_o.T)
    yRow, yCol, yVal = odm.symmetrise(idx_o, val_o, nS)
    Mo = -np.ones((nS, nS))
    Mo[yRow, yCol] = yVal
    assert np.array_equal(M.download(), Mo)                                  # graph bit-exact (integer/index work)
    ls_o, thr = odm.ferguson_logsum(np.sqrt(yVal))
    assert thr == 10
    assert np.allclose(logSumWij, ls_o, rtol=1e-12, atol=1e-12)
    sigma = 3.0 * np.sqrt(np.median(yVal[yVal > 0]))
    L = DMembeddingII.laplacian(M, nS, sigma)
    Lo = odm.laplacian(yVal, yCol, yRow, nS, sigma).toarray()
    assert np.abs(L - Lo).max() <= 1e-12 * np.abs(Lo).max()
    M.free()


def test_knn_ties_and_errors():
    from manifoldem_python_b200 import DMembeddingII, _lib
    nS = 40
    D = np.ones((nS, nS)) * 5.0                                              # all off-diagonal distances tie
    np.fill_diagonal(D, 0.0)
    M, _, _, idx, val = DMembeddingII.graph_and_sweep(D, 7)
    assert np.array_equal(idx[:, 0], np.arange(nS))
    assert (val[:, 1:] == 5.0).all()
    for i in range(nS):                                                      # ties broken by index: the smallest others
        assert list(idx[i, 1:]) == [j for j in range(nS) if j != i][:6]
    M.free()
    lib, ctx = _lib.load(), _lib.default_context()
    with pytest.raises(RuntimeError):
        _lib.check(lib.mem_knn_device(ctx.handle, None, 10, 11, None, None, None))


@pytest.mark.parametrize('nS,dtype', [(16500, np.float32), (20000, np.float32), (16390, np.float64)])
def test_knn_rows_longer_than_one_smem_sort(nS, dtype):
    """C5-sized PDs (nS = 20,000): the chunked bitonic network (shared-memory chunks + global strides) gives the
    same (value, index) order as a host sort, duplicates included."""
    from manifoldem_python_b200 import _lib
    lib, ctx = _lib.load(), _lib.default_context()
    rng = np.random.default_rng(nS)
    D = rng.integers(0, 4000, size=(nS, nS)).astype(dtype)                    # many exact ties per row
    D = np.maximum(D, D.T)
    k = 257
    Dd = _lib.DeviceArray(ctx, (nS, nS), dtype, D)
    idx_d = _lib.DeviceArray(ctx, (nS, k), np.int32)
    val_d = _lib.DeviceArray(ctx, (nS, k), np.float64)
    fn = lib.mem_knn_device_f32 if dtype == np.float32 else lib.mem_knn_device
    _lib.check(lib.mem_knn_mode(1))                                          # the sort (automatic choice here: selection)
    try:
        _lib.check(fn(ctx.handle, Dd.ptr, nS, k, idx_d.ptr, val_d.ptr, None))
        idx, val = idx_d.download(), val_d.download()
    finally:
        _lib.check(lib.mem_knn_mode(0))
        for a in (Dd, idx_d, val_d):
            a.free()
    for i in list(range(0, nS, 1237)) + [nS - 1]:
        row = D[i].astype(np.float64)
        row[i] = -np.inf
        order = np.lexsort((np.arange(nS), row))[:k]
        assert np.array_equal(idx[i], order), i
        expect = row[order]
        expect[0] = 0.0
        assert np.array_equal(val[i], expect), i


def test_dropin_worker_and_driver(tmp_path):
    """GetDistancesS2.op -> getDistanceCTF_local_Conj9combinedS2.op: pickle keys/shapes/dtypes as the
    reference writes them, markers after the dump, finished PDs skipped on a second run."""
    from manifoldem_python_b200 import GetDistancesS2, myio, p, synthetic
    from oracle import pd_distance as opd
    N = 32
    pds = [synthetic.make_pd(n, N, seed=20 + i, snr=0.5) for i, n in enumerate((24, 17, 30))]
    # one augmented data set: concatenate the three PDs' particles
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
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 'demo'
    p.create_dir()
    em = pds[0]['em']
    p.pix_size, p.Cs, p.EkV, p.AmpContrast = em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
    p.relion_data, p.ncpu, p.num_part = False, 1, n_half
    p.img_stack_file = str(tmp_path / 'stack.dat')
    stack.tofile(p.img_stack_file)
    p.numberofJobs = 3
    myio.fout1(p.tess_file, ['CG', 'df', 'q', 'sh'], [CG, df, q, (np.zeros(n_half), np.zeros(n_half))])

    class Sig:
        vals = []

        def emit(self, v):
            self.vals.append(v)
    sig = Sig()
    GetDistancesS2.op(sig)
    assert p.nPix == N and sig.vals[0] == 0 and sig.vals[-1] == 100
    assert sorted(os.listdir(p.dist_prog)) == ['0', '1', '2']
    for prD, pd in enumerate(pds):
        with open('{}prD_{}'.format(p.dist_file, prD), 'rb') as f:
            rec = pickle.load(f)
        nS = len(CG[prD])
        assert list(rec.keys()) == ['D', 'ind', 'q', 'df', 'CTF', 'imgAll', 'msk2', 'PD', 'PDs', 'Psis', 'imgAvg',
                                    'imgAvgFlip', 'imgAllFlip', 'imgLabels', 'Dnom', 'Nom', 'imgAllIntensity',
                                    'version', 'options']
        shapes = dict(D=(nS, nS), CTF=(nS, N * N), imgAll=(nS, N, N), imgAllFlip=(nS, N, N), PD=(3,), PDs=(3, nS),
                      Psis=(nS, 1), imgAvg=(N, N), imgAvgFlip=(N, N), Dnom=(nS, 1), Nom=(nS, 1),
                      imgAllIntensity=(N, N), q=(4, nS), df=(nS,))
        for k, shp in shapes.items():
            assert rec[k].shape == shp and rec[k].dtype == np.float64, k
        assert rec['msk2'] == 1 and rec['version'] == 'getDistanceCTF_local9, V 1.0'
        assert rec['options']['relion_data'] is False
        ref = opd.pd_distance(CG[prD], q[:, CG[prD]], df[CG[prD]], stack, 2 * n_half, N, em['pix_size'], em['Cs'],
                              em['EkV'], em['AmpContrast'], rotate_impl='periodic')
        offd = ~np.eye(nS, dtype=bool)
        assert (np.abs(rec['D'] - ref['D'])[offd] / ref['D'][offd]).max() < 1e-5
        assert np.array_equal(rec['imgLabels'], ref['imgLabels'])
    # resume: remove one marker, only that PD is recomputed
    os.remove(os.path.join(p.dist_prog, '1'))
    mt = {i: os.path.getmtime('{}prD_{}'.format(p.dist_file, i)) for i in range(3)}
    GetDistancesS2.op()
    assert sorted(os.listdir(p.dist_prog)) == ['0', '1', '2']
    assert os.path.getmtime('{}prD_0'.format(p.dist_file)) == mt[0]
    assert os.path.getmtime('{}prD_1'.format(p.dist_file)) >= mt[1]


def test_manifold_trimming_consumer(tmp_path):
    """The consumer chain of the embedding stage (manifoldTrimmingAuto.py:44-70): D from the PD pickle ->
    DMembeddingII.op(D, k=nS) -> leading eigenvector recovers the 1-D latent coordinate of the synthetic PD."""
    from manifoldem_python_b200 import DMembeddingII, p, pd_stage, synthetic
    p.init()
    pd = synthetic.make_pd(160, 48, seed=31, snr=3.0)
    em = pd['em']
    D = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], 48, em['pix_size'], em['Cs'],
                        em['EkV'], em['AmpContrast'], fields=('D',))['D']
    np.random.seed(0)
    lamb, psi, sigma, mu, logEps, logSumWij, popt, R2 = DMembeddingII.op(D.copy(), 160, 3.0, 60000)
    assert lamb[0] == pytest.approx(1.0, abs=1e-6) and (np.diff(lamb) <= 1e-12).all()
    assert abs(mu.sum() - 1.0) < 1e-8
    # the same chain entirely on the CPU oracle (float64 reference D -> oracle embedding, same random draws)
    from oracle import dm_embedding as odm, pd_distance as opd
    Dr = opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], 48, em['pix_size'], em['Cs'],
                         em['EkV'], em['AmpContrast'], rotate_impl='periodic', keep=('D',))['D']
    np.random.seed(0)
    lamb_o, psi_o, sigma_o, mu_o = odm.dm_embedding(Dr.copy(), 160, 3.0)[:4]
    assert abs(sigma - sigma_o) <= 1e-5 * sigma_o
    assert np.allclose(lamb[:6], lamb_o[:6], rtol=1e-4)
    for j in range(3):                                   # north_star: leading eigenvectors |corr| >= 0.9999
        assert abs(np.corrcoef(psi[:, j], psi_o[:, j])[0, 1]) >= 0.9999, j


def test_resident_chain_distance_to_knn():
    """BASELINE config 3 shape of use: D stays on the device, kNN (k < nS) + Ferguson sweep over the compacted
    edges; same lists / curve as the host-D path on the downloaded matrix."""
    from manifoldem_python_b200 import DMembeddingII, pd_stage, synthetic
    pd = synthetic.make_pd(333, 64, seed=77, snr=0.3)
    em = pd['em']
    Dd = pd_stage.run_pd_resident(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], 64, em['pix_size'], em['Cs'],
                                  em['EkV'], em['AmpContrast'])
    k = 25
    M1, logEps, ls1, idx1, val1 = DMembeddingII.graph_and_sweep(Dd, k)
    D = Dd.download().astype(np.float64)
    M2, _, ls2, idx2, val2 = DMembeddingII.graph_and_sweep(D, k)
    assert np.array_equal(idx1, idx2) and np.array_equal(val1, val2)
    assert np.array_equal(M1.download(), M2.download())
    assert np.array_equal(ls1, ls2)
    # compacted sweep (k < nS) == dense sweep of the oracle
    from oracle import dm_embedding as odm
    Do = D.copy()
    io, vo = odm.knn_lists(Do, k)
    yRow, yCol, yVal = odm.symmetrise(io, vo, 333)
    ls_o, _ = odm.ferguson_logsum(np.sqrt(yVal))
    assert np.allclose(ls1, ls_o, rtol=1e-12, atol=1e-12)
    for a in (Dd, M1, M2):
        a.free()


def test_manifold_trimming_dropin(tmp_path):
    """manifoldTrimmingAuto.op drop-in (modules/manifoldTrimmingAuto.py:38-95) on a PD pickle written by the
    distance worker: same trimming loop, psi pickle keys, marker and eig_spec file; the final embedding matches the
    same loop run on the CPU oracle."""
    from manifoldem_python_b200 import manifoldTrimmingAuto as mta, getDistanceCTF_local_Conj9combinedS2 as worker
    from manifoldem_python_b200 import myio, p, synthetic
    from oracle import dm_embedding as odm
    N, nS = 32, 90
    pd = synthetic.make_pd(nS, N, seed=61, snr=2.0)
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 'trim'
    p.create_dir()
    em = pd['em']
    p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast = N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
    stack_file = str(tmp_path / 'stack.dat')
    pd['stack'].tofile(stack_file)
    dist_file = '{}prD_{}'.format(p.dist_file, 0)
    worker.op([pd['ind'], pd['q'], pd['df'], dist_file, 0], dict(type='Butter', Qc=0.5, N=8), stack_file, pd['sh'],
              pd['nStot'], dict(verbose=False, avgOnly=False, visual=False, parallel=False, relion_data=False, thres=2000),
              fields=('D',))
    psi_file = '{}prD_{}'.format(p.psi_file, 0)
    eig_file = str(tmp_path / 'eig_spec.txt')
    rad = 3.0                                                       # small enough to trigger the trimming loop (6 passes)
    np.random.seed(5)
    mta.op([dist_file, psi_file, eig_file, 0], 0, 3.0, rad, False, dict(outputFile='', Is=True))
    rec = myio.fin1(psi_file)
    assert list(rec.keys()) == ['lamb', 'psi', 'sigma', 'mu', 'posPath', 'ind', 'logEps', 'logSumWij', 'popt', 'R_squared']
    assert os.path.exists(os.path.join(p.psi_prog, '0'))
    n_keep = len(rec['posPath'])
    assert 3 < n_keep <= nS and rec['psi'].shape[0] == n_keep and rec['mu'].shape == (n_keep,)
    assert np.all(np.sqrt((rec['psi'][:, :3] ** 2).sum(1)) < rad)
    lines = open(eig_file).read().strip().splitlines()
    assert len(lines) == len(rec['lamb']) - 1 and lines[0].split()[0] == '1'
    # the same loop on the oracle
    D = myio.fin1(dist_file)['D']
    np.random.seed(5)
    pos = np.arange(nS)
    out = odm.dm_embedding(D.copy(), nS, 3.0)
    psi = out[1]
    pos1 = mta.get_psiPath(psi, rad, 0)
    n = nS
    while len(pos1) < n:
        n = len(pos1)
        D1 = D[pos1][:, pos1]
        out = odm.dm_embedding(D1.copy(), n, 3.0)
        pos1 = pos1[mta.get_psiPath(out[1], rad, 0)]
    assert np.array_equal(rec['posPath'], pos[pos1])
    for j in range(3):
        assert abs(np.corrcoef(rec['psi'][:, j], out[1][:, j])[0, 1]) >= 0.9999


def test_sidecar_records_through_the_worker(tmp_path):
    """The per-PD worker with p.record_layout = 'sidecar': every key read back through myio equals the default
    pickle record of the same PD bit for bit (float32 on disk, promoted to the reference's float64 on access), the
    heavy arrays sit in .npy files beside a small manifest, and manifoldTrimmingAuto.op consumes the record."""
    from manifoldem_python_b200 import manifoldTrimmingAuto as mta, getDistanceCTF_local_Conj9combinedS2 as worker
    from manifoldem_python_b200 import myio, p, synthetic
    N, nS = 64, 60
    pd = synthetic.make_pd(nS, N, seed=62, snr=2.0)
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 'rec'
    p.create_dir()
    em = pd['em']
    p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast = N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
    stack_file = str(tmp_path / 'stack.dat')
    pd['stack'].tofile(stack_file)
    opts = dict(verbose=False, avgOnly=False, visual=False, parallel=False, relion_data=False, thres=2000)
    recs = {}
    try:
        for prD, layout in enumerate(('pickle', 'sidecar')):
            p.record_layout = layout
            f = '{}prD_{}'.format(p.dist_file, prD)
            worker.op([pd['ind'], pd['q'], pd['df'], f, prD], dict(type='Butter', Qc=0.5, N=8), stack_file, pd['sh'],
                      pd['nStot'], opts)
            assert os.path.exists(os.path.join(p.dist_prog, str(prD)))
            recs[layout] = myio.fin1(f)
    finally:
        del p.record_layout
    a, b = recs['sidecar'], recs['pickle']
    assert isinstance(a, myio.Record) and list(a.keys()) == list(b.keys())
    assert os.path.getsize('{}prD_1'.format(p.dist_file)) < 0.05 * os.path.getsize('{}prD_0'.format(p.dist_file))
    assert os.path.exists('{}prD_1.imgAll.npy'.format(p.dist_file))
    assert a.raw('imgAll').dtype == np.float32
    assert a._lazy['CTF']['virtual'] == 'ctf' and not os.path.exists('{}prD_1.CTF.npy'.format(p.dist_file))
    for k in b:
        if isinstance(b[k], np.ndarray):
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k
    np.random.seed(5)
    mta.op(['{}prD_1'.format(p.dist_file), '{}prD_1'.format(p.psi_file), str(tmp_path / 'eig.txt'), 1], 0, 3.0, 5.0,
           False, dict(outputFile='', Is=True))
    rec = myio.fin1('{}prD_1'.format(p.psi_file))
    assert rec['psi'].shape[0] == len(rec['posPath']) and os.path.exists(os.path.join(p.psi_prog, '1'))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_device_submatrix_and_resident_embedding(dtype):
    """DMembeddingII.take == D[sel][:, sel] (manifoldTrimmingAuto.py:50,63) on the device, and embedding a resident
    matrix gives what op() gives for the same host matrix (float32 storage included: identical neighbour lists)."""
    from manifoldem_python_b200 import DMembeddingII, _lib, p
    p.init()
    rng = np.random.default_rng(8)
    nS = 300
    tau = rng.random(nS)
    X = np.stack([np.cos(3 * tau), np.sin(3 * tau), 0.2 * rng.standard_normal(nS)], 1)
    D = (((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) * 1e4).astype(np.float32).astype(dtype)
    D = np.maximum(D, D.T)
    Dd = DMembeddingII.upload(D)
    assert Dd.dtype == dtype
    sel = np.sort(rng.choice(nS, 123, replace=False))
    sub = DMembeddingII.take(Dd, sel)
    assert sub.shape == (123, 123) and np.array_equal(sub.download(), D[sel][:, sel])
    np.random.seed(3)
    a = DMembeddingII.embed(sub, 123, 3.0)
    np.random.seed(3)
    b = DMembeddingII.op(D[sel][:, sel].astype(np.float64), 123, 3.0, 0)
    assert np.array_equal(a[5], b[5]) and a[2] == b[2]                        # logSumWij, sigma
    for j in range(3):
        assert abs(np.corrcoef(a[1][:, j], b[1][:, j])[0, 1]) >= 0.9999
    with pytest.raises(RuntimeError):
        DMembeddingII.take(Dd, np.array([0, nS]))
    for x in (Dd, sub):
        x.free()


def _diffusion_like_matrix(nS, seed):
    """Symmetric normalised Gaussian-kernel matrix of points on a noisy curve: top eigenvalue 1, decaying spectrum."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 1, nS))
    X = np.stack([np.cos(3 * t), np.sin(3 * t), t], 1) + 0.02 * rng.standard_normal((nS, 3))
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    W = np.exp(-d2 / 0.05)
    d = W.sum(1)
    W = W / np.outer(d, d)
    d = np.sqrt(W.sum(1))
    L = W / np.outer(d, d)
    return np.abs(L + L.T) / 2


@pytest.mark.parametrize('nS,k', [(40, 16), (333, 16), (2000, 16), (2000, 5)])
def test_device_lanczos_vs_dense_eigh(nS, k):
    """eigsh_device (sembeddingonFly.py:27 on the device) against numpy's dense eigh of the same matrix: eigenvalues to
    1e-10, every eigenvector |cos| >= 1 - 1e-8 (the north-star gate is |corr| >= 0.9999)."""
    from manifoldem_python_b200 import DMembeddingII, _lib
    from manifoldem_python_b200.getDistanceCTF_local_Conj9combinedS2 import _ctx
    L = _diffusion_like_matrix(nS, 5 + nS)
    w, U = np.linalg.eigh(L)
    order = np.argsort(-np.abs(w))[:k]
    Ld = _lib.DeviceArray(_ctx(), (nS, nS), np.float64, L)
    vals, vecs, info = DMembeddingII.eigsh_device(Ld, nS, k)
    Ld.free()
    assert info['converged'] and vals.shape == (k,) and vecs.shape == (nS, k)
    ix = np.argsort(-np.abs(vals))
    assert np.allclose(vals[ix], w[order], rtol=0, atol=1e-10)
    assert np.allclose(np.linalg.norm(vecs, axis=0), 1.0, atol=1e-12)
    for a, b in zip(ix, order):
        assert abs(vecs[:, a] @ U[:, b]) >= 1 - 1e-8, (a, b)
    # orthonormal Ritz vectors
    assert np.abs(vecs.T @ vecs - np.eye(k)).max() < 1e-10


def test_both_eigen_solvers_give_the_same_embedding(golden_dir):
    """embed() through the device Lanczos (default) and through the reference's ARPACK call with the device operator."""
    from manifoldem_python_b200 import DMembeddingII, p
    p.init()
    g = _load_dm(golden_dir)
    out = {}
    for solver in ('lanczos', 'arpack'):
        p.eig_solver = solver
        np.random.seed(7)
        out[solver] = DMembeddingII.embed(g['D'].copy(), 72, 3.0)
    del p.eig_solver
    a, b = out['lanczos'], out['arpack']
    assert np.allclose(a[0], b[0], rtol=1e-9, atol=1e-11)                 # lamb
    assert np.allclose(a[3], b[3], rtol=1e-7, atol=1e-12)                 # mu
    for j in range(8):
        assert abs(np.corrcoef(a[1][:, j], b[1][:, j])[0, 1]) > 0.999999, j


@pytest.mark.parametrize('relion', [False, True])
def test_virtual_image_arrays_in_sidecar_records(tmp_path, relion):
    """p.record_virtual_images: the record keeps the recipe (stack path; ind, q, df are in it anyway) instead of imgAll /
    imgAllFlip; reading the keys rebuilds them with the distance stage's own kernels — bit for bit what the default
    record stores — and every other key is unchanged.  SPIDER and RELION (per-member shifts) stacks."""
    from manifoldem_python_b200 import getDistanceCTF_local_Conj9combinedS2 as worker
    from manifoldem_python_b200 import myio, p, synthetic
    N, nS = 64, 40
    pd = synthetic.make_pd(nS, N, seed=63, snr=2.0)
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 'virt'
    p.create_dir()
    em = pd['em']
    p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast = N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
    stack_file = str(tmp_path / ('stack.mrcs' if relion else 'stack.dat'))
    sh = pd['sh']
    if relion:
        rng = np.random.default_rng(1)
        n_half = pd['nStot'] // 2
        sh = (rng.uniform(-2, 2, n_half), rng.uniform(-2, 2, n_half))
        hdr = np.zeros(256, dtype='<i4')
        hdr[0], hdr[1], hdr[2], hdr[3] = N, N, n_half, 2
        with open(stack_file, 'wb') as f:
            f.write(hdr.tobytes())
            f.write(pd['stack'].astype('<f4').tobytes())
    else:
        pd['stack'].tofile(stack_file)
    opts = dict(verbose=False, avgOnly=False, visual=False, parallel=False, relion_data=relion, thres=2000)
    recs = {}
    try:
        for prD, virt in enumerate((False, True)):
            p.record_layout, p.record_virtual_images = 'sidecar', virt
            f = '{}prD_{}'.format(p.dist_file, prD)
            worker.op([pd['ind'], pd['q'], pd['df'], f, prD], dict(type='Butter', Qc=0.5, N=8), stack_file, sh, pd['nStot'], opts)
            recs[virt] = myio.fin1(f)
    finally:
        del p.record_layout, p.record_virtual_images
    a, b = recs[True], recs[False]
    assert a._lazy['imgAll']['virtual'] == 'images' and a._lazy['imgAllFlip']['virtual'] == 'images'
    assert not os.path.exists('{}prD_1.imgAll.npy'.format(p.dist_file)) and os.path.exists('{}prD_0.imgAll.npy'.format(p.dist_file))
    assert list(a.keys()) == list(b.keys())
    img = a['imgAll']                                   # one recomputation fills both virtual image arrays
    assert 'imgAllFlip' not in a._lazy
    for k in b:
        if isinstance(b[k], np.ndarray):
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k
    assert img.dtype == np.float64 and img.shape == (nS, N, N)


def test_device_lanczos_edge_cases():
    """k close to nS (the recurrence runs to j = nS and becomes exact), a matrix with an invariant start subspace
    (block-diagonal: the solver still returns k pairs of the reachable block or reports non-convergence), tiny nS."""
    from manifoldem_python_b200 import DMembeddingII, _lib
    from manifoldem_python_b200.getDistanceCTF_local_Conj9combinedS2 import _ctx
    L = _diffusion_like_matrix(24, 3)
    w = np.linalg.eigvalsh(L)
    Ld = _lib.DeviceArray(_ctx(), (24, 24), np.float64, L)
    vals, vecs, info = DMembeddingII.eigsh_device(Ld, 24, 20)
    Ld.free()
    assert info['converged'] and info['steps'] <= 24
    assert np.allclose(np.sort(np.abs(vals)), np.sort(np.abs(w))[-20:], atol=1e-10)
    assert np.abs(L @ vecs - vecs * vals).max() < 1e-9
    vals2, vecs2, info2 = DMembeddingII.eigsh_device(_lib.DeviceArray(_ctx(), (3, 3), np.float64, np.diag([3.0, 2.0, 1.0])), 3, 5)
    assert vals2.shape[0] <= 2 and abs(vals2[0]) >= abs(vals2[-1])
