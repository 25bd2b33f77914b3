"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, host-side logic of the
drop-in modules (job list / resume markers, gather, MRC reader, partition, projectMask), and the SciPy
semantics the device kernels were written against."""
import os
import pickle
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from manifoldem_python_b200 import _lib
    lib = _lib.load()                                   # dlopen only, no CUDA call
    hdr = open(os.path.join(ROOT, 'include', 'manifoldem_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(mem_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.SYMBOLS) == declared
    assert lib.mem_version() == 100


def test_host_angles_match_oracle():
    from manifoldem_python_b200 import pd_stage, synthetic
    from oracle import pd_distance as opd
    pd = synthetic.make_pd(30, 16, seed=5)
    PDs, PD, psi_p, Psi, s, c = pd_stage.host_angles(pd['q'])
    PDs_o = opd.calc_avg_pd(pd['q'])
    PD_o = PDs_o.sum(1) / np.linalg.norm(PDs_o.sum(1))
    Psi_o, s_o, c_o = opd.get_psi(pd['q'], PD_o)
    assert np.array_equal(PDs, PDs_o) and np.array_equal(PD, PD_o)
    assert np.array_equal(Psi, Psi_o) and np.array_equal(s, s_o) and np.array_equal(c, c_o)
    assert psi_p == opd.psi_ang(PD_o)


def test_gather_conjugates_and_order():
    from manifoldem_python_b200 import pd_stage
    N, n_half = 4, 6
    stack = np.arange(n_half * N * N, dtype=np.float32)
    ind = np.array([7, 2, 11, 0, 5])                       # >= 6 are conjugates of 1, 5
    raw, flip, base = pd_stage.gather(stack, ind, 2 * n_half, N)
    assert list(base) == [1, 2, 5, 0, 5] and list(flip) == [1, 0, 1, 0, 0]
    for i, b in enumerate(base):
        assert np.array_equal(raw[i], stack[b * N * N:(b + 1) * N * N])
    raw3, _, _ = pd_stage.gather(stack.reshape(n_half, N, N), ind, 2 * n_half, N)
    assert np.array_equal(raw3, raw)


def test_open_stack_spider_and_mrc(tmp_path):
    from manifoldem_python_b200 import pd_stage
    N, n = 6, 5
    data = np.random.default_rng(0).standard_normal((n, N, N)).astype(np.float32)
    f = tmp_path / 'stack.dat'
    data.tofile(f)
    st = pd_stage.open_stack(str(f), N, False)
    assert st.shape == (n * N * N,) and np.array_equal(np.asarray(st), data.reshape(-1))
    hdr = np.zeros(256, dtype='<i4')
    hdr[0], hdr[1], hdr[2], hdr[3], hdr[23] = N, N, n, 2, 80
    m = tmp_path / 'stack.mrcs'
    with open(m, 'wb') as fh:
        fh.write(hdr.tobytes())
        fh.write(b'\0' * 80)
        fh.write(data.tobytes())
    st = pd_stage.open_stack(str(m), N, True)
    assert st.shape == (n, N, N) and np.array_equal(np.asarray(st), data)
    hdr[3] = 1
    with open(m, 'wb') as fh:
        fh.write(hdr.tobytes())
    with pytest.raises(ValueError):
        pd_stage.open_stack(str(m), N, True)


def test_myio_semantics(tmp_path):
    from manifoldem_python_b200 import myio
    f = str(tmp_path / 'x.pkl')
    myio.fout1(f, ['a', 'b'], [1, np.arange(3)])
    d = myio.fin1(f)
    assert d['a'] == 1 and np.array_equal(d['b'], np.arange(3))
    with open(f, 'rb') as fh:
        assert pickle.load(fh).keys() == d.keys()
    open(f, 'wb').write(b'garbage')
    assert myio.fin1(f) is None                          # unreadable pickle -> None (myio.py:21-30)
    with pytest.raises(FileNotFoundError):               # the reference opens the file outside its try (myio.py:22)
        myio.fin1(str(tmp_path / 'missing'))
    myio.fout2(f, dict(z=3))
    assert myio.fin1(f) == dict(z=3)


def test_divide_skips_finished_pds(tmp_path):
    """Resume protocol (GetDistancesS2.py:29-47): PDs with a marker in dist_prog are not re-queued."""
    from manifoldem_python_b200 import GetDistancesS2, p
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 't'
    p.create_dir()
    CG = [np.array([0, 3]), np.array([1, 2, 5]), np.array([4])]
    q = np.arange(24, dtype=float).reshape(4, 6)
    df = np.arange(6, dtype=float)
    jobs = GetDistancesS2.divide(CG, q, df, 3)
    assert [j[4] for j in jobs] == [0, 1, 2]
    assert np.array_equal(jobs[1][1], q[:, [1, 2, 5]]) and np.array_equal(jobs[1][2], df[[1, 2, 5]])
    assert jobs[2][3].endswith('IMGs_prD_2')
    open(os.path.join(p.dist_prog, '1'), 'a').close()
    open(os.path.join(p.dist_prog, '.hidden'), 'a').close()
    assert GetDistancesS2.fileCheck() == [1]
    assert [j[4] for j in GetDistancesS2.divide(CG, q, df, 3)] == [0, 2]
    assert GetDistancesS2.count(3) == 2


def test_lpt_partition_properties():
    from manifoldem_python_b200 import partition
    rng = np.random.default_rng(1)
    nS = rng.integers(100, 2001, size=53)
    costs = [partition.pd_cost(int(n), 256) for n in nS]
    for g in (1, 2, 4, 8):
        shards = partition.lpt_partition(costs, g)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(53))                   # every PD exactly once
        assert partition.imbalance(costs, shards) <= partition.imbalance(costs, partition.round_robin(53, g)) + 1e-12
        assert partition.imbalance(costs, shards) < 1.0 + max(costs) / (sum(costs) / g)
        assert shards == partition.lpt_partition(costs, g)   # deterministic
    eq = partition.lpt_partition([1.0] * 1000, 8)
    assert all(len(s) == 125 for s in eq)
    assert partition.lpt_partition([], 4) == [[], [], [], []]


def test_project_mask_matches_reference_golden(golden_dir):
    from manifoldem_python_b200 import projectMask
    g = np.load(os.path.join(golden_dir, 'pd_volmask_N24.npz'))
    msk2 = projectMask.op(g['mask3d'], g['ref_PD'])
    assert msk2.dtype == bool and np.array_equal(msk2, g['ref_msk2'])


def test_scipy_wrap_shift_model():
    """The algorithm of k_shift (align.cu): mirror prefilter + coordinate wrap with period N-1 + mirrored
    support indices reproduces scipy.ndimage.shift(order=3, mode='wrap') (getDistanceCTF...py:264)."""
    from scipy import ndimage
    rng = np.random.default_rng(0)
    N = 17
    img = rng.standard_normal((N, N))
    for s in ((0.3, -2.7), (5.5, -0.5), (-16.2, 18.9)):
        ref = ndimage.shift(img, s, order=3, mode='wrap')
        c = ndimage.spline_filter(img, order=3, mode='mirror')

        def taps(o, sh):
            x = o - sh
            L = N - 1
            if x < 0:
                x += L * (int(-x / L) + 1)
            elif x > L:
                x -= L * int(x / L)
            i0 = int(np.floor(x))
            t = x - i0
            u = 1 - t
            w = [u ** 3 / 6, 2 / 3 + t * t * (0.5 * t - 1), 2 / 3 + u * u * (0.5 * u - 1), t ** 3 / 6]
            k = []
            for a in range(4):
                kk = i0 - 1 + a
                kk = -kk if kk < 0 else kk
                kk = 2 * (N - 1) - kk if kk > N - 1 else kk
                k.append(kk)
            return k, w
        out = np.zeros_like(img)
        for r in range(N):
            ka, wa = taps(r, s[0])
            for cc in range(N):
                kb, wb = taps(cc, s[1])
                out[r, cc] = sum(wa[a] * wb[b] * c[ka[a], kb[b]] for a in range(4) for b in range(4))
        assert np.abs(out - ref).max() < 1e-12
    # mirror prefilter == periodic prefilter of the (2N-2)-long mirror extension (how the kernel realises it)
    line = rng.standard_normal(N)
    ext = np.concatenate([line, line[-2:0:-1]])
    per = ndimage.spline_filter1d(ext, order=3, mode='grid-wrap')[:N]
    assert np.abs(per - ndimage.spline_filter1d(line, order=3, mode='mirror')).max() < 1e-12


def test_segmented_recursive_prefilter_model():
    """The segmented carry scheme of k_prefilter_rows / _cols (align.cu) in NumPy float64 against
    scipy's periodic cubic spline filter."""
    from scipy import ndimage
    z = np.sqrt(3.0) - 2.0
    rng = np.random.default_rng(3)
    for L, E in ((256, 8), (100, 4), (25, 1), (510, 16), (37, 3)):
        s = rng.standard_normal(L)
        used = (L + E - 1) // E
        H = min(used, 20 // E + 2)
        segs = [np.arange(i * E, min(L, (i + 1) * E)) for i in range(used)]
        v = 6.0 * s.copy()
        ends = np.zeros(used)
        for i, sg in enumerate(segs):
            run = 0.0
            for j in sg:
                run = v[j] + z * run
                v[j] = run
            ends[i] = run
        cplus = v.copy()
        for i, sg in enumerate(segs):
            carry, f = 0.0, 1.0
            for h in range(1, H + 1):
                src = (i - h) % used
                carry += f * ends[src]
                f *= z ** len(segs[src])
            cplus[sg] += carry * z ** (np.arange(len(sg)) + 1)
        v = cplus.copy()
        for i, sg in enumerate(segs):
            run = 0.0
            for j in sg[::-1]:
                run = z * (run - v[j])
                v[j] = run
            ends[i] = run
        out = v.copy()
        for i, sg in enumerate(segs):
            carry, f = 0.0, 1.0
            for h in range(1, H + 1):
                src = (i + h) % used
                carry += f * ends[src]
                f *= z ** len(segs[src])
            out[sg] += carry * z ** (len(sg) - np.arange(len(sg)))
        ref = ndimage.spline_filter1d(s, order=3, mode='grid-wrap')
        assert np.abs(out - ref).max() < 1e-9 * np.abs(ref).max(), (L, E)


def test_dropin_shims_resolve():
    import importlib
    import sys
    d = os.path.join(ROOT, 'manifoldem_python_b200', 'dropin')
    sys.path.insert(0, d)
    try:
        for name in ('getDistanceCTF_local_Conj9combinedS2', 'GetDistancesS2', 'DMembeddingII'):
            sys.modules.pop(name, None)
            m = importlib.import_module(name)
            assert callable(m.op) and m.__file__.startswith(d)
            sys.modules.pop(name, None)
    finally:
        sys.path.remove(d)


def test_fused_lowpass_algebra_model():
    """NumPy model of what lowpass.cu does at N = 256: (1) two real rows as one complex FFT and the separation of their
    half spectra, (2) moments taken on offset-shifted pixels outside the disc and the normalisation applied in
    Fourier space (scale + DC term), (3) the rebuilt full spectrum of a row pair for the inverse transform.
    The result must equal the straightforward (x - mean(b)) / std(b) -> fft2 * G -> ifft2 of the reference (:276-293)."""
    rng = np.random.default_rng(3)
    N = 32
    x = (1000.0 + 10.0 * rng.standard_normal((N, N)))            # un-normalised particle with a large offset
    yy, xx = np.mgrid[:N, :N]
    msk = ((yy - N / 2 + 1) ** 2 + (xx - N / 2) ** 2) < (N / 2) ** 2    # annularMask.py:24-30
    G = rng.random((N, N // 2 + 1))                               # any real filter table in half-spectrum layout
    G[0, 0] = 1.0 / (N * N)
    # reference order of operations
    b = x * (1 - msk)
    xn = (x - b.mean()) / b.std()
    ref = np.fft.irfft2(np.fft.rfft2(xn) * G, s=(N, N)) * (N * N)   # cuFFT-style unnormalised inverse
    # (1) row pairs
    off = x[0, 0]
    xs = x - off
    spec = np.empty((N, N // 2 + 1), complex)
    for r in range(0, N, 2):
        Z = np.fft.fft(xs[r] + 1j * xs[r + 1])
        Zn = np.conj(Z[(-np.arange(N // 2 + 1)) % N])
        spec[r] = 0.5 * (Z[:N // 2 + 1] + Zn)
        spec[r + 1] = -0.5j * (Z[:N // 2 + 1] - Zn)
    assert np.allclose(spec, np.fft.rfft(xs, axis=1), atol=1e-9)
    # (2) moments on shifted values, outside pixels only; undo the offset in the sums
    out = ~msk
    s1, s2, cnt = xs[out].sum(), (xs[out] ** 2).sum(), out.sum()
    S1, S2 = s1 + off * cnt, s2 + 2 * off * s1 + off * off * cnt
    mean = S1 / (N * N)
    inv = 1.0 / np.sqrt(S2 / (N * N) - mean * mean)
    assert np.isclose(mean, b.mean()) and np.isclose(inv, 1.0 / b.std())
    F = np.fft.fft(spec, axis=0)                                   # column pass
    F[0, 0] -= (mean - off) * N * N                                # what is left of the mean after the offset
    F *= G * inv
    colback = np.fft.ifft(F, axis=0) * N                           # unnormalised inverse column pass
    # (3) rebuild z of a row pair and invert
    got = np.empty((N, N))
    k = np.arange(N // 2 + 1)
    for r in range(0, N, 2):
        X1, X2 = colback[r].copy(), colback[r + 1].copy()
        X1[[0, N // 2]] = X1[[0, N // 2]].real                     # a C2R ignores these imaginary parts
        X2[[0, N // 2]] = X2[[0, N // 2]].real
        Z = np.empty(N, complex)
        Z[k] = X1 + 1j * X2
        Z[(N - k[1:-1])] = np.conj(X1[1:-1]) + 1j * np.conj(X2[1:-1])
        z = np.fft.ifft(Z) * N
        got[r], got[r + 1] = z.real, z.imag
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize('N', [16, 15])
def test_contraction_operand_algebra_model(N):
    """NumPy model of the operand layout of DESIGN.md §3 against the reference formula (:391-397)
    D = |C|^2 (|F|^2)^T + transpose - 2 Re(A A^H), A = conj(C) F over the full N x N spectrum:
    radial binning of the first term, Hermitian half spectrum with weight 2 for the second, the self-conjugate
    pixels carried in the swap segments, and the common component F - C M removed from every image."""
    rng = np.random.default_rng(N)
    nS = 7
    imgs = rng.standard_normal((nS, N, N))
    F = np.fft.fft2(imgs)
    f = np.fft.fftfreq(N, 1.0 / N).round().astype(int)
    r2 = f[:, None] ** 2 + f[None, :] ** 2
    df = rng.uniform(1.0, 3.0, nS)
    C = np.sin(0.05 * r2[None] * df[:, None, None]) - 0.1 * np.cos(0.05 * r2[None] * df[:, None, None])   # even, radial
    Cf, Ff = C.reshape(nS, -1), F.reshape(nS, -1)
    A = np.conj(Cf) * Ff
    T1 = (np.abs(Cf) ** 2) @ (np.abs(Ff) ** 2).T
    D_ref = T1 + T1.T - 2 * np.real(A @ np.conj(A).T)
    # half spectrum, common component
    Nh = N // 2 + 1
    Fh, Ch, r2h = F[:, :, :Nh], C[:, :, :Nh], r2[:, :Nh]
    M = (Ch * Fh).sum(0) / (Ch ** 2).sum(0)
    Gh = Fh - Ch * M
    even = N % 2 == 0
    selfcol = np.zeros(Nh, bool)
    selfcol[0] = True
    if even:
        selfcol[N // 2] = True
    w = np.where(selfcol, 1.0, 2.0)[None, :] * np.ones((N, 1))
    bins, inv = np.unique(r2h.ravel(), return_inverse=True)
    Cb = np.stack([np.sin(0.05 * bins * d) - 0.1 * np.cos(0.05 * bins * d) for d in df])
    P = np.zeros((nS, len(bins)))
    for i in range(nS):
        np.add.at(P[i], inv, (w * np.abs(Gh[i]) ** 2).ravel())
    S1, S2 = Cb ** 2 / 4, P
    # S3: one representative per conjugate pair (weight 2), (Re, Im) of A; self-conjugate pixels ride in S1/S2 as -x/4, x
    rep = np.zeros((N, Nh), bool)
    special = []
    for ky in range(N):
        for kx in range(Nh):
            if not selfcol[kx]:
                rep[ky, kx] = True
            else:
                selfrow = ky == 0 or (even and ky == N // 2)
                if selfrow:
                    special.append((ky, kx))
                elif ky < (N + 1) // 2:
                    rep[ky, kx] = True
    Ah = Ch * Gh
    S3 = np.concatenate([Ah[:, rep].real, Ah[:, rep].imag], axis=1)
    xs = np.stack([(Ch[:, ky, kx] * Gh[:, ky, kx]).real for ky, kx in special], axis=1)
    S1 = np.concatenate([S1, -xs / 4], axis=1)
    S2 = np.concatenate([S2, xs], axis=1)
    D = 4 * (S1 @ S2.T + S2 @ S1.T - S3 @ S3.T)
    off = ~np.eye(nS, dtype=bool)
    assert np.abs(D - D_ref)[off].max() < 1e-9 * D_ref[off].max()
    assert np.abs(np.diag(D)).max() < 1e-9 * D_ref[off].max()


def test_knn_selection_model_matches_a_full_sort():
    """The radix selection of k_knn_select (csrc/dm.cu), restated thread for thread on the CPU
    (tests/tools/knn_select_model.py): same lists as sorting the whole row by (value, index) — threshold ties,
    signed zeros, negative values, k = 1 and k = nS included, float32 and float64 keys."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'tools'))
    import knn_select_model as m
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        for nS, k, kind in [(40, 7, 'ties'), (300, 40, 'rand'), (600, 100, 'int'), (777, 1, 'rand'), (513, 128, 'neg'),
                            (64, 16, 'zeros'), (33, 33, 'int')]:
            if kind == 'ties':
                D = np.full((nS, nS), 5.0)
            elif kind == 'rand':
                D = rng.random((nS, nS)) * 1e3
            elif kind == 'int':
                D = rng.integers(0, 20, (nS, nS)).astype(float)
            elif kind == 'neg':
                D = rng.standard_normal((nS, nS))
            else:
                D = np.where(rng.random((nS, nS)) < 0.5, 0.0, -0.0)
            D = D.astype(dt)
            for i in (0, nS // 2, nS - 1):
                idx, val = m.select_row(D[i], i, k)
                idx_r, val_r = m.reference_row(D[i], i, k)
                assert np.array_equal(idx, idx_r) and np.array_equal(val, val_r), (dt, nS, k, kind, i)


def test_sidecar_record_layout(tmp_path):
    """heavy arrays in float32 .npy sidecars, promoted to the reference's float64 when (and only
    when) a consumer reads the key; values identical to the 'pickle' layout; manifest written last, atomically."""
    from manifoldem_python_b200 import myio
    rng = np.random.default_rng(0)
    nS, N = 40, 32
    rec = dict(D=rng.random((nS, nS)).astype(np.float32), ind=np.arange(nS), imgAll=rng.random((nS, N, N)).astype(np.float32),
               CTF=rng.random((nS, N * N)), msk2=1, PD=rng.random(3), imgAvg=rng.random((4, 4)).astype(np.float32),
               version='v', options=dict(a=1), imgAllFlip=None)
    promote = {k: np.float64 for k, v in rec.items() if isinstance(v, np.ndarray) and v.dtype == np.float32}
    f_side, f_pick = str(tmp_path / 'side_prD_0'), str(tmp_path / 'pick_prD_0')
    myio.fout1(f_side, list(rec), list(rec.values()), layout='sidecar', promote=promote)
    myio.fout1(f_pick, list(rec), [v.astype(np.float64) if k in promote else v for k, v in rec.items()], layout='pickle')
    files = sorted(os.listdir(tmp_path))
    assert 'side_prD_0.imgAll.npy' in files and 'side_prD_0.CTF.npy' in files and 'side_prD_0.tmp' not in files
    assert 'side_prD_0.D.npy' not in files                    # 6.4 KB: below the sidecar threshold, stays inside
    assert os.path.getsize(f_side) < 40000
    a, b = myio.fin1(f_side), myio.fin1(f_pick)
    assert isinstance(a, myio.Record) and type(b) is dict
    assert list(a.keys()) != [] and set(a.keys()) == set(b.keys()) and len(a) == len(b) and 'imgAll' in a
    assert a._lazy.keys() == {'imgAll', 'CTF'}                # nothing heavy read yet
    assert a['D'].dtype == np.float64 and np.array_equal(a['D'], b['D'])
    assert a._lazy.keys() == {'imgAll', 'CTF'}
    raw = a.raw('imgAll')
    assert raw.dtype == np.float32 and not raw.flags.writeable
    img = a['imgAll']
    assert img.dtype == np.float64 and img.flags.writeable and np.array_equal(img, b['imgAll'])
    assert a['imgAll'] is img and a._lazy.keys() == {'CTF'}
    assert a.get('CTF').dtype == np.float64 and np.array_equal(a['CTF'], b['CTF'])
    assert a.get('nope', 5) == 5 and a['imgAllFlip'] is None and a['msk2'] == 1 and a['options'] == dict(a=1)
    assert a['imgAvg'].dtype == np.float64 and np.array_equal(a['imgAvg'], b['imgAvg'])
    c = myio.fin1(f_side)
    for (k, v), (k2, v2) in zip(sorted(c.items(), key=lambda t: t[0]), sorted(b.items(), key=lambda t: t[0])):
        assert k == k2 and (np.array_equal(v, v2) if isinstance(v2, np.ndarray) else v == v2), k
    assert pickle.loads(pickle.dumps(myio.fin1(f_side))).keys() == b.keys()
    for plain in (dict(myio.fin1(f_side)), {**myio.fin1(f_side)}, myio.fin1(f_side).copy()):   # copies see the arrays too
        assert type(plain) is dict and np.array_equal(plain['imgAll'], b['imgAll']) and np.array_equal(plain['CTF'], b['CTF'])
    # mutation paths: a key that is written, popped or deleted stops being lazy
    m = myio.fin1(f_side)
    m['imgAll'] = 7
    assert m['imgAll'] == 7 and 'imgAll' not in m._lazy
    ctf = m.pop('CTF')
    assert ctf.dtype == np.float64 and np.array_equal(ctf, b['CTF']) and 'CTF' not in m and m.pop('CTF', 3) == 3
    m = myio.fin1(f_side)
    del m['imgAll']
    assert 'imgAll' not in m and 'imgAll' not in m._lazy
    assert np.array_equal(m.setdefault('CTF', 0), b['CTF']) and m.setdefault('new', 4) == 4
    m.update(CTF=1, other=2)
    assert m['CTF'] == 1 and m['other'] == 2 and not m._lazy
    m = myio.fin1(f_side)
    assert '<lazy>' in repr(m) and all(v is not None or k == 'imgAllFlip' for k, v in zip(m.keys(), m.values()))
    # rewriting a record with fewer heavy arrays removes the stale sidecars
    f_re = str(tmp_path / 're_prD_1')
    myio.fout1(f_re, ['A', 'B'], [rec['imgAll'], rec['CTF']], layout='sidecar')
    myio.fout1(f_re, ['A'], [rec['imgAll']], layout='sidecar')
    assert 're_prD_1.B.npy' not in os.listdir(tmp_path) and 're_prD_1.A.npy' in os.listdir(tmp_path)
    # a missing / truncated sidecar or an unreadable manifest gives None, like the reference's 'None on any failure'
    with open(str(tmp_path / 'side_prD_0.imgAll.npy'), 'r+b') as fh:
        fh.truncate(1000)
    assert myio.fin1(f_side) is None
    os.remove(str(tmp_path / 'side_prD_0.CTF.npy'))
    assert myio.fin1(f_side) is None
    with pytest.raises(ValueError):
        myio.fout1(f_side, ['a'], [1], layout='hdf5')
    # layout selection: argument > p.record_layout > environment > 'pickle'
    from manifoldem_python_b200 import p
    assert myio.default_layout() == 'pickle'
    os.environ['MANIFOLDEM_B200_RECORD'] = 'sidecar'
    try:
        assert myio.default_layout() == 'sidecar'
        p.record_layout = 'pickle'
        assert myio.default_layout() == 'pickle'
    finally:
        del os.environ['MANIFOLDEM_B200_RECORD']
        del p.record_layout


def test_worker_writes_both_layouts_without_touching_the_gpu(tmp_path, monkeypatch):
    """The per-PD worker's record logic (keys, dtypes, marker after the dump) with the device call replaced by a
    stub: the 'sidecar' record read back through myio equals the 'pickle' record of the same results."""
    from manifoldem_python_b200 import getDistanceCTF_local_Conj9combinedS2 as worker
    from manifoldem_python_b200 import myio, p, pd_stage
    p.init()
    nS, N = 12, 64
    p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast, p.mask_vol_file = N, 1.0, 2.0, 300.0, 0.1, ''
    p.dist_prog = str(tmp_path / 'prog') + os.sep
    os.makedirs(p.dist_prog)
    stackf = tmp_path / 'stack.dat'
    np.zeros((nS, N, N), np.float32).tofile(stackf)
    rng = np.random.default_rng(3)
    f32 = dict(D=rng.random((nS, nS)), imgAll=rng.random((nS, N, N)), imgAllFlip=rng.random((nS, N, N)),
               imgAvg=rng.random((N, N)), imgAvgFlip=rng.random((N, N)), imgAllIntensity=rng.random((N, N)))
    f32 = {k: v.astype(np.float32) for k, v in f32.items()}
    ctf = rng.random((nS, N * N))
    seen = {}

    def fake_run_pd(ind, q, df, stack, nStot, Nn, *a, float64=True, **kw):
        seen['float64'] = float64
        cast = (lambda x: x.astype(np.float64)) if float64 else (lambda x: x)
        res = {k: cast(v) for k, v in f32.items()}
        res.update(ind=ind, q=q, df=df, CTF=ctf, msk2=1, PD=np.ones(3), PDs=np.ones((3, nS)), Psis=np.zeros((nS, 1)),
                   imgLabels=np.ones(nS, int), Dnom=np.ones((nS, 1)), Nom=np.ones((nS, 1)), version=pd_stage.VERSION)
        for k in ('imgAll', 'imgAllFlip', 'CTF'):              # like run_pd: heavy arrays not asked for are None
            if k not in kw['fields']:
                res[k] = None
        return res
    monkeypatch.setattr(pd_stage, 'run_pd', fake_run_pd)
    import types
    monkeypatch.setattr(worker, '_Slot', lambda dev: types.SimpleNamespace(ctx=None, arena=None))
    q = np.tile(np.array([[1.0], [0.0], [0.0], [0.0]]), (1, nS))
    opts = dict(verbose=False, avgOnly=False, visual=False, parallel=False, relion_data=False, thres=2000)
    recs = {}
    p.record_virtual_ctf = False                             # the CTF field stored like the other arrays
    for prD, layout in enumerate(('pickle', 'sidecar')):
        p.record_layout = layout
        out = str(tmp_path / ('IMGs_prD_%d' % prD))
        worker.op([np.arange(nS), q, np.full(nS, 1e4), out, prD], dict(type='Butter', Qc=0.5, N=8), str(stackf),
                  (np.zeros(nS), np.zeros(nS)), 2 * nS, opts)
        assert seen['float64'] == (layout == 'pickle')
        assert os.path.exists(os.path.join(p.dist_prog, str(prD)))
        recs[layout] = myio.fin1(out)
    # default sidecar record: the CTF field is virtual (df + microscope constants in the manifest, no file), and
    # p.record_skip drops arrays no consumer reads
    del p.record_virtual_ctf
    p.record_skip = ('imgAllFlip',)
    out = str(tmp_path / 'IMGs_prD_2')
    worker.op([np.arange(nS), q, np.full(nS, 1e4), out, 2], dict(type='Butter', Qc=0.5, N=8), str(stackf),
              (np.zeros(nS), np.zeros(nS)), 2 * nS, opts)
    del p.record_skip, p.record_layout
    c = myio.fin1(out)
    assert list(c.keys()) == worker._KEYS and c['imgAllFlip'] is None
    assert c._lazy['CTF'] == dict(virtual='ctf', df_key='df', N=N, pix_size=1.0, Cs=2.0, EkV=300.0, gaussEnv=np.inf,
                                  AmpContrast=0.1, shape=(nS, N * N))
    assert not os.path.exists(out + '.CTF.npy') and not os.path.exists(out + '.imgAllFlip.npy')
    assert os.path.exists(out + '.imgAll.npy') and np.array_equal(c['df'], np.full(nS, 1e4))
    a, b = recs['sidecar'], recs['pickle']
    assert isinstance(a, myio.Record) and type(b) is dict and list(a.keys()) == list(b.keys()) == worker._KEYS
    assert os.path.getsize(str(tmp_path / 'IMGs_prD_1.imgAll.npy')) < 0.51 * nS * N * N * 8 + 200
    for k in worker._KEYS:
        va, vb = a[k], b[k]
        if isinstance(vb, np.ndarray):
            assert va.dtype == vb.dtype and np.array_equal(va, vb), k
        else:
            assert va == vb, k


def test_ferguson_sorted_chunk_model():
    """k_ferguson_sorted (csrc/dm.cu) sorts the CTA's entries and cuts the sorted chunk, per tile of 32 eps, into
    [saturated: d2/(2 eps) < 2^-54 for the tile's largest 1/(2 eps) -> counted as 1 | degree-4 Taylor: < 1e-3 |
    exp | past the cut for the tile's smallest 1/(2 eps) -> skipped] with three binary searches.  NumPy restatement
    of exactly that rule against the plain definition (fergusonE.py:36-43)."""
    rng = np.random.default_rng(4)
    d2 = np.concatenate([10.0 ** rng.uniform(-3, 9, 500), [0.0, 0.0, 1e-30, 1e300], -np.ones(7)])   # -1 = absent
    logEps = np.arange(-150, 150.2, 0.2)
    s = 1.0 / (2.0 * np.exp(logEps))
    sv = np.sort(np.where(d2 >= 0, d2, np.inf))
    tiny = 2.0 ** -54
    for thr in (10.0, 37.5, np.inf, 1e-6):
        present = d2[d2 >= 0]
        x = present[:, None] * s[None, :]
        with np.errstate(over='ignore'):
            direct = np.where(x < thr, np.exp(-x), 0.0).sum(0)
        tiled = np.zeros_like(direct)
        shortcuts = thr >= 1e-3
        n_exp_calls = 0
        for t0 in range(0, len(s), 32):
            st = s[t0:t0 + 32]
            smax, smin = st.max(), st.min()
            with np.errstate(invalid='ignore', over='ignore'):
                n_sat = int((sv * smax < tiny).sum()) if shortcuts else 0      # monotone in d: a prefix of the chunk
                n_poly = max(int((sv * smax < 1e-3).sum()) if shortcuts else 0, n_sat)
                n_exp = max(int((sv * smin < thr).sum()), n_poly)
                assert (np.diff((sv * smin < thr).astype(int)) <= 0).all()
            acc = np.zeros(len(st))
            for d in sv[n_sat:n_poly]:
                xt = d * st
                acc += 1 + xt * (-1 + xt * (0.5 + xt * (-1 / 6 + xt / 24)))
            for d in sv[n_poly:n_exp]:
                xt = d * st
                acc += np.where(xt < thr, np.exp(-xt), 0.0)
                n_exp_calls += len(st)
            tiled[t0:t0 + 32] = acc + n_sat
        assert np.allclose(tiled, direct, rtol=1e-13, atol=0), thr
        if thr == 10.0:
            assert n_exp_calls < 0.08 * present.size * len(s)          # the point of the exercise


def test_host_arena_reuses_buffers(monkeypatch):
    """pd_stage.HostArena: one buffer per name, reused while it is big enough, regrown otherwise; falls back to
    pageable memory when pinned memory cannot be had (here: no GPU, cudaMallocHost fails)."""
    from manifoldem_python_b200 import pd_stage
    arena = pd_stage.HostArena()
    a = arena.get('raw', (10, 16), np.float32)
    a[...] = 3.0
    b = arena.get('raw', (5, 16), np.float32)
    assert b.shape == (5, 16) and b.dtype == np.float32 and np.shares_memory(a, b) and (b == 3.0).all()
    c = arena.get('raw', (40, 16), np.float64)
    assert c.shape == (40, 16) and c.dtype == np.float64 and not np.shares_memory(a, c)
    d = arena.get('D', (4, 4), np.float32)
    assert not np.shares_memory(c, d)
    c[...] = 1.0
    d[...] = 2.0
    assert (c == 1.0).all() and (d == 2.0).all()
    arena.close()


def test_header_is_plain_c():
    """include/manifoldem_b200.h is the C ABI a maintainer binds: it must compile as C99 on its own."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    hdr = os.path.join(ROOT, 'include', 'manifoldem_b200.h')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-fsyntax-only', '-x', 'c', hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr


def test_lanczos_host_selection_against_dense_eigh():
    """Host half of eigsh_device (DMembeddingII.lanczos_ritz_selection): a NumPy Lanczos with full re-orthogonalisation
    stands in for the device kernels (same recurrence, same coefficient layout); the selection must report convergence
    only when the residual estimates allow it, pick the largest-magnitude pairs, and treat j = nS / a vanishing beta as exact."""
    from manifoldem_python_b200.DMembeddingII import lanczos_ritz_selection
    rng = np.random.default_rng(3)
    nS, k = 60, 6
    Q, _ = np.linalg.qr(rng.standard_normal((nS, nS)))
    lam = np.concatenate([[1.0, -0.9, 0.8, 0.7, 0.65, 0.6], 0.3 * rng.uniform(-1, 1, nS - 6)])
    L = (Q * lam) @ Q.T
    L = (L + L.T) / 2

    def lanczos(L, steps, v0):
        n = L.shape[0]
        V = np.zeros((steps + 1, n))
        alpha, beta = np.zeros(steps), np.zeros(steps + 1)
        V[0] = v0 / np.linalg.norm(v0)
        for j in range(steps):
            w = L @ V[j]
            for _ in range(2):
                h = V[:j + 1] @ w
                alpha[j] += h[j]
                w = w - V[:j + 1].T @ h
            beta[j + 1] = np.linalg.norm(w)
            V[j + 1] = w / beta[j + 1] if beta[j + 1] > 0 else 0
        return V, alpha, beta

    V, alpha, beta = lanczos(L, 12, rng.standard_normal(nS))
    theta, S, j_eff, done = lanczos_ritz_selection(alpha, beta, k, nS, 1e-12)
    assert not done and j_eff == 12 and theta.shape == (k,)
    V, alpha, beta = lanczos(L, nS, rng.standard_normal(nS))
    theta, S, j_eff, done = lanczos_ritz_selection(alpha, beta, k, nS, 1e-12)
    assert done and np.allclose(np.sort(theta), np.sort(lam[:6]), atol=1e-10)
    X = V[:j_eff].T @ S
    assert np.abs(L @ X - X * theta).max() < 1e-9
    # invariant subspace: a start vector inside the span of 4 eigenvectors breaks down after 4 steps (exact pairs, k' = 4)
    v0 = Q[:, :4] @ np.array([1.0, 2.0, -1.0, 0.5])
    V, alpha, beta = lanczos(L, 8, v0)
    theta, S, j_eff, done = lanczos_ritz_selection(alpha, beta, k, nS, 1e-12)
    assert done and j_eff == 4 and np.allclose(np.sort(theta), np.sort(lam[:4]), atol=1e-9)


def test_vectorised_manifold_fit_equals_the_per_point_restatement(golden_dir):
    """fit_1D_open_manifold_3D drop-in (one batched eigvals per iteration) vs the oracle's per-point np.roots loop, on
    the reference's own psirec outputs and on random curves (incl. a degenerate a_3 = 0 start)."""
    from manifoldem_python_b200 import fit_1D_open_manifold_3D as fit
    from oracle import nlsa as onl
    g = np.load(os.path.join(golden_dir, 'nlsa_nS80_N24.npz'))
    cases = [g['m1_psi0_psirec'], g['md_psi1_psirec']]
    rng = np.random.default_rng(2)
    t = rng.uniform(0, 1, 120)
    cases.append(np.stack([np.cos(np.pi * t), 0.5 * np.cos(2 * np.pi * t), 0.2 * np.cos(3 * np.pi * t)], 1)
                 + 0.01 * rng.standard_normal((120, 3)))
    for psi in cases:
        a, b, tau = fit.op_host(psi)
        a0, b0, tau0 = onl.fit_1d_open_manifold_3d(psi)
        assert np.allclose(a, a0, rtol=1e-12, atol=0) and np.allclose(b, b0, rtol=1e-12, atol=1e-15)
        assert tau.shape == tau0.shape and np.abs(tau - tau0).max() < 1e-12
    ref_tau = g['m1_psi0_tau']
    assert np.abs(fit.op_host(g['m1_psi0_psirec'])[2] - ref_tau).max() < 1e-9
    x = np.stack([np.cos(np.pi * t), 0.5 * np.cos(2 * np.pi * t), np.zeros_like(t)], 1)
    assert np.array_equal(fit._taus(x, np.array([1.0, 0.5, 0.0]), np.zeros(3)),
                          np.array([onl._tau_of_point(x[p], np.array([1.0, 0.5, 0.0]), np.zeros(3)) for p in range(120)]))


def test_psi_analysis_driver_bookkeeping(tmp_path):
    """psiAnalysis.divid / fileCheck (modules/psiAnalysis.py:20-50): markers '<prD>_<psi>' under p.psi2_prog mark finished
    (PD, psi) pairs; a job lists only the psis still to do."""
    from manifoldem_python_b200 import psiAnalysis, p
    p.init()
    p.psi2_prog = str(tmp_path / 'prog')
    os.makedirs(p.psi2_prog)
    p.num_psis, p.numberofJobs = 3, 4
    p.dist_file, p.psi_file, p.psi2_file, p.EL_file = 'D_', 'P_', 'Q_', 'E_'
    for name in ('0_0', '0_1', '0_2', '2_1', '.hidden'):
        open(os.path.join(p.psi2_prog, name), 'a').close()
    fin = psiAnalysis.fileCheck(4)
    assert fin.shape == (4, 3) and fin.sum() == 4 and fin[0].all() and fin[2, 1] == 1
    rc = dict(psiNumsAll=np.tile(np.arange(3), (4, 1)), sensesAll=np.ones((4, 3)))
    jobs = psiAnalysis.divid(4, rc, fin)
    assert [j[7] for j in jobs] == [[], [0, 1, 2], [0, 2], [0, 1, 2]]
    assert jobs[2][:4] == ['D_prD_2', 'P_prD_2', 'Q_prD_2', 'E_prD_2'] and jobs[3][6] == 3


def test_partition_with_rank_speeds():
    """LPT with per-rank speeds: ranks 4-7 one and a half times as fast get one and a half times the PDs; every job once."""
    from manifoldem_python_b200 import partition
    costs = [1.0] * 1000
    speeds = [1.0] * 4 + [1.5] * 4
    shards = partition.lpt_partition(costs, 8, speeds)
    assert sorted(i for s in shards for i in s) == list(range(1000))
    n = [len(s) for s in shards]
    assert all(abs(x - 100) <= 1 for x in n[:4]) and all(abs(x - 150) <= 1 for x in n[4:])
    assert partition.imbalance(costs, shards, speeds) < 1.02
    assert partition.counts_by_speed(384, [23.3] * 4 + [35.5] * 4) == [38] * 4 + [58] * 4
    assert sum(partition.counts_by_speed(100, [1, 2, 3])) == 100
    rng = np.random.default_rng(0)
    costs = list(rng.uniform(1, 5, 57))
    assert partition.lpt_partition(costs, 3) == partition.lpt_partition(costs, 3, [1, 1, 1])


def test_manifold_analysis_driver_bookkeeping(tmp_path):
    """manifoldAnalysis.divide / fileCheck / count (modules/manifoldAnalysis.py:30-54): one job per PD without a marker under
    p.psi_prog, the eigenvalue file under out_dir/topos/PrD_<prD + 1>/."""
    from manifoldem_python_b200 import manifoldAnalysis, p
    p.init()
    p.user_dir, p.proj_name = str(tmp_path), 'ma'
    p.create_dir()
    p.numberofJobs = 5
    for name in ('1', '3', '.hidden'):
        open(os.path.join(p.psi_prog, name), 'a').close()
    assert manifoldAnalysis.fileCheck() == [1, 3] and manifoldAnalysis.count(5) == 3
    jobs = manifoldAnalysis.divide(5)
    assert [j[3] for j in jobs] == [0, 2, 4]
    assert jobs[1][0] == p.dist_file + 'prD_2' and jobs[1][1] == p.psi_file + 'prD_2'
    assert jobs[2][2] == '{}/topos/PrD_5/eig_spec.txt'.format(p.out_dir)


def _queue_consumer(queue, out, tag):
    """Spawned stand-in for a GPU worker: the queue loop of the driver with a recording job function."""
    from manifoldem_python_b200 import GetDistancesS2
    import time as _t

    def rec(job, *a):
        _t.sleep(0.002)
        out.put((tag, job[4]))
    n = GetDistancesS2._run_queue(queue, None, None, None, 0, {}, 1e9, op=rec)
    out.put((tag, -1 - n))


def test_job_queue_hands_every_pd_out_once():
    """The box-wide job queue of GetDistancesS2 (the reference's imap_unordered, :110-113): two consumer processes with two
    slots each drain it, every PD runs exactly once and every slot sees an end marker (both processes exit)."""
    import multiprocessing
    from manifoldem_python_b200 import GetDistancesS2
    ctx = multiprocessing.get_context('spawn')
    queue, out = ctx.Queue(), ctx.Queue()
    n_jobs, n_proc = 60, 2
    for prD in range(n_jobs):
        queue.put([np.arange(3), None, None, 'f%d' % prD, prD])
    for _ in range(n_proc * GetDistancesS2._inflight(1e9)):
        queue.put(None)
    procs = [ctx.Process(target=_queue_consumer, args=(queue, out, t)) for t in range(n_proc)]
    [pr.start() for pr in procs]
    got = [out.get(timeout=120) for _ in range(n_jobs + n_proc)]
    [pr.join() for pr in procs]
    assert all(pr.exitcode == 0 for pr in procs)
    ran = sorted(j for _, j in got if j >= 0)
    assert ran == list(range(n_jobs))
    counts = {t: -1 - j for t, j in got if j < 0}
    assert sum(counts.values()) == n_jobs and sorted(counts) == [0, 1]
