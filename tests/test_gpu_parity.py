"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs,
against the committed golden vectors, and through size-independent properties at larger sizes.

Tolerances (north_star): off-diagonal D within 1e-5 relative of the float64 reference; diagonal
|D_ii| <= 1e-5 * max D; images / averages within 2e-6 of their max (fp32 pipeline)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

D_RTOL = 1e-5
IMG_TOL = 2e-6


@pytest.fixture(scope='module')
def env():
    from manifoldem_python_b200 import _lib, pd_stage, synthetic
    return _lib, pd_stage, synthetic


def _oracle(pd, N, **kw):
    from oracle import pd_distance as opd
    em = pd['em']
    return opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                           em['EkV'], em['AmpContrast'], **kw)


def _gpu(pd_stage, pd, N, **kw):
    em = pd['em']
    return pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                           em['EkV'], em['AmpContrast'], **kw)


def _check_D(D, Dr, rtol=D_RTOL):
    n = D.shape[0]
    off = ~np.eye(n, dtype=bool)
    rel = np.abs(D - Dr)[off] / Dr[off]
    assert rel.max() <= rtol, 'max rel err %.3e' % rel.max()
    assert np.abs(np.diag(D) - np.diag(Dr)).max() <= 1e-5 * Dr.max()
    assert np.array_equal(D, D.T)


def _check_fields(res, ref, tol=IMG_TOL):
    for k in ('imgAll', 'imgAllFlip', 'imgAvg', 'imgAvgFlip', 'imgAllIntensity'):
        a, b = res[k].reshape(ref[k].shape), ref[k]
        assert np.abs(a - b).max() <= tol * np.abs(b).max(), (k, np.abs(a - b).max() / np.abs(b).max())
    assert np.abs(res['CTF'].reshape(ref['CTF'].shape) - ref['CTF']).max() < 1e-11
    for k in ('PD', 'PDs', 'Psis', 'Dnom', 'Nom', 'imgLabels'):
        assert np.array_equal(np.asarray(res[k]), np.asarray(ref[k])), k


@pytest.mark.parametrize('name', ['pd_spider_N32.npz', 'pd_spider_N25_lownoise.npz'])
def test_golden_reference_vectors(env, golden_dir, name):
    """CUDA path vs outputs of the reference itself (fixtures from tests/golden/make_golden.py)."""
    _lib, pd_stage, _ = env
    g = np.load(os.path.join(golden_dir, name))
    N = int(g['N'])
    res = pd_stage.run_pd(g['ind'], g['q'], g['df'], g['stack'], int(g['nStot']), N, float(g['em_pix_size']),
                          float(g['em_Cs']), float(g['em_EkV']), float(g['em_AmpContrast']))
    ref = {k[4:]: g[k] for k in g.files if k.startswith('ref_')}
    # N<=32: the reference's 3x3-tile rotate still feels its mirror border at 0.268^(0.79N) ~ 1e-9..4e-12
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


def test_golden_volmask(env, golden_dir):
    _lib, pd_stage, _ = env
    g = np.load(os.path.join(golden_dir, 'pd_volmask_N24.npz'))
    N = int(g['N'])
    res = pd_stage.run_pd(g['ind'], g['q'], g['df'], g['stack'], int(g['nStot']), N, float(g['em_pix_size']),
                          float(g['em_Cs']), float(g['em_EkV']), float(g['em_AmpContrast']), msk2=g['ref_msk2'])
    ref = {k[4:]: g[k] for k in g.files if k.startswith('ref_')}
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)
    assert np.array_equal(res['msk2'], g['ref_msk2'])


def test_mask_at_bench_box_size(env):
    """msk2 at N = 256: the fused low-pass column kernel, the second (unmasked) spectrum for the Wiener average and
    the one-pass operand writer with the in-place phase flip, all fields against the float64 oracle."""
    _lib, pd_stage, synthetic = env
    nS, N = 24, 256
    pd = synthetic.make_pd(nS, N, seed=21, snr=0.5)
    yy, xx = np.mgrid[:N, :N]
    msk2 = ((yy - N / 2) ** 2 / (0.42 * N) ** 2 + (xx - N / 2) ** 2 / (0.3 * N) ** 2) < 1
    res = _gpu(pd_stage, pd, N, msk2=msk2)
    ref = _oracle(pd, N, rotate_impl='periodic', msk2=msk2)
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


def test_unnormalised_stack_at_bench_box_size(env):
    """Raw particles with a large offset and scale (1000 +- 10) through the fused ingest: the moments are taken on
    offset-shifted values and the normalisation is applied in Fourier space; same bar as everywhere else."""
    _lib, pd_stage, synthetic = env
    nS, N = 16, 256
    pd = synthetic.make_pd(nS, N, seed=22, snr=0.5)
    pd = dict(pd, stack=(pd['stack'] * 10.0 + 1000.0).astype(np.float32))
    res = _gpu(pd_stage, pd, N)
    ref = _oracle(pd, N, rotate_impl='periodic')
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


def test_golden_relion(env, golden_dir):
    """RELION branch: cubic 'wrap' shift by (shy-0.5, shx-0.5) on the device vs the reference's ndimage.shift."""
    _lib, pd_stage, _ = env
    g = np.load(os.path.join(golden_dir, 'pd_relion_N24.npz'))
    N = int(g['N'])
    stack3d = g['stack'].reshape(-1, N, N)
    res = pd_stage.run_pd(g['ind'], g['q'], g['df'], stack3d, int(g['nStot']), N, float(g['em_pix_size']),
                          float(g['em_Cs']), float(g['em_EkV']), float(g['em_AmpContrast']), relion=True,
                          sh=(g['shx'], g['shy']))
    ref = {k[4:]: g[k] for k in g.files if k.startswith('ref_')}
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


@pytest.mark.parametrize('N', [100, 128, 256, 300])
def test_relion_shift_large_boxes(env, N):
    """RELION branch at box sizes whose mirror extension (2N-2) takes the long-line prefilter kernels."""
    _lib, pd_stage, synthetic = env
    nS = 6
    pd = synthetic.make_pd(nS, N, seed=50 + N, snr=1.0)
    rng = np.random.default_rng(N)
    sh = (rng.uniform(-4, 4, nS), rng.uniform(-4, 4, nS))
    stack3d = pd['stack'].reshape(nS, N, N)
    em = pd['em']
    res = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], stack3d, pd['nStot'], N, em['pix_size'], em['Cs'], em['EkV'],
                          em['AmpContrast'], relion=True, sh=sh)
    ref = _oracle(dict(pd, stack=stack3d), N, relion=True, sh=sh, rotate_impl='periodic')
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


@pytest.mark.parametrize('nS,N,snr,seed', [(150, 64, 0.1, 2), (257, 64, 10.0, 5), (300, 128, 10.0, 3), (129, 96, 0.5, 7),
                                           (24, 320, 0.5, 8), (40, 256, 0.2, 9)])
def test_oracle_parity_tc(env, nS, N, snr, seed):
    """tcgen05 3xTF32 product path vs the float64 oracle, noisy and low-noise (worst cancellation)."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(nS, N, seed=seed, snr=snr)
    res = _gpu(pd_stage, pd, N)
    ref = _oracle(pd, N, rotate_impl='periodic')
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


@pytest.mark.parametrize('N,nS', [(128, 70), (256, 70), (128, 2), (256, 3), (128, 301), (320, 24), (320, 3)])
@pytest.mark.parametrize('relion', [False, True])
def test_own_fft_kernels_match_cufft_path(env, N, nS, relion):
    """The fused ingest / low-pass / a10 kernels of lowpass.cu (N = 128: 16 x 8, N = 256: 16 x 16; N = 320: column pass 20 x 16) against the generic
    kernels + cuFFT (context option cufft_lowpass): the same images to fp32 FFT round-off, the same D to 1e-5.
    SPIDER stacks take the transposing ingest, RELION stacks the row-major one behind the sub-pixel shift."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(nS, N, seed=300 + N, snr=0.5)
    em = pd['em']
    kw = {}
    stack = pd['stack']
    if relion:
        rng = np.random.default_rng(N)
        kw = dict(relion=True, sh=(rng.uniform(-3, 3, nS), rng.uniform(-3, 3, nS)))
        stack = stack.reshape(nS, N, N)
    ctx = _lib.Context(0)
    try:
        out = []
        for generic in (0, 1):
            ctx.set_option('cufft_lowpass', generic)
            out.append(pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], stack, pd['nStot'], N, em['pix_size'], em['Cs'],
                                       em['EkV'], em['AmpContrast'], ctx=ctx, **kw))
    finally:
        ctx.close()
    own, gen = out
    _check_D(own['D'], gen['D'])
    for k in ('imgAll', 'imgAllFlip', 'imgAvg', 'imgAvgFlip', 'imgAllIntensity'):
        assert np.abs(own[k] - gen[k]).max() <= IMG_TOL * np.abs(gen[k]).max(), k


def test_oracle_parity_single_cta_tiles(env):
    """The single-CTA tcgen05 kernel (contraction=2) stays selectable and meets the same bar."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(300, 128, seed=3, snr=10.0)
    res = _gpu(pd_stage, pd, 128, contraction=2, fields=('D',))
    ref = _oracle(pd, 128, rotate_impl='periodic', keep=('D',))
    _check_D(res['D'], ref['D'])


def test_oracle_parity_simt_checker(env):
    """The fp64-accumulate SIMT kernel isolates operand error from tensor-core accumulation error."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(100, 64, seed=9, snr=10.0)
    res = _gpu(pd_stage, pd, 64, contraction=1)
    ref = _oracle(pd, 64, rotate_impl='periodic')
    _check_D(res['D'], ref['D'], rtol=3e-6)


def test_gauss_filter_and_avg_only(env):
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(40, 48, seed=4, snr=0.3)
    fp = dict(type='Gauss', Qc=0.4, N=8)
    res = _gpu(pd_stage, pd, 48, filterPar=fp)
    ref = _oracle(pd, 48, rotate_impl='periodic', filterPar=fp)
    _check_D(res['D'], ref['D'])
    _check_fields(res, ref)
    res = _gpu(pd_stage, pd, 48, avg_only=True)
    assert not res['D'].any()
    with pytest.raises(ValueError):
        _gpu(pd_stage, pd, 48, filterPar=dict(type='Box', Qc=0.4, N=8))


def test_split_and_chunk_invariance(env):
    """K-slicing and chunking change the summation order only: results agree to fp32 round-off."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(260, 64, seed=12, snr=0.2)
    base = _gpu(pd_stage, pd, 64, fields=('D',))['D']
    for chunk, split in ((2, 1), (4, 3), (8, 0)):
        other = _gpu(pd_stage, pd, 64, fields=('D',), k_chunk_blocks=chunk, split_k=split)['D']
        off = ~np.eye(260, dtype=bool)
        assert (np.abs(other - base)[off] / base[off]).max() < 2e-5


def test_invariances_full_size(env):
    """Size-independent properties at BASELINE config-2 size (1000 x 128^2), no oracle needed:
    symmetry, ~zero diagonal, permutation equivariance, and D(i,j)=0 for duplicated particles."""
    _lib, pd_stage, synthetic = env
    nS, N = 1000, 128
    rng = np.random.default_rng(0)
    stack = rng.standard_normal((nS, N, N)).astype(np.float32)
    stack[1] = stack[0]                                   # duplicate particle
    ind = np.arange(nS)
    pd0 = synthetic.make_pd(4, 16, seed=0)                # only for the EM constants
    em = pd0['em']
    q = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS), 1.1 + 0.03 * rng.standard_normal(nS),
                                rng.uniform(0, 2 * np.pi, nS))
    df = rng.uniform(10000, 30000, nS)
    q[:, 1], df[1] = q[:, 0], df[0]
    run = lambda idx: pd_stage.run_pd(idx, q[:, idx], df[idx], stack.reshape(-1), 2 * nS, N, em['pix_size'], em['Cs'],
                                      em['EkV'], em['AmpContrast'], fields=('D',))['D']
    D = run(ind)
    assert np.array_equal(D, D.T)
    assert np.abs(np.diag(D)).max() <= 1e-5 * D.max()
    assert abs(D[0, 1]) <= 1e-5 * D.max()
    assert (D[~np.eye(nS, dtype=bool)] > -1e-5 * D.max()).all()
    # permuting the members permutes D.  (PD and psi_p are means over the same set -> unchanged to round-off.)
    perm = rng.permutation(nS)
    Dp = run(ind[perm])
    ref = D[np.ix_(perm, perm)]
    off = ~np.eye(nS, dtype=bool)
    big = ref[off] > 1e-3 * D.max()
    assert (np.abs(Dp - ref)[off][big] / ref[off][big]).max() < 2e-5


def test_contraction_kernel_vs_fp64(env):
    """mem_contract_device alone: random operands, tcgen05 vs numpy float64 on the same hi+lo values."""
    _lib, _, _ = env
    lib = _lib.load()
    ctx = _lib.default_context()
    rng = np.random.default_rng(3)
    nS, n1, n3 = 300, 5, 37
    K = 32 * (2 * n1 + n3)
    Z = rng.standard_normal((nS, K)).astype(np.float32)
    Z[:, :64 * n1] = np.abs(Z[:, :64 * n1])
    hi = (Z.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = ((Z - hi).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    z = (hi.astype(np.float64) + lo.astype(np.float64))
    S1, S2, S3 = z[:, :32 * n1], z[:, 32 * n1:64 * n1], z[:, 64 * n1:]
    exact = 4 * (S1 @ S2.T + S2 @ S1.T - S3 @ S3.T)
    zhi = _lib.DeviceArray(ctx, Z.shape, np.float32, hi)
    zlo = _lib.DeviceArray(ctx, Z.shape, np.float32, lo)
    Dd = _lib.DeviceArray(ctx, (nS, nS), np.float32)
    shp = _lib.ContractShape(nS=nS, n1_blocks=n1, n3_blocks=n3, ldz=K)
    scale = 4 * (np.abs(S1) @ np.abs(S2).T + np.abs(S2) @ np.abs(S1).T + np.abs(S3) @ np.abs(S3).T)
    for kind, tol in ((1, 2e-7), (0, 2e-6), (2, 2e-6)):    # SIMT checker, CTA-pair tiles, single-CTA tiles
        _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, kind, 0, 0, None))
        ctx.sync()
        D = Dd.download().astype(np.float64)
        iu = np.triu_indices(nS)
        assert (np.abs(D - exact)[iu] / scale[iu]).max() < tol, kind
        assert np.array_equal(D, D.T)


@pytest.mark.parametrize('nS', [1, 2, 5])
def test_tiny_pds(env, nS):
    """Edge cases: PDs far smaller than one 256 x 256 tile (TMA zero-fills the missing rows)."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(nS, 32, seed=40 + nS, snr=1.0)
    res = _gpu(pd_stage, pd, 32)
    ref = _oracle(pd, 32, rotate_impl='periodic')
    assert res['D'].shape == (nS, nS)
    if nS > 1:
        _check_D(res['D'], ref['D'])
    else:
        assert abs(res['D'][0, 0]) <= 1e-5 * max(1.0, np.abs(ref['imgAll']).max() ** 2 * 32 * 32)
    _check_fields(res, ref)


@pytest.mark.parametrize('nS', [1, 3])
def test_tiny_pds_at_bench_box_size(env, nS):
    """The persistent column kernel and the block-per-image row kernels with fewer images than CTAs."""
    _lib, pd_stage, synthetic = env
    pd = synthetic.make_pd(nS, 256, seed=60 + nS, snr=1.0)
    res = _gpu(pd_stage, pd, 256)
    ref = _oracle(pd, 256, rotate_impl='periodic')
    if nS > 1:
        _check_D(res['D'], ref['D'])
    _check_fields(res, ref)


def test_bad_inputs_fail_loudly(env):
    """No silent fallback: unsupported shapes and parameters surface as errors."""
    _lib, pd_stage, synthetic = env
    lib, ctx = _lib.load(), _lib.default_context()
    prm = _lib.PdParams(nS=0, N=64, transposed=1, filter_order=8, filter_Qc=0.5, pix_size=1.0, Cs=2.0, EkV=300.0,
                        gaussEnv=float('inf'), AmpContrast=0.1)
    io = _lib.PdIO()
    assert lib.mem_pd_distance_host(ctx.handle, C.byref(prm), C.byref(io)) != 0        # empty PD
    assert b'bad shape' in lib.mem_last_error()
    pd = synthetic.make_pd(4, 8, seed=1)
    with pytest.raises(RuntimeError):                                                   # N < 16
        _gpu(pd_stage, pd, 8)
    shp = _lib.ContractShape(nS=4, n1_blocks=1, n3_blocks=1, ldz=32)                    # ldz too small
    assert lib.mem_contract_device(ctx.handle, C.byref(shp), None, None, None, 0, 0, 0, None) != 0
    shp = _lib.ContractShape(nS=4, n1_blocks=1, n3_blocks=1, ldz=96)
    assert lib.mem_contract_device(ctx.handle, C.byref(shp), None, None, None, 7, 0, 0, None) != 0   # unknown kind


@pytest.mark.parametrize('N', [128, 256, 320])
def test_reference_vectors_at_baseline_box_sizes(env, golden_dir, N):
    """The CUDA path against outputs of the unmodified reference at the box sizes that take size-specific kernels (FFT-128 /
    256 / 320, rotation kernels for multiples of 32): tests/golden/pd_box_N*.npz (make_golden_boxes.py)."""
    from _box_golden import _box_case, check_box_outputs
    _lib, pd_stage, synthetic = env
    g, pd = _box_case(golden_dir, N)
    res = _gpu(pd_stage, pd, N, fields=('D', 'imgAll'))
    check_box_outputs(res, g, d_rtol=D_RTOL, img_tol=IMG_TOL)
