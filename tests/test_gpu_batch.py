"""GPU tests of the batched distance stage (`mem_pd_distance_batch_device`): a group of reference-sized PDs through ONE
launch sequence + one grouped tcgen05 launch must give, PD by PD, what the one-PD-at-a-time path gives, and match the
float64 oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _group(sizes, N, seed):
    """PDs sharing one SPIDER stack: members of PD g are a random subset of the stack (conjugates included)."""
    from manifoldem_python_b200 import synthetic
    big = synthetic.make_pd_fast(int(sum(sizes)), N, seed=seed, snr=0.5)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(int(sum(sizes)))
    jobs, a = [], 0
    for g, n in enumerate(sizes):
        sel = np.sort(perm[a:a + n])
        a += n
        q = big['q'][:, sel].copy()
        # a different mean in-plane convention per PD: rotate the PD's quaternions so psi_p differs between PDs
        jobs.append((big['ind'][sel], q, big['df'][sel]))
    return big, jobs


@pytest.mark.parametrize('N,sizes', [(64, (117, 260, 33, 450, 300)), (128, (130, 257, 512, 64)), (256, (150, 206, 300))])
def test_batch_equals_one_pd_at_a_time(N, sizes):
    from manifoldem_python_b200 import pd_stage
    big, jobs = _group(sizes, N, seed=3 + N)
    em = big['em']
    args = (big['stack'], big['nStot'], N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast'])
    got = pd_stage.run_pd_batch(jobs, *args, want_imgAll=True)
    assert len(got) == len(jobs)
    worst = 0.0
    for (ind, q, df), h in zip(jobs, got):
        one = pd_stage.run_pd(ind, q, df, *args, fields=('D', 'imgAll'), float64=False)
        assert h['D'].shape == one['D'].shape == (len(ind), len(ind))
        assert np.array_equal(h['imgAll'], one['imgAll'])          # same kernels, same values: the aligned images are identical
        off = ~np.eye(len(ind), dtype=bool)
        rel = np.abs(h['D'] - one['D'])[off] / one['D'][off]
        worst = max(worst, rel.max())
        assert np.allclose(np.diag(h['D']), np.diag(one['D']), atol=1e-5 * one['D'].max())
    # K is cut into other slices than in the single-PD launch: fp32 partial sums are added in another order
    assert worst < 2e-6, worst


def test_batch_against_the_oracle():
    from manifoldem_python_b200 import pd_stage
    from oracle import pd_distance as opd
    N, sizes = 64, (48, 90, 40)
    big, jobs = _group(sizes, N, seed=9)
    em = big['em']
    got = pd_stage.run_pd_batch(jobs, big['stack'], big['nStot'], N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast'])
    for (ind, q, df), h in zip(jobs, got):
        ref = opd.pd_distance(ind, q, df, big['stack'], big['nStot'], N, em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast'],
                              rotate_impl='periodic')
        off = ~np.eye(len(ind), dtype=bool)
        assert (np.abs(h['D'] - ref['D'])[off] / ref['D'][off]).max() < 1e-5


def test_batch_argument_errors():
    import ctypes as C
    from manifoldem_python_b200 import _lib
    lib = _lib.load()
    ctx = _lib.default_context()
    prm = _lib.PdParams(nS=10, N=64, transposed=1, filter_type=0, filter_order=8, filter_Qc=0.5, pix_size=1.0, Cs=2.0, EkV=300.0,
                        gaussEnv=float('inf'), AmpContrast=0.1)
    io = _lib.PdIO()
    start = np.array([0, 4, 9], dtype=np.int32)               # does not end at nS
    pp = np.zeros(2)
    assert lib.mem_pd_distance_batch_device(ctx.handle, C.byref(prm), C.byref(io), 2, start.ctypes.data, pp.ctypes.data, None) != 0
    start[2] = 10
    assert lib.mem_pd_distance_batch_device(ctx.handle, C.byref(prm), C.byref(io), 2, start.ctypes.data, pp.ctypes.data, None) != 0  # no D
