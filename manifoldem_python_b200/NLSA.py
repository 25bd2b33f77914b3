"""Drop-in for the NLSA step of the psi analysis (modules/NLSA.py:23-158).

op(NLSAPar, DD, posPath, posPsi1, imgAll, msk2, CTF, ExtPar)
    -> (IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau)                     same signature / return values as the reference

Device (C ABI, nlsa.cu, float64): ConD, the Wiener-filtered supervector products (taken in Fourier space: one transform per
particle, ConOrder x psiTrunc inverse transforms instead of ConOrder x (num - ConOrder) fft2 / ifft2 pairs), A^T A, U = A V S^-1,
Topo_mean, the rank-2 frame reconstruction, frame normalisation, L2_distance**2; both DMembeddingII calls run through the
device chain of DMembeddingII.embed (kNN, graph, Ferguson sweep, Laplacian, Lanczos).  Host: the psiTrunc x psiTrunc eigh of
svdRF (same np.linalg.eigh call) and the 1-D manifold fit (fit_1D_open_manifold_3D, vectorised).

`PdState` keeps what does not depend on the psi being analysed — D, the CTF-weighted spectra of every particle, the CTF
half planes — on the device, so psiAnalysisParS2 uploads a PD once for all of its psis; `analyse` is op() on that state and
can leave IMGT on the device (only the class representatives are downloaded in the first pass).
"""
import numpy as np

from . import _lib, DMembeddingII, fit_1D_open_manifold_3D
from .getDistanceCTF_local_Conj9combinedS2 import _ctx


class PdState:
    """Device-resident inputs of one PD: D (float32 or float64, any square size), H = rfft2(imgAll) * CTF, CTF half planes."""

    def __init__(self, D, imgAll, CTF, ctx=None):
        self.ctx = ctx or _ctx()
        lib = _lib.load()
        imgAll = np.asarray(imgAll)
        n, N = imgAll.shape[0], imgAll.shape[1]
        CTF = np.asarray(CTF).reshape(n, N, N)
        c0 = CTF[0]
        if not np.allclose(c0, np.roll(c0[::-1, ::-1], (1, 1), (0, 1)), rtol=1e-12, atol=1e-14 * max(1.0, np.abs(c0).max())):
            raise ValueError('NLSA: the CTF planes must be even, CTF(-k) = CTF(k) (every CTF of the distance stage is)')
        self.n, self.N, self.Nh = n, N, N // 2 + 1
        self.D = DMembeddingII.upload(D, self.ctx) if D is not None else None
        img_d = _lib.DeviceArray(self.ctx, (n, N, N), np.float64, np.ascontiguousarray(imgAll, dtype=np.float64))
        ctf_d = _lib.DeviceArray(self.ctx, (n, N, N), np.float64, np.ascontiguousarray(CTF, dtype=np.float64))
        self.H = _lib.DeviceArray(self.ctx, (n, N, self.Nh, 2), np.float64)
        self.Ch = _lib.DeviceArray(self.ctx, (n, N, self.Nh), np.float64)
        _lib.check(lib.mem_nlsa_spectra_device(self.ctx.handle, img_d.ptr, ctf_d.ptr, n, N, self.H.ptr, self.Ch.ptr, None))
        self.ctx.sync()
        img_d.free()
        ctf_d.free()

    def free(self):
        for a in (self.D, self.H, self.Ch):
            if a is not None:
                a.free()
        self.D = self.H = self.Ch = None


def analyse(state, sel_D, sel_img, NLSAPar, msk2, keep_IMGT_on_device=False, timings=None):
    """NLSA.op on a PdState.  sel_D: rows / columns of state.D in snapshot order (DD = D[sel_D][:, sel_D]); sel_img: particle
    index of every snapshot (posPath[posPsi1]).  Returns the reference's 8-tuple; with keep_IMGT_on_device the first entry is
    a `_lib.DeviceArray` [nC][Npix] (frame-major) instead of the (Npix, nC) NumPy array."""
    import time
    lib = _lib.load()
    ctx = state.ctx
    clock = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            ctx.sync()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - clock[0]) * 1e3
            clock[0] = now
    num, ConOrder, k, tune = int(NLSAPar['num']), int(NLSAPar['ConOrder']), int(NLSAPar['k']), NLSAPar['tune']
    nS, psiTrunc = int(NLSAPar['nS']), int(NLSAPar['psiTrunc'])
    N, nI = state.N, num - ConOrder
    if ConOrder < 1 or nI - ConOrder < 1:
        raise ValueError('NLSA needs 1 <= ConOrder and num > 2 ConOrder (num=%d ConOrder=%d)' % (num, ConOrder))
    selD = _lib.DeviceArray(ctx, (num,), np.int32, np.ascontiguousarray(sel_D, dtype=np.int32))
    selI = _lib.DeviceArray(ctx, (num,), np.int32, np.ascontiguousarray(sel_img, dtype=np.int32))
    # ---- :30-45 ConD and its diffusion map
    ConD = _lib.DeviceArray(ctx, (nI, nI), np.float64)
    _lib.check(lib.mem_nlsa_cond_device(ctx.handle, state.D.ptr, state.D.dtype.itemsize, state.D.shape[0], selD.ptr, num, ConOrder,
                                        ConD.ptr, None))
    lap('ConD')
    lambdaC, psiC, sigmaC, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.embed(ConD, k, tune)
    ConD.free()
    lap('embed_ConD')
    psiC1 = np.copy(psiC)
    ell = psiTrunc - 1
    E = ell + 1
    psiC = np.hstack((np.ones((nI, 1)), psiC[:, 0:ell]))                       # :61
    mu_psi = np.ascontiguousarray(mu.reshape((-1, 1)) * psiC)                    # :62
    # ---- :66-86 supervector products (Wiener filter of get_wiener.py inside)
    rows = ConOrder * N * N
    A = _lib.DeviceArray(ctx, (rows, E), np.float64)
    m2 = None
    if np.ndim(msk2) == 2:
        m2 = _lib.DeviceArray(ctx, (N, N), np.float64, np.ascontiguousarray(msk2, dtype=np.float64))
    elif np.ndim(msk2) == 0 and float(msk2) != 1.0:
        m2 = _lib.DeviceArray(ctx, (N, N), np.float64, np.full((N, N), float(msk2)))
    _lib.check(lib.mem_nlsa_supervectors_device(ctx.handle, state.H.ptr, state.Ch.ptr, selI.ptr, mu_psi.ctypes.data, num, ConOrder,
                                                E, N, m2.ptr if m2 is not None else None, A.ptr, None))
    lap('supervectors')
    # ---- svdRF.op :19-26 (D1 > D2 branch): eigh of A^T A on the host, U = A V S^-1 on the device
    AtA = np.empty((E, E))
    _lib.check(lib.mem_nlsa_gram_small_device(ctx.handle, A.ptr, rows, E, AtA.ctypes.data, None))
    Dv, V = np.linalg.eigh(AtA)
    order = np.argsort(Dv)[::-1]
    Dv = np.sort(Dv)[::-1]
    V = V[:, order]
    sqrtD = np.sqrt(Dv)
    S = np.diag(sqrtD)
    M = np.ascontiguousarray(np.matmul(V, np.diag(1. / sqrtD)))
    U = _lib.DeviceArray(ctx, (rows, E), np.float64)
    Topo_mean = np.empty((N * N, E))
    _lib.check(lib.mem_nlsa_project_device(ctx.handle, A.ptr, rows, E, M.ctypes.data, U.ptr, N * N, ConOrder, Topo_mean.ctypes.data,
                                           None))
    A.free()
    lap('svd')
    VX = np.matmul(V.T, psiC.T)                                                  # :89
    sdiag = np.diag(S)                                                           # :91 — 1-D: np.diag of svdRF's diagonal matrix
    # ---- :106-144 frames from the first two singular triplets, normalised; squared L2 distances
    Q = np.ascontiguousarray((sdiag[:2, None] * V[:, :2].T) @ psiC.T)            # (2, nI)
    nC = nS - 2 * ConOrder
    IMGT_d = _lib.DeviceArray(ctx, (nC, N * N), np.float64)
    D2 = _lib.DeviceArray(ctx, (nC, nC), np.float64)
    _lib.check(lib.mem_nlsa_reconstruct_device(ctx.handle, U.ptr, N * N, ConOrder, E, Q.ctypes.data, nI, nC, IMGT_d.ptr, D2.ptr, None))
    U.free()
    lap('reconstruct_l2')
    lamb, psirec, sigma, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.embed(D2, nC, tune)     # :146
    D2.free()
    lap('embed_recon')
    a, b, tau = fit_1D_open_manifold_3D.op(psirec)                               # :149
    lap('manifold_fit')
    for d in (selD, selI, m2):
        if d is not None:
            d.free()
    if keep_IMGT_on_device:
        IMGT = IMGT_d
    else:
        IMGT = np.ascontiguousarray(IMGT_d.download().T)
        IMGT_d.free()
    return (IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau)


def frames(IMGT_d, cols):
    """Columns `cols` of the reference's (Npix, nC) IMGT from the device-resident [nC][Npix] array: (Npix, len(cols))."""
    lib = _lib.load()
    Npix = IMGT_d.shape[1]
    out = np.empty((len(cols), Npix))
    for j, c in enumerate(cols):
        _lib.check(lib.mem_copy_d2h(IMGT_d.ctx.handle, out[j].ctypes.data, IMGT_d.ptr + int(c) * Npix * 8, Npix * 8))
    return np.ascontiguousarray(out.T)


def op(NLSAPar, DD, posPath, posPsi1, imgAll, msk2, CTF, ExtPar):
    if 'prD' not in ExtPar:
        raise NotImplementedError("NLSA drop-in: only the 'prD' branch (psiAnalysisParS2) is on the path; 'cuti' is not")
    posPath = np.asarray(posPath)
    state = PdState(np.asarray(DD), imgAll, CTF)
    try:
        out = analyse(state, np.arange(int(NLSAPar['num'])), posPath[posPsi1], NLSAPar, msk2)
    finally:
        state.free()
    if NLSAPar.get('save'):
        from . import myio
        a, b, _ = fit_1D_open_manifold_3D.op(out[2])
        myio.fout1(ExtPar['filename'], ['psirec', 'tau', 'a', 'b'], [out[2], out[7], a, b])     # :155-156
    return out
