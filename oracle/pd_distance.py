"""Oracle: per-PD CTF-corrected pairwise distances (float64 NumPy/SciPy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, function by function,
what the reference computes in

    modules/getDistanceCTF_local_Conj9combinedS2.py:216-420   (op)
    modules/annularMask.py:19-32, modules/rotatefill.py:21-41,
    modules/ctemh_cryoFrank.py:24-44, modules/q2Spider.py:19-60,
    modules/qMult_bsx.py:15-35

Third-party arithmetic the reference delegates to (not under /root/reference):
scipy.fftpack.fft2/ifft2, scipy.ndimage.rotate/shift (cubic spline),
scipy.optimize.least_squares, numpy.dot — the same SciPy 1.18.1 / NumPy 2.3.5
that this image ships, called here at the same call sites.

All arrays are float64 / complex128 exactly as in the reference.
"""
import math

import numpy as np
from scipy import ndimage, optimize
from scipy.fftpack import fft2, ifft2, ifftshift

VERSION = 'getDistanceCTF_local9, V 1.0'   # getDistanceCTF...py:41


# --------------------------------------------------------------------- a4
def annular_mask(a, b, N, M):
    """annularMask.py:19-32 — centre is (N/2-1, M/2); a^2 <= r^2 < b^2."""
    x = np.arange(N, dtype=np.float64) - N / 2 + 1
    y = np.arange(M, dtype=np.float64) - M / 2
    r2 = (x * x)[:, None] + (y * y)[None, :]
    return ((r2 >= a * a) & (r2 < b * b)).astype(np.float64)


# --------------------------------------------------------------------- a5
def create_grid(N):
    """getDistanceCTF...py:139-154 — Q = |(X,Y)| / (N/2)."""
    if N <= 0:
        raise AssertionError('non-positive image size')
    if N % 2 == 1:
        a = np.arange(-(N - 1) / 2, (N - 1) / 2 + 1)
    else:
        a = np.arange(-N / 2., N / 2)
    X, Y = np.meshgrid(a, a)
    return (1. / (N / 2.)) * np.sqrt(X ** 2 + Y ** 2)


def create_filter(filter_type, NN, Qc, Q):
    """getDistanceCTF...py:156-167."""
    if filter_type == 'Gauss':
        return np.exp(-(np.log(2) / 2.) * (Q / Qc) ** 2)
    if filter_type == 'Butter':
        return np.sqrt(1. / (1 + (Q / Qc) ** (2 * NN)))
    raise ValueError('%s filter is unsupported' % filter_type)


# --------------------------------------------------------------------- a8
def ctemh_cryo_frank(k, Cs_mm, df, kev, B, ampc):
    """ctemh_cryoFrank.py:24-44."""
    Cs = Cs_mm * 1.0e7
    wav = 12.3986 / np.sqrt((2 * 511.0 + kev) * kev)
    w1 = np.pi * Cs * wav * wav * wav
    w2 = np.pi * wav * df
    k2 = k * k
    sigm = B / math.sqrt(2 * math.log(2))
    wi = np.exp(-k2 / (2 * sigm ** 2))
    wr = (0.5 * w1 * k2 - w2) * k2
    return (np.sin(wr) - ampc * np.cos(wr)) * wi


# --------------------------------------------------------------------- a7
def rotatefill(img, angle_deg, impl='tile'):
    """rotatefill.py:21-41.

    impl='tile'     : literally what the reference does — 3x3 tile,
                      ndimage.rotate(order=3, reshape=False), centre crop.
    impl='periodic' : the equivalent closed form (SURVEY §7 hard part 2):
                      periodic cubic-B-spline resampling of the NxN image at
                      R(o-c)+c, c=(N-1)/2.  Agrees with 'tile' to ~1e-13
                      (checked in tests/test_oracle_golden.py); this is the
                      statement the CUDA kernel implements.
    """
    n = img.shape[0]
    if impl == 'tile':
        rep = np.tile(img, (3, 3))
        out = ndimage.rotate(rep, angle_deg, reshape=False)
        return out[n:2 * n, n:2 * n]
    th = np.deg2rad(angle_deg)
    c, s = math.cos(th), math.sin(th)
    ctr = (n - 1) / 2.0
    o0, o1 = np.meshgrid(np.arange(n) - ctr, np.arange(n) - ctr, indexing='ij')
    # ndimage.rotate builds matrix [[c, s], [-s, c]] mapping output->input coords
    i0 = c * o0 + s * o1 + ctr
    i1 = -s * o0 + c * o1 + ctr
    return ndimage.map_coordinates(img, [i0, i1], order=3, mode='grid-wrap')


# --------------------------------------------------------------------- a6
def q_mult(q, s):
    """qMult_bsx.py:15-35 — Hamilton product, columns are quaternions."""
    q = np.asarray(q, dtype=np.float64).reshape(4, -1)
    s = np.asarray(s, dtype=np.float64).reshape(4, -1)
    q0, qv = q[0], q[1:4]
    s0, sv = s[0], s[1:4]
    cross = np.vstack((qv[1] * sv[2] - qv[2] * sv[1],
                       qv[2] * sv[0] - qv[0] * sv[2],
                       qv[0] * sv[1] - qv[1] * sv[0]))
    return np.vstack((q0 * s0 - np.sum(qv * sv, axis=0), q0 * sv + s0 * qv + cross))


def q2spider(q):
    """q2Spider.py:19-60 — (phi,theta,psi) with q = q3(psi) q2(theta) q1(phi),
    found by least_squares from a zero start, ftol=1e-12."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.sqrt(np.sum(q ** 2))

    def dev(a):
        q1 = np.array([np.cos(a[0] / 2.), 0., 0., -np.sin(a[0] / 2.)])
        q2 = np.array([np.cos(a[1] / 2.), 0., -np.sin(a[1] / 2.), 0.])
        q3 = np.array([np.cos(a[2] / 2.), 0., 0., -np.sin(a[2] / 2.)])
        return q - q_mult(q3, q_mult(q2, q1)).flatten()

    res = optimize.least_squares(dev, np.array([0, 0, 0]), bounds=(-np.inf, np.inf), ftol=1e-12)
    return res.x[0], res.x[1], res.x[2]


def calc_avg_pd(q):
    """getDistanceCTF...py:169-185 — per-image projection directions (3,nS)."""
    if q.shape[0] <= 3:
        raise AssertionError('quaternion has wrong dimensions')
    return 2 * np.vstack((q[1] * q[3] - q[0] * q[2],
                          q[0] * q[1] + q[2] * q[3],
                          q[0] ** 2 + q[3] ** 2 - 0.5))


def get_psi(q, PD):
    """getDistanceCTF...py:187-202 vectorised over images; NaN -> 0 (:319-320)."""
    s = -(1 + PD[2]) * q[3] - PD[0] * q[1] - PD[1] * q[2]
    c = (1 + PD[2]) * q[0] + PD[1] * q[1] - PD[0] * q[2]
    with np.errstate(divide='ignore', invalid='ignore'):
        psi = 2 * np.arctan(s / c)
    psi = np.where(np.isnan(psi), 0.0, psi)
    return psi, s, c


def psi_ang(PD):
    """getDistanceCTF...py:206-213 — in-plane angle of the PD itself, degrees."""
    Qr = np.array([1 + PD[2], PD[1], -PD[0], 0.0])
    Qr = Qr / np.sqrt(np.sum(Qr ** 2))
    _, _, psi = q2spider(Qr)
    return np.mod(psi, 2 * np.pi) * (180 / np.pi)


# --------------------------------------------------------------------- a2/a3
def read_particle(stack, idx, N, relion, sh):
    """getDistanceCTF...py:253-264.

    stack: for SPIDER a flat float32 array/memmap (image i at offset N*N*i,
    read (N,N) then transposed); for RELION an (n,N,N) float32 array (what
    mrcfile.mmap(...).data is), followed by the cubic 'wrap' shift.
    """
    if not relion:
        tmp = np.asarray(stack[N * N * idx:N * N * (idx + 1)]).reshape(N, N).T
    else:
        tmp = stack[idx]
        shi = (sh[1][idx] - 0.5, sh[0][idx] - 0.5)
        tmp = ndimage.shift(tmp, shi, order=3, mode='wrap')
    return tmp


def ingest(stack, ind, nStot, N, msk, relion=False, sh=None):
    """getDistanceCTF...py:246-283 — conjugate handling + background normalise.
    Returns (imgs (nS,N,N) f64, imgLabels (nS,) int)."""
    nS = len(ind)
    out = np.zeros((nS, N, N))
    labels = np.zeros(nS, dtype=int)
    for iS in range(nS):
        conj = not (ind[iS] < nStot / 2)
        idx = int(ind[iS] - nStot / 2) if conj else int(ind[iS])
        labels[iS] = -1 if conj else 1
        tmp = read_particle(stack, idx, N, relion, sh)
        if conj:
            tmp = np.flipud(tmp)
        backg = tmp * (1 - msk)
        tmp = (tmp - backg.mean()) / backg.std()
        out[iS] = tmp
    return out, labels


def lowpass(imgs, G):
    """getDistanceCTF...py:286-293 (G already ifftshift-ed)."""
    return ifft2(fft2(imgs, axes=(-2, -1)) * G, axes=(-2, -1)).real


# --------------------------------------------------------------------- op
def pd_distance(ind, q, df, stack, nStot, N, pix_size, Cs, EkV, AmpContrast,
                gaussEnv=np.inf, filterPar=None, msk2=1, relion=False, sh=None,
                avg_only=False, direct=False, rotate_impl='tile', keep=None, pd_override=None,
                timings=None):
    """getDistanceCTF...py:216-420 without the file I/O: returns the dict that
    the reference pickles (same keys, shapes, dtypes), plus a few named
    intermediates under '_'-prefixed keys for per-stage parity checks.

    direct=True evaluates D by the definitional per-pair norm
    (conquer, :106-122) instead of the matrix identity (:391-397).

    pd_override=PD (3,): use this mean projection direction instead of the mean
    over `q` (:299-301).  D_ij depends on particles i and j and, through Psi and
    psi_p, on PD only — so a SUBSET of a large PD evaluated with the full PD's
    direction reproduces exactly the entries D[sub][:, sub] of the full matrix
    (sampled-pair parity checks at BASELINE config sizes).

    timings: optional dict that receives the wall seconds of the stages (ingest, lowpass, rotate, ctf, fft, averages,
    gemm) — the per-stage CPU split bench.py reports for the reference algorithm.
    """
    import time as _time
    _t = [_time.perf_counter()]

    def _lap(name):
        now = _time.perf_counter()
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + now - _t[0]
        _t[0] = now
    if filterPar is None:
        filterPar = dict(type='Butter', Qc=0.5, N=8)            # GetDistancesS2.py:83
    ind = np.asarray(ind)
    nS = ind.shape[0]
    msk = annular_mask(0, N / 2., N, N)                          # :242
    y, imgLabels = ingest(stack, ind, nStot, N, msk, relion, sh)  # :246-283
    _lap('ingest')
    Q = create_grid(N)                                           # :286
    G = ifftshift(create_filter(filterPar['type'], filterPar['N'], filterPar['Qc'], Q))
    y = lowpass(y, G)                                            # :290-293
    _lap('lowpass')
    PDs = calc_avg_pd(q)                                         # :297
    PD = np.sum(PDs, 1)
    PD = PD / np.linalg.norm(PD)                                 # :299-301
    if pd_override is not None:
        PD = np.asarray(pd_override, dtype=np.float64)
    psi_p = psi_ang(PD)                                          # :313
    Psi, s, c = get_psi(q, PD)                                   # :317-323

    CTF = np.zeros((nS, N, N))
    fy = np.zeros((nS, N, N), dtype=np.complex128)
    imgAll = np.zeros((nS, N, N))
    imgAllFlip = np.zeros((nS, N, N))
    for iS in range(nS):                                         # :315-349
        img = y[iS] * msk
        img = rotatefill(img, -(180 / math.pi) * Psi[iS], rotate_impl)
        img = rotatefill(img, -psi_p, rotate_impl)
        _lap('rotate')
        CTF[iS] = ifftshift(ctemh_cryo_frank(Q / (2 * pix_size), Cs, df[iS], EkV, gaussEnv, AmpContrast))
        _lap('ctf')
        fy[iS] = fft2(img * msk2)
        imgAllFlip[iS] = ifft2(np.sign(CTF[iS]) * fy[iS]).real
        imgAll[iS] = img
        _lap('fft')
    imgAvgFlip = imgAllFlip.sum(0)

    wiener_dom = -(np.sum(CTF ** 2, axis=0) + 1. / 5)            # :354, :422-430 (SNR=5)
    imgAvg = ifft2(fft2(imgAll, axes=(-2, -1)) * (CTF / wiener_dom), axes=(-2, -1)).real.sum(0)
    imgAvg = imgAvg * msk2 / nS                                  # :366-367
    imgAvgFlip = imgAvgFlip * msk2 / nS
    _lap('averages')

    D = np.zeros((nS, nS))
    CTF_out = CTF
    if not avg_only:
        if direct:                                               # :106-122
            for x in range(nS):
                for yy in range(x + 1, nS):
                    D[x, yy] = np.linalg.norm(CTF[x] * fy[yy] - CTF[yy] * fy[x]) ** 2
            D = D + D.T
        else:                                                    # :391-397
            fyf = fy.reshape(nS, N * N)
            CTF_out = CTF.reshape(nS, N * N)
            CTFfy = CTF_out.conj() * fyf
            D = np.dot(np.abs(CTF_out) ** 2, (np.abs(fyf) ** 2).T)
            D = D + D.T - 2 * np.real(np.dot(CTFfy, CTFfy.conj().T))
    _lap('gemm')
    imgAllIntensity = np.mean(imgAllFlip ** 2, axis=0)           # :400

    out = dict(D=D, ind=ind, q=q, df=df, CTF=CTF_out, imgAll=imgAll, msk2=msk2, PD=PD, PDs=PDs,
               Psis=Psi.reshape(nS, 1), imgAvg=imgAvg, imgAvgFlip=imgAvgFlip, imgAllFlip=imgAllFlip,
               imgLabels=imgLabels, Dnom=c.reshape(nS, 1), Nom=s.reshape(nS, 1),
               imgAllIntensity=imgAllIntensity, version=VERSION)
    out['_psi_p'] = psi_p
    out['_filtered'] = y
    out['_fy'] = fy
    if keep is not None:
        out = {k: v for k, v in out.items() if k in keep}
    return out
