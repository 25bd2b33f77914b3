"""Seeded synthetic projection-direction (PD) stacks (SURVEY.md §8d generator).

A PD is a set of nS particle snapshots that look along (nearly) the same
direction.  Each snapshot is a smooth 2-D template deformed along a 1-D latent
coordinate tau (one blob translating and widening), rotated in-plane by the
angle its quaternion encodes, modulated by the CTF of its defocus, plus white
noise; about one third of the members are "conjugates" (index >= nStot/2,
stored un-flipped under index - nStot/2 exactly as the reference's augmented
data set addresses them, getDistanceCTF_local_Conj9combinedS2.py:246-275).

The stack is returned in the SPIDER raw-float32 layout the reference reads
with np.memmap(...).T (:254-258): image i occupies floats [i*N*N, (i+1)*N*N)
and holds the *transpose* of the picture.
"""
import numpy as np

_PIX, _CS, _EKV, _AMPC = 1.255, 2.26, 300.0, 0.1


def euler_to_quat(phi, theta, psi):
    """q = q3(psi) q2(theta) q1(phi) in the Spider convention that
    q2Spider.py:31-34 inverts (unit quaternions as columns, shape (4,n))."""
    phi, theta, psi = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (phi, theta, psi))
    z = np.zeros_like(phi)
    q1 = np.vstack((np.cos(phi / 2), z, z, -np.sin(phi / 2)))
    q2 = np.vstack((np.cos(theta / 2), z, -np.sin(theta / 2), z))
    q3 = np.vstack((np.cos(psi / 2), z, z, -np.sin(psi / 2)))
    return quat_mult(q3, quat_mult(q2, q1))


def quat_mult(q, s):
    """Hamilton product of column quaternions."""
    q0, q1, q2, q3 = q
    s0, s1, s2, s3 = s
    return np.vstack((q0 * s0 - q1 * s1 - q2 * s2 - q3 * s3,
                      q0 * s1 + q1 * s0 + q2 * s3 - q3 * s2,
                      q0 * s2 - q1 * s3 + q2 * s0 + q3 * s1,
                      q0 * s3 + q1 * s2 - q2 * s1 + q3 * s0))


def ctf_2d(N, df, pix=_PIX, Cs=_CS, kev=_EKV, ampc=_AMPC):
    """Unshifted (DC at [0,0]) CTF image for one defocus — used only to make
    the synthetic data CTF-modulated; not the product CTF."""
    f = np.fft.fftfreq(N, d=1.0 / N)
    k2 = (f[:, None] ** 2 + f[None, :] ** 2) / (N / 2.0) ** 2 / (2 * pix) ** 2
    wav = 12.3986 / np.sqrt((1022.0 + kev) * kev)
    g = (0.5 * np.pi * Cs * 1e7 * wav ** 3 * k2 - np.pi * wav * df) * k2
    return np.sin(g) - ampc * np.cos(g)


def make_pd(nS, N, seed=0, snr=0.1, conj_frac=1.0 / 3, n_blobs=20, tilt_sigma=0.03,
            pd_phi=0.7, pd_theta=1.1, with_ctf=True):
    """Returns a dict with
       stack   (n_half*N*N,) float32, SPIDER raw layout
       ind     (nS,) int64 indices into the augmented set [0, nStot)
       q       (4,nS) float64 quaternions of the PD members
       df      (nS,) float64 defocus [A]
       sh      (shx, shy) zero shifts for the un-augmented particles
       nStot   augmented particle count (= 2*n_half)
       tau     (nS,) latent coordinate (ground truth for embedding tests)
       em      dict(nPix, pix_size, Cs, EkV, AmpContrast)
    """
    rng = np.random.default_rng(seed)
    n_half = nS                     # every member has its own stored image
    nStot = 2 * n_half
    conj = rng.random(nS) < conj_frac
    ind = np.arange(nS, dtype=np.int64) + np.where(conj, n_half, 0)
    tau = rng.random(nS)
    psi = rng.uniform(0, 2 * np.pi, nS)
    phi = pd_phi + tilt_sigma * rng.standard_normal(nS)
    theta = pd_theta + tilt_sigma * rng.standard_normal(nS)
    q = euler_to_quat(phi, theta, psi)
    df = rng.uniform(10000.0, 30000.0, nS)
    # in-plane angle the alignment will undo (getDistanceCTF...py:187-202): Psi = 2 atan(s/c) w.r.t. the mean PD
    PDs = 2 * np.vstack((q[1] * q[3] - q[0] * q[2], q[0] * q[1] + q[2] * q[3], q[0] ** 2 + q[3] ** 2 - 0.5))
    PD = PDs.sum(1) / np.linalg.norm(PDs.sum(1))
    sn = -(1 + PD[2]) * q[3] - PD[0] * q[1] - PD[1] * q[2]
    cn = (1 + PD[2]) * q[0] + PD[1] * q[1] - PD[0] * q[2]
    Psi = 2 * np.arctan(sn / cn)

    # template: blobs inside radius 0.35 N; blob 0 moves/widens with tau
    r = 0.35 * N * np.sqrt(rng.random(n_blobs))
    a = rng.uniform(0, 2 * np.pi, n_blobs)
    bx, by = r * np.cos(a), r * np.sin(a)
    bs = rng.uniform(0.02, 0.06, n_blobs) * N
    amp = rng.uniform(0.5, 1.5, n_blobs)
    ctr = (N - 1) / 2.0
    grid = np.arange(N) - ctr
    stack = np.empty((n_half, N, N), dtype=np.float32)
    for i in range(nS):
        cx, cy, s = bx.copy(), by.copy(), bs.copy()
        cx[0] += 0.25 * N * (tau[i] - 0.5)
        s[0] *= 1.0 + 0.5 * tau[i]
        # ndimage.rotate(img, -Psi) must give back the template: blob (row,col) -> R(-Psi) (row,col)
        ca, sa = np.cos(Psi[i]), np.sin(Psi[i])
        ry, rx = ca * cy - sa * cx, sa * cy + ca * cx
        U = np.exp(-(grid[:, None] - ry[None, :]) ** 2 / (2 * s ** 2))      # (N, blobs) rows
        V = np.exp(-(grid[:, None] - rx[None, :]) ** 2 / (2 * s ** 2))      # (N, blobs) cols
        img = (U * amp) @ V.T
        if with_ctf:
            img = np.fft.ifft2(np.fft.fft2(img) * ctf_2d(N, df[i])).real
        sig = img.std()
        img = img + rng.standard_normal((N, N)) * (sig / np.sqrt(snr))
        if conj[i]:
            img = img[::-1, :]       # the reader flips conjugates back (np.flipud, :275)
        stack[i] = img.T             # reader transposes (:258)
    shx = np.zeros(n_half)
    shy = np.zeros(n_half)
    return dict(stack=stack.reshape(-1), ind=ind, q=q, df=df, sh=(shx, shy), nStot=nStot, tau=tau,
                em=dict(nPix=N, pix_size=_PIX, Cs=_CS, EkV=_EKV, AmpContrast=_AMPC))


def make_pd_fast(nS, N, seed=0, snr=0.1, conj_frac=1.0 / 3, n_blobs=20, tilt_sigma=0.03, pd_phi=0.7, pd_theta=1.1,
                 chunk=256, out=None, with_ctf=True):
    """Same model as make_pd, vectorised over chunks of particles (BLAS for the blobs, batched multi-threaded FFT
    for the CTF modulation, float32 noise) so that PDs of BASELINE config sizes (2,000 - 20,000 particles at
    256^2 / 320^2) are generated in seconds.  Different random stream from make_pd; same dict.
    `out`: optional preallocated (nS, N*N) float32 array (e.g. pinned memory) that receives the stack."""
    from scipy import fft as sfft
    rng = np.random.default_rng(seed)
    nStot = 2 * nS
    conj = rng.random(nS) < conj_frac
    ind = np.arange(nS, dtype=np.int64) + np.where(conj, nS, 0)
    tau = rng.random(nS)
    psi = rng.uniform(0, 2 * np.pi, nS)
    q = euler_to_quat(pd_phi + tilt_sigma * rng.standard_normal(nS), pd_theta + tilt_sigma * rng.standard_normal(nS), psi)
    df = rng.uniform(10000.0, 30000.0, nS)
    PDs = 2 * np.vstack((q[1] * q[3] - q[0] * q[2], q[0] * q[1] + q[2] * q[3], q[0] ** 2 + q[3] ** 2 - 0.5))
    PD = PDs.sum(1) / np.linalg.norm(PDs.sum(1))
    sn = -(1 + PD[2]) * q[3] - PD[0] * q[1] - PD[1] * q[2]
    cn = (1 + PD[2]) * q[0] + PD[1] * q[1] - PD[0] * q[2]
    Psi = 2 * np.arctan(sn / cn)
    r = 0.35 * N * np.sqrt(rng.random(n_blobs))
    a = rng.uniform(0, 2 * np.pi, n_blobs)
    bx, by = r * np.cos(a), r * np.sin(a)
    bs = rng.uniform(0.02, 0.06, n_blobs) * N
    amp = rng.uniform(0.5, 1.5, n_blobs).astype(np.float32)
    grid = (np.arange(N) - (N - 1) / 2.0).astype(np.float32)
    f = np.fft.fftfreq(N, d=1.0 / N)
    k2 = ((f[:, None] ** 2 + f[None, :N // 2 + 1] ** 2) / (N / 2.0) ** 2 / (2 * _PIX) ** 2)
    wav = 12.3986 / np.sqrt((1022.0 + _EKV) * _EKV)
    stack = out if out is not None else np.empty((nS, N * N), dtype=np.float32)
    for i0 in range(0, nS, chunk):
        sl = slice(i0, min(nS, i0 + chunk))
        m = sl.stop - sl.start
        cx = np.tile(bx, (m, 1)); cy = np.tile(by, (m, 1)); s = np.tile(bs, (m, 1))
        cx[:, 0] += 0.25 * N * (tau[sl] - 0.5)
        s[:, 0] *= 1.0 + 0.5 * tau[sl]
        ca, sa = np.cos(Psi[sl])[:, None], np.sin(Psi[sl])[:, None]
        ry, rx = (ca * cy - sa * cx).astype(np.float32), (sa * cy + ca * cx).astype(np.float32)
        s32 = s.astype(np.float32)
        U = np.exp(-(grid[None, :, None] - ry[:, None, :]) ** 2 / (2 * s32[:, None, :] ** 2)) * amp     # (m, N, blobs)
        V = np.exp(-(grid[None, :, None] - rx[:, None, :]) ** 2 / (2 * s32[:, None, :] ** 2))
        img = np.matmul(U, V.transpose(0, 2, 1))                                                      # (m, N, N)
        if with_ctf:
            g = (0.5 * np.pi * _CS * 1e7 * wav ** 3 * k2[None] - np.pi * wav * df[sl, None, None]) * k2[None]
            ctf = (np.sin(g) - _AMPC * np.cos(g)).astype(np.float32)
        if with_ctf:
            img = sfft.irfft2(sfft.rfft2(img, workers=-1) * ctf, s=(N, N), workers=-1).astype(np.float32)
        sig = img.reshape(m, -1).std(axis=1)
        img += rng.standard_normal((m, N, N), dtype=np.float32) * (sig / np.sqrt(snr)).astype(np.float32)[:, None, None]
        img[conj[sl]] = img[conj[sl], ::-1, :]
        stack[sl] = img.transpose(0, 2, 1).reshape(m, N * N)
    return dict(stack=stack.reshape(-1), ind=ind, q=q, df=df, sh=(np.zeros(nS), np.zeros(nS)), nStot=nStot, tau=tau,
                PD=PD, em=dict(nPix=N, pix_size=_PIX, Cs=_CS, EkV=_EKV, AmpContrast=_AMPC))
