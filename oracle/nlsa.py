"""Oracle: NLSA / psi-analysis stage.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in plain NumPy at the reference's call sites,

    modules/NLSA.py:23-158                         (op)
    modules/get_wiener.py:10-22                    (op)
    modules/svdRF.py:9-32                          (tidyUp, op)
    modules/L2_distance.py:8-41                    (op)
    modules/fit_1D_open_manifold_3D.py:60-146      (op)  with get_fit_1D_open_manifold_3D_param.py:13-90,
                                                   solve_d_R_d_tau_p_3D.py:37-52 and R_p.py:32-39
    modules/psiAnalysisParS2.py:84-124             (tau rescaling and the IMG1 class representatives)

Third-party arithmetic at the same call sites as the reference: scipy.fftpack fft2 / ifft2, np.linalg.eigh, np.linalg.lstsq,
np.roots; DMembeddingII.op is oracle.dm_embedding.dm_embedding.  Parity: PINNED by tests/golden/nlsa_nS80_N24.npz (outputs of
the unmodified reference, tests/golden/make_golden_nlsa.py) in tests/test_oracle_golden.py.
"""
import numpy as np
from scipy.fftpack import fft2, ifft2

from . import dm_embedding as odm


def con_d(DD, num, ConOrder):
    """NLSA.py:30-33 — sum of ConOrder diagonal-shifted (num - ConOrder)^2 blocks of DD."""
    n = num - ConOrder
    ConD = np.zeros((n, n))
    for i in range(ConOrder):
        ConD += DD[i:i + n][:, i:i + n]
    return ConD


def wiener(CTF, posPath, posPsi1, ConOrder, num):
    """get_wiener.op :10-22 — wiener_dom[i] = sum_{ii < ConOrder} CTF1[ConOrder - ii + i]^2 + 1/SNR, SNR = 5."""
    CTF1 = CTF[posPath[posPsi1], :, :]
    dim = CTF.shape[1]
    wd = np.zeros((num - ConOrder, dim, dim))
    for i in range(num - ConOrder):
        for ii in range(ConOrder):
            wd[i] = wd[i] + CTF1[ConOrder - ii + i] ** 2
    return wd + 1. / 5, CTF1


def svd_rf(A):
    """svdRF.op :19-32 for D1 > D2 (the only branch NLSA reaches): eigh of A^T A, descending, U = A V S^-1."""
    D, V = np.linalg.eigh(np.matmul(A.T, A))
    order = np.argsort(D)[::-1]
    D = np.sort(D)[::-1]
    V = V[:, order]
    sq = np.sqrt(D)
    return np.matmul(A, np.matmul(V, np.diag(1. / sq))), np.diag(sq), V


def l2_distance(a, b):
    """L2_distance.op :33-41."""
    aa = np.sum(a ** 2, axis=0)
    bb = np.sum(b ** 2, axis=0)
    tmp = aa[:, None] + bb[None, :] - 2 * np.matmul(a.T, b)
    tmp[np.nonzero(tmp < 1e-8)] = 0
    return np.sqrt(tmp)


# ------------------------------------------------------------------------------------------------ 1-D manifold fit
def _tau_of_point(x_p, a, b):
    """solve_d_R_d_tau_p_3D.op :37-52 with R_p.op :32-39 for one data point x_p (3,)."""
    coeff = np.array([48 * a[2] ** 2, 0, 8 * a[1] ** 2 - 48 * a[2] ** 2, -12 * a[2] * (x_p[2] - b[2]),
                      a[0] ** 2 - 4 * a[1] ** 2 + 9 * a[2] ** 2 - 4 * a[1] * (x_p[1] - b[1]),
                      -a[0] * (x_p[0] - b[0]) + 3 * a[2] * (x_p[2] - b[2])])
    beta = np.roots(coeff)
    beta = beta[~(np.absolute(np.imag(beta)) > 0)]
    beta = np.real(beta[~(np.absolute(beta) > 1)])
    cand = np.vstack((np.arccos(beta.reshape(-1, 1)) / np.pi, 0, 1))
    err = x_p - b - a * np.cos(cand * np.array([1, 2, 3]) * np.pi)
    return cand[np.argmin(np.sum(err ** 2, axis=1))]


def fit_1d_open_manifold_3d(psi, max_iter=100, da_max=1.0, db_max=1.0, dtau_max=0.01):
    """fit_1D_open_manifold_3D.op :60-146 — x_ij = a_j cos(j pi tau_i) + b_j, j = 1..3, alternating fits.
    NB (kept from the reference): `tau_old = tau` aliases the array that is then updated in place, so delta_tau is
    always 0 and only the a / b criteria stop the iteration."""
    x = psi[:, 0:3]
    nS = x.shape[0]
    # ---- get_fit_1D_open_manifold_3D_param.op :13-90: cubic / quadratic moment fits for the first a, b
    X, Y, Z = psi[:, 0], psi[:, 1], psi[:, 2]
    X2 = X * X
    X3 = X2 * X
    X4 = X2 * X2
    X5 = X3 * X2
    X6 = X3 * X3
    A = np.array([[np.sum(X6), np.sum(X5), np.sum(X4), np.sum(X3)],
                  [np.sum(X5), np.sum(X4), np.sum(X3), np.sum(X2)],
                  [np.sum(X4), np.sum(X3), np.sum(X2), np.sum(X)],
                  [np.sum(X3), np.sum(X2), np.sum(X), nS]])
    bvec = np.array([np.dot(X3.T, Z), np.dot(X2.T, Z), np.dot(X.T, Z), np.sum(Z)])
    D_, E_, F_, G_ = np.linalg.lstsq(A, bvec)[0]
    disc = E_ * E_ - 3 * D_ * F_
    if disc < 0:
        disc = 0.
    if np.absolute(D_) < 1e-8:
        D_ = 1e-8
    a1 = (2. * np.sqrt(disc)) / (3. * D_)
    a3 = (2. * disc ** (3 / 2.)) / (27. * D_ * D_)
    b1 = -E_ / (3 * D_)
    b3 = (2. * E_ * E_ * E_) / (27. * D_ * D_) - (E_ * F_) / (3 * D_) + G_
    Xb = X - 2 * b1
    XXb = X * Xb
    A2 = np.array([[np.sum(XXb * XXb), np.sum(XXb)], [np.sum(XXb), nS]])
    Ac, Cc = np.linalg.lstsq(A2, np.array([np.dot(XXb.T, Y), np.sum(Y)]))[0]
    a2 = 2. * Ac * disc / (9. * D_ * D_)
    b2 = Cc + (Ac * E_ * E_) / (9. * D_ * D_) - (2. * Ac * F_) / (3. * D_)
    a = np.array([a1, a2, a3])
    b = np.array([b1, b2, b3])
    tau = np.zeros((nS, 1))
    for p in range(nS):
        tau[p] = _tau_of_point(x[p], a, b)
    # ---- the alternating iteration :66-144
    for _ in range(1, max_iter + 1):
        a_old, b_old = a, b
        cosj = np.cos(np.dot(tau, np.pi * np.array([[1, 2, 3]])))
        A11 = np.sum(cosj ** 2, axis=0)
        A12 = np.sum(cosj, axis=0)
        b1v = np.sum(x * cosj, axis=0)
        b2v = np.sum(x, axis=0)
        coeff = np.zeros((2, 3))
        for qq in range(3):
            coeff[:, qq] = np.linalg.lstsq(np.array([[A11[qq], A12[qq]], [A12[qq], nS]]), np.array([b1v[qq], b2v[qq]]))[0]
        a, b = coeff[0, :], coeff[1, :]
        for p in range(nS):
            tau[p] = _tau_of_point(x[p], a, b)
        delta_a = max(np.fabs(a - a_old) / (np.fabs(a) + 1e-4)) * 100
        delta_b = max(np.fabs(b - b_old) / (np.fabs(b) + 1e-4)) * 100
        if delta_a < da_max and delta_b < db_max:          # delta_tau == 0 always (aliasing, see above)
            break
    return a, b, tau


# ------------------------------------------------------------------------------------------------ NLSA.op
def nlsa(NLSAPar, DD, posPath, posPsi1, imgAll, msk2, CTF, num_eigs=15, rng=None):
    """NLSA.op :23-158 for the 'prD' branch (the one psiAnalysisParS2 takes).  DD is mutated by DMembeddingII like the
    reference (diag <- -inf is applied to ConD, a fresh array, so DD itself stays).  Returns
    (IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau)."""
    num, ConOrder, k, tune = NLSAPar['num'], NLSAPar['ConOrder'], NLSAPar['k'], NLSAPar['tune']
    nS, psiTrunc = NLSAPar['nS'], NLSAPar['psiTrunc']
    ConD = con_d(DD, num, ConOrder)
    lambdaC, psiC, sigmaC, mu, _, _, _, _ = odm.dm_embedding(ConD, k, tune, 600000, num_eigs=num_eigs, rng=rng)
    psiC1 = np.copy(psiC)
    IMG1 = imgAll[posPath[posPsi1], :, :]
    wd, CTF1 = wiener(CTF, posPath, posPsi1, ConOrder, num)
    dim = CTF.shape[1]
    ell = psiTrunc - 1
    N = psiC.shape[0]
    psiC = np.hstack((np.ones((N, 1)), psiC[:, 0:ell]))
    mu_psi = mu.reshape((-1, 1)) * psiC
    A = np.zeros((ConOrder * dim * dim, ell + 1))
    tmp = np.zeros((dim * dim, num - ConOrder))
    for ii in range(ConOrder):
        for i in range(num - ConOrder):
            ind3 = ConOrder - ii + i - 1
            img_f = fft2(IMG1[ind3])
            img = ifft2(img_f * (CTF1[ind3] / wd[i])).real * msk2
            tmp[:, i] = np.squeeze(img.T.reshape(-1, 1))
        A[ii * dim * dim:(ii + 1) * dim * dim, :] = np.matmul(tmp, mu_psi)
    U, S, V = svd_rf(A)
    VX = np.matmul(V.T, psiC.T)
    sdiag = np.diag(S)
    Npixel = dim * dim
    Topo_mean = np.zeros((Npixel, psiTrunc))
    for ii in range(psiTrunc):
        Topo = np.stack([U[kk * Npixel:(kk + 1) * Npixel, ii] for kk in range(ConOrder)], axis=1)
        Topo_mean[:, ii] = np.mean(Topo, axis=1)
    ConImgT = np.zeros((max(U.shape), ell + 1))
    for i in range(0, 2):
        ConImgT = ConImgT + np.matmul(U[:, i].reshape(-1, 1), sdiag[i] * (V[:, i].reshape(1, -1)))
    IMGT = np.zeros((Npixel, nS - 2 * ConOrder))
    for i in range(ConOrder):
        t = np.matmul(ConImgT[i * Npixel:(i + 1) * Npixel, :], psiC.T)
        for ii in range(num - 2 * ConOrder):
            IMGT[:, ii] = IMGT[:, ii] + t[:, i + ii]
    for i in range(IMGT.shape[1]):
        IMGT[:, i] = (IMGT[:, i] - np.mean(IMGT[:, i])) / np.std(IMGT[:, i])
    Drecon = l2_distance(IMGT, IMGT)
    lamb, psirec, sigma, mu, _, _, _, _ = odm.dm_embedding(Drecon ** 2, min(IMGT.shape), tune, 30, num_eigs=num_eigs, rng=rng)
    a, b, tau = fit_1d_open_manifold_3d(psirec)
    return IMGT, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau


def class_representatives(IMGT, tau, nClass):
    """psiAnalysisParS2.py:96-124 — tau rescaled to [0, 1]; for each of numclass bins the first snapshot inside
    (bins widened by 1/(2 numclass) of their bounds while empty).  Returns (IMG1, tau, tauinds)."""
    nSrecon = min(IMGT.shape)
    numclass = int(min(nClass, np.floor(nSrecon / 2.)))
    tau = (tau - min(tau)) / (max(tau) - min(tau))
    tauinds = []
    IMG1 = np.zeros((IMGT.shape[0], numclass))
    for i in range(numclass):
        ind1 = float(i) / numclass
        ind2 = ind1 + 1. / numclass
        if i == numclass - 1:
            tauind = ((tau >= ind1) & (tau <= ind2)).nonzero()[0]
        else:
            tauind = ((tau >= ind1) & (tau < ind2)).nonzero()[0]
        while tauind.size == 0:
            sc = 1. / (numclass * 2.)
            ind1 = ind1 - sc * ind1
            ind2 = ind2 + sc * ind2
            tauind = ((tau >= ind1) & (tau < ind2)).nonzero()[0]
        IMG1[:, i] = IMGT[:, tauind[0]]
        tauinds.append(tauind[0])
    return IMG1, tau, tauinds
