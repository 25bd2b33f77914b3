"""Host side of the NLSA stage: fit of the leading three diffusion coordinates to x_ij = a_j cos(j pi tau_i) + b_j
(modules/fit_1D_open_manifold_3D.py:60-146, get_fit_1D_open_manifold_3D_param.py:13-90, solve_d_R_d_tau_p_3D.py:37-52,
R_p.py:32-39).

op(psi) -> (a (3,), b (3,), tau (nS,1))   — same signature and return values as the reference's module.

The reference solves one quintic per data point and iteration with np.roots inside a Python loop (nS x iterations calls).
Here the nS companion matrices of an iteration go through ONE batched np.linalg.eigvals call (the LAPACK routine
np.roots itself ends in, so the roots and their order are the same) and the candidate selection is vectorised.
Kept from the reference: `tau_old = tau` aliases the array updated in place, so the tau criterion never blocks."""
import numpy as np

eps = 1e-4


def _taus(x, a, b):
    """solve_d_R_d_tau_p_3D.op for every row of x (nS,3) at once -> tau (nS,1)."""
    nS = x.shape[0]
    c0 = 48 * a[2] ** 2
    coeff = np.empty((nS, 6))
    coeff[:, 0] = c0
    coeff[:, 1] = 0
    coeff[:, 2] = 8 * a[1] ** 2 - 48 * a[2] ** 2
    coeff[:, 3] = -12 * a[2] * (x[:, 2] - b[2])
    coeff[:, 4] = a[0] ** 2 - 4 * a[1] ** 2 + 9 * a[2] ** 2 - 4 * a[1] * (x[:, 1] - b[1])
    coeff[:, 5] = -a[0] * (x[:, 0] - b[0]) + 3 * a[2] * (x[:, 2] - b[2])
    tau = np.zeros((nS, 1))
    jj = np.array([1, 2, 3])
    # np.roots strips leading / trailing zero coefficients before it builds the companion matrix; those rows (a_3 == 0 or a
    # vanishing constant term) take the reference's own per-point path
    regular = (c0 != 0) & (coeff[:, 5] != 0)
    if regular.any():
        cr = coeff[regular]
        comp = np.zeros((cr.shape[0], 5, 5))
        comp[:, np.arange(1, 5), np.arange(0, 4)] = 1.0
        comp[:, 0, :] = -cr[:, 1:] / cr[:, :1]
        beta = np.linalg.eigvals(comp)                                   # (n, 5), complex
        ok = ~(np.absolute(np.imag(beta)) > 0)
        ok &= ~(np.absolute(beta) > 1)
        br = np.where(ok, np.real(beta), 0.0)
        cand = np.concatenate([np.arccos(br) / np.pi, np.zeros((cr.shape[0], 1)), np.ones((cr.shape[0], 1))], axis=1)   # (n, 7)
        xr = x[regular]
        err = xr[:, None, :] - b - a * np.cos(cand[:, :, None] * jj * np.pi)
        R = np.sum(err ** 2, axis=2)
        R[:, :5][~ok] = np.inf
        pick = np.argmin(R, axis=1)                                      # first minimum in the reference's candidate order
        tau[regular, 0] = cand[np.arange(cr.shape[0]), pick]
    for p in np.nonzero(~regular)[0]:
        beta = np.roots(coeff[p])
        beta = beta[~(np.absolute(np.imag(beta)) > 0)]
        beta = np.real(beta[~(np.absolute(beta) > 1)])
        cand = np.vstack((np.arccos(beta.reshape(-1, 1)) / np.pi, 0, 1))
        err = x[p] - b - a * np.cos(cand * jj * np.pi)
        tau[p] = cand[np.argmin(np.sum(err ** 2, axis=1))]
    return tau


def initial_parameters(psi):
    """get_fit_1D_open_manifold_3D_param.op :21-83 — cubic fit of z(x) and quadratic fit of y(x) -> first (a, b)."""
    nS = psi.shape[0]
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
    rhs = np.array([np.dot(X3.T, Z), np.dot(X2.T, Z), np.dot(X.T, Z), np.sum(Z)])
    D, E, F, G = np.linalg.lstsq(A, rhs)[0]
    disc = E * E - 3 * D * F
    if disc < 0:
        disc = 0.
    if np.absolute(D) < 1e-8:
        D = 1e-8
    a1 = (2. * np.sqrt(disc)) / (3. * D)
    a3 = (2. * disc ** (3 / 2.)) / (27. * D * D)
    b1 = -E / (3 * D)
    b3 = (2. * E * E * E) / (27. * D * D) - (E * F) / (3 * D) + G
    XXb = X * (X - 2 * b1)
    A2 = np.array([[np.sum(XXb * XXb), np.sum(XXb)], [np.sum(XXb), nS]])
    Aq, Cq = np.linalg.lstsq(A2, np.array([np.dot(XXb.T, Y), np.sum(Y)]))[0]
    a2 = 2. * Aq * disc / (9. * D * D)
    b2 = Cq + (Aq * E * E) / (9. * D * D) - (2. * Aq * F) / (3. * D)
    return np.array([a1, a2, a3]), np.array([b1, b2, b3])


def op(psi, maxIter=100, delta_a_max=1, delta_b_max=1):
    """Device path (C ABI mem_manifold_fit_host, nlsa.cu k_manifold_fit): the first (a, b) on the host exactly as the
    reference computes them, the alternating iteration in one kernel launch."""
    import ctypes as C
    from . import _lib
    from .getDistanceCTF_local_Conj9combinedS2 import _ctx
    lib = _lib.load()
    x = np.ascontiguousarray(psi[:, 0:3], dtype=np.float64)
    nS = x.shape[0]
    a, b = initial_parameters(psi)
    ab = np.ascontiguousarray(np.concatenate([a, b]), dtype=np.float64)
    tau = np.empty(nS)
    iters = C.c_int32(0)
    _lib.check(lib.mem_manifold_fit_host(_ctx().handle, x.ctypes.data, nS, ab.ctypes.data, tau.ctypes.data, int(maxIter),
                                         float(delta_a_max), float(delta_b_max), C.byref(iters)))
    return ab[:3].copy(), ab[3:].copy(), tau.reshape(-1, 1)


def op_host(psi, maxIter=100, delta_a_max=1, delta_b_max=1):
    """The same fit in NumPy on the host (one batched eigvals per iteration, the LAPACK routine np.roots ends in): the
    checker of the device kernel in tests/, and a reference for its root isolation."""
    x = psi[:, 0:3]
    nS = x.shape[0]
    a, b = initial_parameters(psi)
    tau = _taus(x, a, b)
    for _ in range(1, maxIter + 1):
        a_old, b_old = a, b
        cos_j_pi_tau = np.cos(np.dot(tau, np.pi * np.array([[1, 2, 3]])))
        A11 = np.sum(cos_j_pi_tau ** 2, axis=0)
        A12 = np.sum(cos_j_pi_tau, axis=0)
        b1 = np.sum(x * cos_j_pi_tau, axis=0)
        b2 = np.sum(x, axis=0)
        coeff = np.zeros((2, 3))
        for qq in range(3):
            coeff[:, qq] = np.linalg.lstsq(np.array([[A11[qq], A12[qq]], [A12[qq], nS]]), np.array([b1[qq], b2[qq]]))[0]
        a, b = coeff[0, :], coeff[1, :]
        tau = _taus(x, a, b)
        delta_a = max(np.fabs(a - a_old) / (np.fabs(a) + eps)) * 100
        delta_b = max(np.fabs(b - b_old) / (np.fabs(b) + eps)) * 100
        if delta_a < delta_a_max and delta_b < delta_b_max:       # the reference's delta_tau is identically 0 (aliasing)
            break
    return a, b, tau
