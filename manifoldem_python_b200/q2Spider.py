"""Quaternion -> Spider Euler angles, and the quaternion product it needs (host-side scalars, row a6).

Same formulation and the same SciPy call as the reference (modules/q2Spider.py:19-60,
modules/qMult_bsx.py:15-35) so that the optimiser lands on the same branch of (phi, theta, psi):
least_squares from a zero start with ftol=1e-12 on  q - q3(psi) q2(theta) q1(phi)."""
import numpy as np
from scipy import optimize


def qmult(q, s):
    q = np.asarray(q, dtype=np.float64).reshape(4, -1)
    s = np.asarray(s, dtype=np.float64).reshape(4, -1)
    q0, qv, s0, sv = q[0], q[1:4], s[0], s[1:4]
    c = np.vstack((qv[1] * sv[2] - qv[2] * sv[1], qv[2] * sv[0] - qv[0] * sv[2], qv[0] * sv[1] - qv[1] * sv[0]))
    return np.vstack((q0 * s0 - np.sum(qv * sv, axis=0), q0 * sv + s0 * qv + c))


def op(q):
    q = np.asarray(q, dtype=np.float64)
    q = q / np.sqrt(np.sum(q ** 2))

    def dev(a):
        q1 = np.array([np.cos(a[0] / 2.), 0., 0., -np.sin(a[0] / 2.)])
        q2 = np.array([np.cos(a[1] / 2.), 0., -np.sin(a[1] / 2.), 0.])
        q3 = np.array([np.cos(a[2] / 2.), 0., 0., -np.sin(a[2] / 2.)])
        return q - qmult(q3, qmult(q2, q1)).flatten()

    res = optimize.least_squares(dev, np.array([0, 0, 0]), bounds=(-np.inf, np.inf), ftol=1e-12)
    return res.x[0], res.x[1], res.x[2]
