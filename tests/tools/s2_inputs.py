"""Seeded orientation sets for the S2-tessellation tests and golden vectors (shared by tests/golden/make_golden_s2.py
and the tests, so the 1 MB of quaternions need not be committed)."""
import numpy as np

CASES = (('a', 30000, 4 * 5.0 / 360, 25, 400, 3), ('b', 5000, 0.2, 10, 30, 4))   # tag, n, shAngWidth, PDsizeThL, thres, seed


def quats(n, seed):
    """n orientations: 70 % uniform on the sphere, 30 % clustered around three views (gives occupied PDs)."""
    from manifoldem_python_b200 import synthetic
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0, 2 * np.pi, n)
    theta = np.arccos(rng.uniform(-1, 1, n))
    m = int(0.3 * n)
    centres = np.array([[0.3, 0.5], [2.0, 1.2], [4.0, 2.2]])
    which = rng.integers(0, 3, m)
    phi[:m] = centres[which, 0] + 0.05 * rng.standard_normal(m)
    theta[:m] = centres[which, 1] + 0.05 * rng.standard_normal(m)
    return synthetic.euler_to_quat(phi, theta, rng.uniform(0, 2 * np.pi, n))
