"""Oracle: optional volumetric mask projected along the PD (row a9).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
modules/projectMask.py:47-135 (eulerRotMatrix3DSpider, rotateVolumeEuler,
getEuler_from_PD, op); third-party: scipy.ndimage.affine_transform (order 3,
mode='nearest') at the same call site.
"""
from math import cos, sin

import numpy as np
from scipy.ndimage import affine_transform

from .pd_distance import q2spider


def euler_matrix_spider(Phi, Theta, Psi):
    """projectMask.py:47-57 (radians)."""
    return np.array([
        [cos(Phi) * cos(Psi) * cos(Theta) - sin(Phi) * sin(Psi), cos(Psi) * cos(Theta) * sin(Phi) + cos(Phi) * sin(Psi), -cos(Psi) * sin(Theta)],
        [-cos(Psi) * sin(Phi) - cos(Phi) * cos(Theta) * sin(Psi), cos(Phi) * cos(Psi) - cos(Theta) * sin(Phi) * sin(Psi), sin(Psi) * sin(Theta)],
        [cos(Phi) * sin(Theta), sin(Phi) * sin(Theta), cos(Theta)]])


def project_mask(vol, PD):
    """projectMask.op :118-135 -> boolean (N,N) msk2."""
    vol = np.swapaxes(vol, 0, 2)
    n = vol.shape[0]
    Qr = np.array([1 + PD[2], PD[1], -PD[0], 0.0])
    phi, theta, _ = q2spider(Qr / np.sqrt(np.sum(Qr ** 2)))       # :95-101
    sym = -np.array([phi, theta, 0.0])                            # :126-127
    R = euler_matrix_spider(sym[2], sym[1], sym[0])               # :62 (argument order as in the reference)
    c = 0.5 * np.array(vol.shape)
    rho = affine_transform(input=vol, matrix=R, offset=c - R @ c, output_shape=vol.shape, mode='nearest')
    return np.sum(rho, axis=2).reshape(n, n).T > 1
