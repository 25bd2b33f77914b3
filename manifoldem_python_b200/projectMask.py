"""Optional volumetric mask projected along the PD (row a9; host side, once per PD, only when
p.mask_vol_file is set).  Same construction as the reference (modules/projectMask.py:47-135): Spider ZYZ
Euler matrix from the PD's (phi, theta), scipy.ndimage.affine_transform (cubic, mode='nearest') about the
volume centre, sum along z, threshold > 1."""
from math import cos, sin

import numpy as np
from scipy.ndimage import affine_transform

from . import q2Spider


def euler_matrix_spider(Phi, Theta, Psi):
    return np.array([
        [cos(Phi) * cos(Psi) * cos(Theta) - sin(Phi) * sin(Psi), cos(Psi) * cos(Theta) * sin(Phi) + cos(Phi) * sin(Psi), -cos(Psi) * sin(Theta)],
        [-cos(Psi) * sin(Phi) - cos(Phi) * cos(Theta) * sin(Psi), cos(Phi) * cos(Psi) - cos(Theta) * sin(Phi) * sin(Psi), sin(Psi) * sin(Theta)],
        [cos(Phi) * sin(Theta), sin(Phi) * sin(Theta), cos(Theta)]])


def op(vol, PD):
    vol = np.swapaxes(vol, 0, 2)
    n = vol.shape[0]
    Qr = np.array([1 + PD[2], PD[1], -PD[0], 0.0])
    phi, theta, _ = q2Spider.op(Qr / np.sqrt(np.sum(Qr ** 2)))
    sym = -np.array([phi, theta, 0.0])            # psi = 0: images are already rotated in-plane; inverse transform
    R = euler_matrix_spider(sym[2], sym[1], sym[0])
    c = 0.5 * np.array(vol.shape)
    rho = affine_transform(input=vol, matrix=R, offset=c - R @ c, output_shape=vol.shape, mode='nearest')
    return np.sum(rho, axis=2).reshape(n, n).T > 1
