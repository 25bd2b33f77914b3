"""Drop-in for the embedding-stage worker that consumes the per-PD pickle
(modules/manifoldTrimmingAuto.py:38-95): load D, embed with DMembeddingII.op(D, k = nS, tune), iteratively drop
the points whose leading three diffusion coordinates lie outside radius `rad` and re-embed, then write the
psi pickle, the resume marker (after the dump) and the eigenvalue spectrum text file.

Same signature and outputs as the reference; the heavy part of every DMembeddingII.op call (kNN, graph,
Ferguson sweep, Gaussian-kernel Laplacian) runs on the GPU.
"""
import os

import numpy as np

from . import DMembeddingII, myio
from .getDistanceCTF_local_Conj9combinedS2 import _cfg


def get_psiPath(psi, rad, plotNum):
    """(:18-28) indices of the points inside the sphere of radius `rad` in (psi_n, psi_n+1, psi_n+2)."""
    d = np.sqrt(psi[:, plotNum] ** 2 + psi[:, plotNum + 1] ** 2 + psi[:, plotNum + 2] ** 2)
    return (d < rad).nonzero()[0]


def op(input_data, posPath, tune, rad, visual, doSave):
    p = _cfg()
    dist_file, psi_file, eig_file, prD = input_data[0], input_data[1], input_data[2], input_data[3]
    data = myio.fin1(dist_file)
    D = data['D']
    ind = data['ind']
    nS = D.shape[1]
    if isinstance(posPath, int) and posPath == 0:                 # :48-49
        posPath = np.arange(nS)
    D = D[posPath][:, posPath]
    nS = D.shape[1]
    k = nS
    lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.op(D, k, tune, 60000)   # :54
    posPath1 = get_psiPath(psi, rad, 0)
    while len(posPath1) < nS:                                      # :61-70
        nS = len(posPath1)
        D1 = D[posPath1][:, posPath1]
        k = D1.shape[0]
        lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.op(D1, k, tune, 600000)
        lamb = lamb[lamb > 0]
        posPathInt = get_psiPath(psi, rad, 0)
        posPath1 = posPath1[posPathInt]
    posPath = posPath[posPath1]
    if doSave['Is']:                                               # :79-87
        myio.fout1(psi_file, ['lamb', 'psi', 'sigma', 'mu', 'posPath', 'ind', 'logEps', 'logSumWij', 'popt', 'R_squared'],
                   [lamb, psi, sigma, mu, posPath, ind, logEps, logSumWij, popt, R_squared])
        open(os.path.join(p.psi_prog, '%s' % (prD)), 'a').close()  # marker after the dump
    if os.path.exists(eig_file):                                   # :89-92
        os.remove(eig_file)
    with open(eig_file, 'a') as f:
        for i in range(len(lamb) - 1):
            f.write("%d\t%.5f\n" % (i + 1, lamb[i + 1]))
    return None
