"""Drop-in for the embedding-stage worker that consumes the per-PD pickle
(modules/manifoldTrimmingAuto.py:38-95): load D, embed with DMembeddingII.op(D, k = nS, tune), iteratively drop
the points whose leading three diffusion coordinates lie outside radius `rad` and re-embed, then write the
psi pickle, the resume marker (after the dump) and the eigenvalue spectrum text file.

Same signature and outputs as the reference; the heavy part of every embedding (kNN, graph, Ferguson sweep,
Gaussian-kernel Laplacian) runs on the GPU, and so does the sub-matrix selection of the trimming loop: D is uploaded
once per PD and stays on the device.
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
    ind = data['ind']
    # D goes to the device once, in the dtype it is stored in (a sidecar record holds the float32 the distance stage
    # computed: half the upload, same neighbour lists); every D[posPath][:, posPath] of the loop (:50, :63) is a device
    # gather, so per pass only an index list goes up and the embedding comes back
    stored = data.raw('D') if isinstance(data, myio.Record) else data['D']
    nS = stored.shape[1]
    resident = [DMembeddingII.upload(stored)]
    try:
        if isinstance(posPath, int) and posPath == 0:             # :48-49
            posPath = np.arange(nS)
        else:
            posPath = np.asarray(posPath)
            resident.append(DMembeddingII.take(resident[-1], posPath))
            nS = len(posPath)
        D = resident[-1]                                          # D[posPath][:, posPath]
        lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.embed(D, nS, tune)     # :54, k = nS
        keep = get_psiPath(psi, rad, 0)                           # indices into D
        while len(keep) < nS:                                     # :61-70 drop the outliers, embed what is left
            nS = len(keep)
            sub = DMembeddingII.take(D, keep)
            try:
                lamb, psi, sigma, mu, logEps, logSumWij, popt, R_squared = DMembeddingII.embed(sub, nS, tune)
            finally:
                sub.free()
            lamb = lamb[lamb > 0]
            keep = keep[get_psiPath(psi, rad, 0)]
    finally:
        for a in resident:
            a.free()
    posPath = posPath[keep]
    if doSave['Is']:                                              # :79-87
        myio.fout1(psi_file, ['lamb', 'psi', 'sigma', 'mu', 'posPath', 'ind', 'logEps', 'logSumWij', 'popt', 'R_squared'],
                   [lamb, psi, sigma, mu, posPath, ind, logEps, logSumWij, popt, R_squared], layout='pickle')
        open(os.path.join(p.psi_prog, '%s' % (prD)), 'a').close()  # marker after the dump
    if os.path.exists(eig_file):                                  # :89-92
        os.remove(eig_file)
    with open(eig_file, 'a') as f:
        for i in range(len(lamb) - 1):
            f.write("%d\t%.5f\n" % (i + 1, lamb[i + 1]))
    return None
