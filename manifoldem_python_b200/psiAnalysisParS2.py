"""Drop-in for the per-PD worker of the psi analysis (modules/psiAnalysisParS2.py:45-170).

op(input_data, conOrderRange, traj_name, isFull, psiTrunc, *argv) -> 'ok'       same signature, files and markers as the reference

The PD record (D, imgAll, msk2, CTF) and the embedding (psi, posPath) are read through myio; the PD goes to the device ONCE
(NLSA.PdState: D, CTF-weighted spectra, CTF half planes) and every psi of psi_list is analysed on that state.  First pass
(isFull = 0): only the <= nClass class representatives of IMGT are downloaded.  Output keys / file names / resume markers as
:127-163."""
import os

import numpy as np

from . import myio, NLSA
from .getDistanceCTF_local_Conj9combinedS2 import _cfg


def corr(a, b, n, m):
    A = a[:, n] - np.mean(a[:, n])
    B = b[:, m] - np.mean(b[:, m])
    return np.dot(A, B) / (np.std(A) * np.std(B))


def diff_corr(a, b, maxval):
    return corr(a, b, 0, 0) + corr(a, b, maxval, maxval) - (corr(a, b, 0, maxval) + corr(a, b, maxval, 0))


def fileCheck(N):
    p = _cfg()
    fin_PDs = np.zeros(shape=(N, p.num_psis), dtype=int)
    for root, dirs, files in os.walk(p.psi2_prog):
        for file in sorted(files):
            if not file.startswith('.'):
                fin_PD, fin_psi = file.split('_')
                fin_PDs[int(fin_PD), int(fin_psi)] = int(1)
    return fin_PDs


def class_representatives(tau, numclass):
    """:100-124 — for each of numclass tau bins the first snapshot inside (bins widened while empty) -> tauinds."""
    tauinds = []
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
        tauinds.append(tauind[0])
    return tauinds


def op(input_data, conOrderRange, traj_name, isFull, psiTrunc, *argv):
    p = _cfg()
    dist_file, psi_file, psi2_file, EL_file, psinums, senses, prD = input_data[:7]
    psi_list = input_data[7] if len(input_data) == 8 else psinums
    data_IMG = myio.fin1(dist_file)
    data_psi = myio.fin1(psi_file)
    D = np.asarray(data_IMG['D'])
    imgAll = np.asarray(data_IMG['imgAll'])
    msk2 = np.array(data_IMG['msk2'])
    CTF = np.asarray(data_IMG['CTF'])
    psi = data_psi['psi']
    posPath = np.squeeze(data_psi['posPath'])
    nS = len(posPath)
    ConOrder = int(np.floor(float(nS) / conOrderRange))
    dim = int(np.sqrt(imgAll.size / D.shape[0]))
    CTF = CTF.reshape(D.shape[0], dim, dim)
    state = NLSA.PdState(D, imgAll, CTF)                  # D[posPath][:, posPath][PosPsi1][:, PosPsi1] = D[sel][:, sel]
    try:
        for psinum in psi_list:
            if psinum == -1:
                continue
            PosPsi1 = np.argsort(psi[:, psinum])
            sel = posPath[PosPsi1]
            num = nS
            NLSAPar = dict(num=num, ConOrder=ConOrder, k=num - ConOrder, tune=p.tune, nS=nS, save=False, psiTrunc=psiTrunc)
            IMGT_d, Topo_mean, psirec, psiC1, sdiag, VX, mu, tau = NLSA.analyse(state, sel, sel, NLSAPar, msk2,
                                                                                keep_IMGT_on_device=True)
            nSrecon = min(IMGT_d.shape)
            numclass = int(min(p.nClass, np.floor(nSrecon / 2.)))
            tau = (tau - min(tau)) / (max(tau) - min(tau))
            tauinds = class_representatives(tau, numclass)
            IMG1 = NLSA.frames(IMGT_d, tauinds)
            if isFull:                                     # second pass for EL1D (:127-150)
                data = myio.fin1('{}_psi_{}'.format(psi2_file, psinum))
                dc = diff_corr(IMG1, data['IMG1'], numclass - 1)
                if (senses[0] == -1 and dc > 0) or senses[0] == 1 and dc < 0:
                    tau = 1 - tau
                IMGT = np.ascontiguousarray(IMGT_d.download().T)
                myio.fout1('{}_{}_1'.format(EL_file, traj_name),
                           ['IMG1', 'IMGT', 'posPath', 'PosPsi1', 'psirec', 'tau', 'psiC1', 'mu', 'VX', 'sdiag', 'Topo_mean', 'tauinds'],
                           [IMG1, IMGT, posPath, PosPsi1, psirec, tau, psiC1, mu, VX, sdiag, Topo_mean, tauinds])
                open(os.path.join(p.EL_prog, '%s' % (prD)), 'a').close()
            else:                                          # first pass (:152-170)
                myio.fout1('{}_psi_{}'.format(psi2_file, psinum),
                           ['IMG1', 'psirec', 'tau', 'psiC1', 'mu', 'VX', 'sdiag', 'Topo_mean', 'tauinds'],
                           [IMG1, psirec, tau, psiC1, mu, VX, sdiag, Topo_mean, tauinds])
                open(os.path.join(p.psi2_prog, '%s_%s' % (prD, psinum)), 'a').close()
                if argv:
                    fin_PDs = fileCheck(p.numberofJobs)
                    offset = np.count_nonzero(fin_PDs == 1)
                    argv[0].emit(int((offset / float((p.numberofJobs) * p.num_psis)) * 100))
            IMGT_d.free()
    finally:
        state.free()
    return 'ok'
