"""Drop-in for the driver of the psi-analysis stage (modules/psiAnalysis.py:52-115 and its _mpi twin :37-76).

op(*argv): argv[0], if present, has .emit(int_percent) (Qt signal `progress3`).  One job per projection direction
[dist_file, psi_file, psi2_file, EL_file, psinums, senses, prD, psi_list] (divid, :30-50), skipping the (PD, psi) pairs
whose marker exists under p.psi2_prog (resume, :20-28); the PDs are sharded over the visible B200s with a static
cost-balanced partition (one spawned process per GPU, no collective — the results are the per-psi files and markers), each
worker runs psiAnalysisParS2.op (first pass, isFull = 0) on its PDs.
"""
import multiprocessing
import os
import time

import numpy as np

from . import partition
from . import psiAnalysisParS2
from .getDistanceCTF_local_Conj9combinedS2 import _cfg
from .GetDistancesS2 import _n_gpus, _set_params

fileCheck = psiAnalysisParS2.fileCheck


def divid(N, rc, fin_PDs):
    """(:30-50) job list; psi_list = the psis of the PD that are not finished yet."""
    p = _cfg()
    ll = []
    for prD in range(N):
        psinums = rc['psiNumsAll'][prD, :]
        senses = rc['sensesAll'][prD, :]
        psi_list = [psi for psi in range(len(psinums)) if fin_PDs[int(prD), int(psi)] != 1]
        ll.append(['{}prD_{}'.format(p.dist_file, prD), '{}prD_{}'.format(p.psi_file, prD),
                   '{}prD_{}'.format(p.psi2_file, prD), '{}prD_{}'.format(p.EL_file, prD), psinums, senses, prD, psi_list])
    return ll


_CFG_KEYS = ('tune', 'nClass', 'num_psis', 'num_eigs', 'numberofJobs', 'psi2_prog', 'EL_prog', 'record_layout', 'eig_solver')


def _gpu_worker(device, jobs, conOrderRange, trajName, psiTrunc, cfg):
    os.environ['MANIFOLDEM_B200_DEVICE'] = str(device)
    p = _cfg()
    for k, v in cfg.items():
        setattr(p, k, v)
    for job in jobs:
        psiAnalysisParS2.op(job, conOrderRange, trajName, 0, psiTrunc)


def op(*argv):
    p = _cfg()
    _set_params(1)
    psiNumsAll = np.tile(np.array(range(p.num_psis)), (p.numberofJobs, 1))           # :58-60
    sensesAll = np.tile(np.ones(p.num_psis), (p.numberofJobs, 1))
    rc = {'psiNumsAll': psiNumsAll, 'sensesAll': sensesAll}
    print("Computing the NLSA snapshots...")
    fin_PDs = fileCheck(p.numberofJobs)
    input_data = [job for job in divid(p.numberofJobs, rc, fin_PDs) if len(job[7])]
    progress3 = argv[0] if argv else None
    total = float(p.numberofJobs * p.num_psis)

    def emit():
        if progress3 is not None:
            progress3.emit(int((np.count_nonzero(fileCheck(p.numberofJobs) == 1) / total) * 100))
    emit()
    print("Processing {} projection directions.".format(len(input_data)))
    n_workers = min(_n_gpus(), max(1, len(input_data)))
    if n_workers <= 1:
        for job in input_data:                                                       # :93-98
            if progress3 is not None:
                psiAnalysisParS2.op(job, p.conOrderRange, p.trajName, 0, p.num_psiTrunc, progress3)
            else:
                psiAnalysisParS2.op(job, p.conOrderRange, p.trajName, 0, p.num_psiTrunc)
    else:
        # cost of a PD ~ psis left x particles^2 (ConD, the two embeddings and the Gram matrix of the snapshots)
        costs = []
        for job in input_data:
            try:
                n = os.path.getsize(job[1])
            except OSError:
                n = 1
            costs.append(float(len(job[7])) * max(n, 1))
        shards = partition.lpt_partition(costs, n_workers)
        cfg = {k: getattr(p, k) for k in _CFG_KEYS if hasattr(p, k)}
        ctx = multiprocessing.get_context('spawn')
        procs = [ctx.Process(target=_gpu_worker, args=(r, [input_data[i] for i in shards[r]], p.conOrderRange, p.trajName,
                                                       p.num_psiTrunc, cfg)) for r in range(n_workers)]
        for pr in procs:
            pr.start()
        while any(pr.is_alive() for pr in procs):
            emit()
            time.sleep(0.2)
        for pr in procs:
            pr.join()
            if pr.exitcode != 0:
                raise RuntimeError('GPU worker exited with code %s' % pr.exitcode)
        emit()
    _set_params(0)
    return
