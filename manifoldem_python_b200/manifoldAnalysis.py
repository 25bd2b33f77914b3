"""Drop-in for the driver of the embedding stage (modules/manifoldAnalysis.py:56-116 and its _mpi twin).

op(*argv): argv[0], if present, has .emit(int_percent) (Qt signal `progress2`).  One job per projection direction
[dist_file, psi_file, eig_file, prD] (divide, :43-54), skipping PDs whose marker exists under p.psi_prog (resume, :30-41); the
`topos/PrD_<n>` directories the eigenvalue files go to are created as the reference does (:85-87); the jobs are sharded over
the visible B200s (one spawned process per GPU, static partition by record size, no collective — markers are the gather) and
every worker runs manifoldTrimmingAuto.op(job, posPath = 0, p.tune, p.rad, visual = False, doSave)."""
import multiprocessing
import os
import time

from . import manifoldTrimmingAuto, partition
from .getDistanceCTF_local_Conj9combinedS2 import _cfg
from .GetDistancesS2 import _n_gpus, _set_params


def fileCheck():
    p = _cfg()
    fin_PDs = []
    for root, dirs, files in os.walk(p.psi_prog):
        for file in sorted(files):
            if not file.startswith('.'):
                fin_PDs.append(int(file))
    return fin_PDs


def count(N):
    return N - len(fileCheck())


def divide(N):
    p = _cfg()
    ll = []
    fin_PDs = fileCheck()
    for prD in range(N):
        if prD not in fin_PDs:
            ll.append(['{}prD_{}'.format(p.dist_file, prD), '{}prD_{}'.format(p.psi_file, prD),
                       '{}/topos/PrD_{}/eig_spec.txt'.format(p.out_dir, prD + 1), prD])
    return ll


_CFG_KEYS = ('tune', 'rad', 'num_eigs', 'psi_prog', 'record_layout', 'eig_solver')


def _gpu_worker(device, jobs, tune, rad, cfg):
    os.environ['MANIFOLDEM_B200_DEVICE'] = str(device)
    p = _cfg()
    for k, v in cfg.items():
        setattr(p, k, v)
    for job in jobs:
        manifoldTrimmingAuto.op(job, 0, tune, rad, False, dict(outputFile='', Is=True))


def op(*argv):
    p = _cfg()
    _set_params(1)
    print("Computing the eigenfunctions...")
    doSave = dict(outputFile='', Is=True)
    input_data = divide(p.numberofJobs)
    progress2 = argv[0] if argv else None
    offset = p.numberofJobs - len(input_data)
    if progress2 is not None:
        progress2.emit(int((offset / float(p.numberofJobs)) * 100))
    print("Processing {} projection directions.".format(len(input_data)))
    for i in range(p.numberofJobs):
        os.makedirs(p.out_dir + '/topos/PrD_{}'.format(i + 1), exist_ok=True)
    n_workers = min(_n_gpus(), max(1, len(input_data)))
    if n_workers <= 1:
        for job in input_data:
            manifoldTrimmingAuto.op(job, 0, p.tune, p.rad, False, doSave)
            offset += 1
            if progress2 is not None:
                progress2.emit(int((offset / float(p.numberofJobs)) * 100))
    else:
        costs = []
        for job in input_data:                       # the trimming loop embeds an nS x nS matrix: cost ~ size of the record's D
            try:
                costs.append(float(max(1, os.path.getsize(job[0]))))
            except OSError:
                costs.append(1.0)
        shards = partition.lpt_partition(costs, n_workers)
        cfg = {k: getattr(p, k) for k in _CFG_KEYS if hasattr(p, k)}
        ctx = multiprocessing.get_context('spawn')
        procs = [ctx.Process(target=_gpu_worker, args=(r, [input_data[i] for i in shards[r]], p.tune, p.rad, cfg))
                 for r in range(n_workers)]
        for pr in procs:
            pr.start()
        while any(pr.is_alive() for pr in procs):
            if progress2 is not None:
                progress2.emit(int(((p.numberofJobs - count(p.numberofJobs)) / float(p.numberofJobs)) * 100))
            time.sleep(0.2)
        for pr in procs:
            pr.join()
            if pr.exitcode != 0:
                raise RuntimeError('GPU worker exited with code %s' % pr.exitcode)
        if progress2 is not None:
            progress2.emit(int(((p.numberofJobs - count(p.numberofJobs)) / float(p.numberofJobs)) * 100))
    _set_params(0)
    return
