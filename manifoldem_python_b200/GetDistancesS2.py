"""Drop-in for the distance-stage driver (modules/GetDistancesS2.py:53-124 and its _mpi twin :37-96).

op(*argv): argv[0], if present, has .emit(int_percent) (Qt signal).  Reads the tessellation pickle
p.tess_file (keys CG, df, q, sh), builds one job per projection direction [ind, q1, df1, dist_file, prD],
skips PDs whose marker exists in p.dist_prog (resume), and runs the rest on the visible B200s:
one spawned process per GPU (never fork after CUDA init), one job queue for the box (or a static LPT partition), no collective —
results are the per-PD pickles, progress is read back from the marker directory exactly as the
reference's MPI mode does (:69-73).
"""
import multiprocessing
import os
import time

import numpy as np

from . import myio, partition
from . import getDistanceCTF_local_Conj9combinedS2 as worker
from .getDistanceCTF_local_Conj9combinedS2 import _cfg


def fileCheck():
    """Finished PDs = names of the marker files under p.dist_prog (:29-35)."""
    p = _cfg()
    fin = []
    for _root, _dirs, files in os.walk(p.dist_prog):
        for f in sorted(files):
            if not f.startswith('.'):
                fin.append(int(f))
    return fin


def divide(CG, q, df, N):
    """(:37-47) job list, skipping PDs with a marker."""
    p = _cfg()
    fin = set(fileCheck())
    ll = []
    for prD in range(N):
        ind = CG[prD]
        if prD not in fin:
            ll.append([ind, q[:, ind], df[ind], '{}prD_{}'.format(p.dist_file, prD), prD])
    return ll


def count(N):
    return N - len(fileCheck())


def _set_params(do):
    try:
        import set_params          # the reference's persistence of p.* (modules/set_params.py), if present
        set_params.op(do)
    except ImportError:
        pass


def _n_gpus():
    env = os.environ.get('MANIFOLDEM_B200_GPUS')
    if env:
        return max(1, int(env))
    try:
        import torch
        return max(1, torch.cuda.device_count())
    except Exception:
        return 1


def _gpu_worker(device, jobs, filterPar, imgFileName, sh, size, options, cfg, queue=None, work=0.0):
    """Runs in a spawned process: restore the config, bind the device, loop over this rank's PDs (a static shard) or over
    whatever the box's job queue still holds."""
    os.environ['MANIFOLDEM_B200_DEVICE'] = str(device)
    p = _cfg()
    for k, v in cfg.items():
        setattr(p, k, v)
    if queue is not None:
        _run_queue(queue, filterPar, imgFileName, sh, size, options, work)
    else:
        _run_jobs(jobs, filterPar, imgFileName, sh, size, options, p.nPix, None)


def _inflight(work):
    """PDs in flight per GPU.  Small PDs are launch/latency-bound: 4 streams double the throughput at nS ~ 200
    (scripts/overlap_check.py); `work` = median particles per PD x box area."""
    return max(1, int(os.environ.get('MANIFOLDEM_B200_INFLIGHT', '4' if work < 3e7 else '2')))


def _run_jobs(jobs, filterPar, imgFileName, sh, size, options, nPix, on_done):
    """The PDs of one GPU, a few in flight: the float64 conversion + pickle dump of PD k (host; NumPy and file I/O
    release the GIL) overlaps the device work of PD k+1.  Every host thread owns its context (stream + workspace)."""
    from concurrent.futures import ThreadPoolExecutor, as_completed
    work = float(np.median([len(j[0]) for j in jobs])) * nPix * nPix if jobs else 0.0
    with ThreadPoolExecutor(max_workers=_inflight(work)) as pool:
        futs = [pool.submit(worker.op, job, filterPar, imgFileName, sh, size, options) for job in jobs]
        for fut in as_completed(futs):
            fut.result()                                                             # re-raise worker errors
            if on_done is not None:
                on_done()


def _run_queue(queue, filterPar, imgFileName, sh, size, options, work, op=None):
    """One GPU's share of the box's job queue — the reference hands its PDs out the same way (Pool.imap_unordered,
    GetDistancesS2.py:110-113): every in-flight slot takes the next PD when it is free, so a GPU whose host link delivers
    more (profiles/r02_h2d_concurrent_8gpu.txt: 23 against 35 GB/s on one box) simply takes more PDs.  The parent puts one
    end marker (None) per slot of every process behind the jobs.  Returns the number of PDs this process ran."""
    from concurrent.futures import ThreadPoolExecutor
    run = op or worker.op

    def slot():
        n = 0
        while True:
            job = queue.get()
            if job is None:
                return n
            run(job, filterPar, imgFileName, sh, size, options)
            n += 1

    k = _inflight(work)
    with ThreadPoolExecutor(max_workers=k) as pool:
        futs = [pool.submit(slot) for _ in range(k)]
        return sum(f.result() for f in futs)                                        # re-raises worker errors


_CFG_KEYS = ('nPix', 'pix_size', 'Cs', 'EkV', 'AmpContrast', 'gaussEnv', 'mask_vol_file', 'dist_prog', 'dist_file',
             'relion_data', 'ncpu', 'record_layout', 'record_virtual_ctf', 'record_skip')


def op(*argv):
    p = _cfg()
    _set_params(1)
    data = myio.fin1(p.tess_file)
    CG = data['CG']
    df, q, sh = data['df'], data['q'], data['sh']
    size = len(df)
    filterPar = dict(type='Butter', Qc=0.5, N=8)                                     # :83
    options = dict(verbose=False, avgOnly=False, visual=False, parallel=False,
                   relion_data=p.relion_data, thres=getattr(p, 'PDsizeThH', 2000))  # :84-85
    if p.relion_data is False:                                                       # SPIDER: box from file size (:90-92)
        p.nPix = int(np.sqrt(os.path.getsize(p.img_stack_file) / (4 * p.num_part)))
        _set_params(0)
    if not getattr(p, 'numberofJobs', 0):
        p.numberofJobs = len(CG)
    input_data = divide(CG, q, df, p.numberofJobs)
    progress = argv[0] if argv else None
    offset = p.numberofJobs - len(input_data)
    if progress is not None:
        progress.emit(int((offset / float(p.numberofJobs)) * 100))
    print('Processing {} projection directions.'.format(len(input_data)))

    # one worker process per visible B200, whatever p.ncpu says (the GUI's default p.ncpu = 1 would otherwise leave
    # seven GPUs of a box idle); MANIFOLDEM_B200_GPUS=n restricts it
    n_workers = min(_n_gpus(), max(1, len(input_data)))
    if n_workers != max(1, int(getattr(p, 'ncpu', 1))):
        print('GetDistancesS2: %d GPU worker(s) (p.ncpu = %s is a CPU setting and is not used)' % (n_workers, getattr(p, 'ncpu', 1)))
    if n_workers <= 1:
        state = {'offset': offset}                                                   # :102-108

        def on_done():
            state['offset'] += 1
            if progress is not None:
                progress.emit(int((state['offset'] / float(p.numberofJobs)) * 100))
        _run_jobs(input_data, filterPar, p.img_stack_file, sh, size, options, p.nPix, on_done)
    else:
        costs = [partition.pd_cost(len(job[0]), p.nPix) for job in input_data]
        cfg = {k: getattr(p, k) for k in _CFG_KEYS if hasattr(p, k)}
        ctx = multiprocessing.get_context('spawn')
        queue = None
        if os.environ.get('MANIFOLDEM_B200_SCHEDULE', 'queue') == 'static':
            # MANIFOLDEM_B200_GPU_SPEEDS="1,1,1,1,1.5,1.5,1.5,1.5": relative feed rate of the GPUs when the stage is bound by
            # the host-to-device copies and the box does not serve its GPUs equally
            speeds = os.environ.get('MANIFOLDEM_B200_GPU_SPEEDS')
            speeds = [float(x) for x in speeds.split(',')][:n_workers] if speeds else None
            shards = partition.lpt_partition(costs, n_workers, speeds if speeds and len(speeds) == n_workers else None)
            procs = [ctx.Process(target=_gpu_worker, args=(r, [input_data[i] for i in shards[r]], filterPar,
                                                           p.img_stack_file, sh, size, options, cfg))
                     for r in range(n_workers)]
        else:
            # one queue for the box, largest PDs first (the LPT order), as the reference's Pool.imap_unordered (:110-113)
            queue = ctx.Queue()
            for i in sorted(range(len(input_data)), key=lambda i: (-costs[i], i)):
                queue.put(input_data[i])
            work = float(np.median([len(j[0]) for j in input_data])) * p.nPix * p.nPix
            for _ in range(n_workers * _inflight(work)):
                queue.put(None)
            procs = [ctx.Process(target=_gpu_worker, args=(r, None, filterPar, p.img_stack_file, sh, size, options, cfg,
                                                           queue, work))
                     for r in range(n_workers)]
        for pr in procs:
            pr.start()
        while any(pr.is_alive() for pr in procs):                                    # marker polling, as :69-73
            if progress is not None:
                done = p.numberofJobs - count(p.numberofJobs)
                progress.emit(int((done / float(p.numberofJobs)) * 100))
            time.sleep(0.2)
        for pr in procs:
            pr.join()
        if queue is not None:
            queue.cancel_join_thread()                                               # markers of slots that died stay in the pipe
            queue.close()
        for pr in procs:
            if pr.exitcode != 0:
                raise RuntimeError('GPU worker exited with code %s' % pr.exitcode)
        if progress is not None:
            progress.emit(int(((p.numberofJobs - count(p.numberofJobs)) / float(p.numberofJobs)) * 100))
    _set_params(0)
    return
