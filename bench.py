#!/usr/bin/env python
"""bench.py — CTF pairwise distances/s of the per-PD distance stage on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W]                    one process per GPU (torchrun for N>1)
  python bench.py --impl reference [--gpus N --steps K --warmup W]   the reference's CPU algorithm (oracle port)

Workload (weak scaling): every rank processes PDS_PER_GPU synthetic projection directions of BASELINE config 4's
shape (2,000 particles x 256^2) per step; at 8 GPUs one step is the whole 1,000-PD S2 run.  A "pair" is one entry of
the nS x nS matrix D the reference materialises.

`value`      whole-job G pairs/s with the raw particle stacks already resident in HBM (CUDA events, max over ranks).
`e2e`        the same through the host-buffer C-ABI call (pinned host stacks -> D on the host), H2D and D2H inside
             the timed region, three PDs in flight per GPU; `h2d_ceiling` = the box's concurrent pinned H2D rate
             measured by a bare copy loop on every rank at the same moment (the end-to-end roofline).
`roofline`   dominant kernel (tcgen05 contraction): executed TF32 flops per launch / its CUDA-event time, against half
             the measured bf16 rate of MEASURED_PEAKS.json; the SM clock inside the kernel from a clock64 probe.
`hbm_roofline`  the HBM-bound pre-processing stages: algorithmic bytes (20 N^2 per image) / stage time.
`details`    other BASELINE configs (1: the demo's 53 PDs, 2, 3, 5: one PD) and the record-producing variants.
"""
import argparse
import contextlib
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NS, NPIX = 2000, 256
PDS_PER_GPU = 125                 # 1000 PDs / 8 GPUs (BASELINE config 4)
POOL = 4                          # distinct raw stacks cycled through (each 524 MB >> 126 MB L2)
EM = dict(pix_size=1.255, Cs=2.26, EkV=300.0, AmpContrast=0.1)
METRIC = 'CTF pairwise distances/sec'
UNIT = 'Gpairs/s'
REF_VS_PORT = os.path.join(ROOT, 'profiles', 'r02_reference_vs_port_cpu.txt')
CONTRACT_NCU = os.path.join(ROOT, 'profiles', 'r02_contract_tc2_ncu.json')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--pds', type=int, default=PDS_PER_GPU, help='PDs per GPU per step')
    ap.add_argument('--nS', type=int, default=NS)
    ap.add_argument('--N', type=int, default=NPIX)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-details', action='store_true', help='skip the other BASELINE configs / record variants')
    return ap.parse_args()


def workload_config(nS, N, P):
    """The `config` object — identical in the B200 arm and the reference arm."""
    return dict(workload='BASELINE config 4 shape: PDs of %d particles x %d^2, %d PDs per GPU per step '
                         '(1000 PDs at 8 GPUs)' % (nS, N, P), pds_per_gpu=P, nS=nS, N=N,
                l2='inputs larger than L2: %d distinct %d MB stacks cycled' % (POOL, nS * N * N * 4 // 1000000))


# ----------------------------------------------------------------------------------------- helpers
def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'MEASURED_PEAKS.json'
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback of B200_PROFILING.md'


def reference_vs_port():
    """Ratio (unmodified reference op() time / oracle port time) measured in the build container, where
    /root/reference exists (scripts/ref_vs_port_cpu.py -> profiles/r02_reference_vs_port_cpu.txt)."""
    try:
        vals = []
        for ln in open(REF_VS_PORT):
            if 'reference / port =' in ln:
                vals.append(float(ln.split('reference / port =')[1].split(';')[0]))
        return dict(ratio_mean=float(np.mean(vals)), ratios=vals, file=os.path.relpath(REF_VS_PORT, ROOT),
                    where='build container (the reference does not travel to the GPU box)') if vals else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measure_cublas_tf32(index):
    """cuBLAS TF32 GEMM (8192^3, best of 5 after warm-up) measured in this run: context for the reader."""
    try:
        import torch
        dev = torch.device('cuda', index)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c < os.cpu_count()]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def make_inputs(nS, N, pool, seed):
    """Synthetic PD stacks of the named shape: white-noise particles (timing does not depend on the
    pixel values), random defocus 1-3 um, orientations scattered 0.03 rad around one PD."""
    from manifoldem_python_b200 import pd_stage, synthetic
    rng = np.random.default_rng(seed)
    pds = []
    for _ in range(pool):
        q = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS), 1.1 + 0.03 * rng.standard_normal(nS),
                                    rng.uniform(0, 2 * np.pi, nS))
        df = rng.uniform(10000.0, 30000.0, nS)
        PDs, PD, psi_p, Psi, s, c = pd_stage.host_angles(q)
        pds.append(dict(df=df, psi_deg=np.ascontiguousarray(-(180 / np.pi) * Psi), psi_p=float(psi_p),
                        flip=(rng.random(nS) < 1 / 3).astype(np.uint8)))
    return pds, rng


def pd_params(_lib, nS, N, psi_p):
    return _lib.PdParams(nS=nS, N=N, transposed=1, relion_shift=0, filter_type=0, filter_order=8, filter_Qc=0.5,
                         pix_size=EM['pix_size'], Cs=EM['Cs'], EkV=EM['EkV'], gaussEnv=float('inf'),
                         AmpContrast=EM['AmpContrast'], psi_p_deg=psi_p, avg_only=0, contraction=0,
                         k_chunk_blocks=0, split_k=0)


# ----------------------------------------------------------------------------------------- reference arm
def _ref_images_worker(args):
    """Per-image part of the reference algorithm (ingest .. FFT) for a slice of particles; returns
    (wall seconds, per-stage seconds)."""
    nS_s, N, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from manifoldem_python_b200 import synthetic
    from oracle import pd_distance as opd
    rng = np.random.default_rng(seed)
    stack = rng.standard_normal(nS_s * N * N).astype(np.float32)
    q = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS_s), 1.1 + 0.03 * rng.standard_normal(nS_s),
                                rng.uniform(0, 2 * np.pi, nS_s))
    df = rng.uniform(10000.0, 30000.0, nS_s)
    tm = {}
    t0 = time.perf_counter()
    opd.pd_distance(np.arange(nS_s), q, df, stack, 2 * nS_s, N, EM['pix_size'], EM['Cs'], EM['EkV'], EM['AmpContrast'],
                    avg_only=True, rotate_impl='tile', keep=('imgAvg',), timings=tm)
    return time.perf_counter() - t0, tm


def _gemm_block(rows, N, seed):
    rng = np.random.default_rng(seed)
    K = N * N
    CTF = rng.standard_normal((rows, K))
    fy = rng.standard_normal((rows, K)) + 1j * rng.standard_normal((rows, K))
    t0 = time.perf_counter()
    CTFfy = CTF.conj() * fy
    D = np.dot(np.abs(CTF) ** 2, (np.abs(fy) ** 2).T)
    D = D + D.T - 2 * np.real(np.dot(CTFfy, CTFfy.conj().T))
    return time.perf_counter() - t0


def _ref_gemm_worker(args):
    """The two GEMMs of :391-397 on a rows x rows block with ONE BLAS thread (a Pool / MPI worker's share)."""
    rows, N, seed = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    return _gemm_block(rows, N, seed)


def _pickle_seconds(nS_s, N):
    """myio.fout1 (:409-412) of the per-image float64 record arrays of nS_s particles (3 x nS_s x N^2 x 8 bytes)."""
    import pickle
    import tempfile
    a = np.zeros((3, nS_s, N * N))
    with tempfile.NamedTemporaryFile(dir='/dev/shm' if os.path.isdir('/dev/shm') else None) as f:
        t0 = time.perf_counter()
        pickle.dump(dict(a=a), f, protocol=pickle.HIGHEST_PROTOCOL)
        f.flush()
        return time.perf_counter() - t0


def cpu_reference_sample(nS, N, n_img, cores, contraction_rows):
    """Bounded sample of the reference's CPU algorithm on one PD of shape (nS, N):
    (1) the per-image stages (ingest, low-pass, 2x rotatefill on the 3x3 tile, CTF, 3 FFTs) on n_img particles,
        spread over `cores` worker processes — these stages are linear in nS;
    (2) the two GEMMs of :391-397 (float64 / complex128, all BLAS threads) on `contraction_rows`^2 pairs;
    (3) the pickle dump of the per-image record arrays of n_img particles (linear in nS).
    Per-PD time = nS/n_img * t_images + (nS/contraction_rows)^2 * t_gemm + nS/n_img * t_pickle: the reference's
    serial mode (p.ncpu = 1) with the per-image loop ideally parallel.  Also timed: the Pool mode
    (GetDistancesS2.py:110-113) and its MPI twin's static round-robin (GetDistancesS2_mpi.py:14-15) — one whole PD per
    worker process, one BLAS thread each, `cores` PDs at once; for equal PDs the two schedules coincide.
    Returns (pairs/s of the faster mode, detail dict with the per-stage split)."""
    import multiprocessing as mp
    per = max(1, n_img // cores)
    jobs = [(per, N, 100 + i) for i in range(max(1, n_img // per))]
    rows1 = min(contraction_rows, 200)
    with mp.get_context('spawn').Pool(min(cores, len(jobs))) as pool:
        res = pool.map(_ref_images_worker, jobs)
        t_img = max(r[0] for r in res)                      # slowest worker's compute time (spawn/import excluded)
        t_gemm1 = max(pool.map(_ref_gemm_worker, [(rows1, N, 7 + i) for i in range(min(cores, len(jobs)))]))
    stages = {}
    for _, tm in res:
        for k, v in tm.items():
            stages[k] = max(stages.get(k, 0.0), v)
    n_done = per * len(jobs)
    rows = contraction_rows
    try:                                  # torchrun exports OMP_NUM_THREADS=1: give the GEMMs every core back
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=cores)
    except Exception:
        limiter = None
    t_gemm = _gemm_block(rows, N, 0)
    if limiter is not None:
        limiter.restore_original_limits()
    t_pick = _pickle_seconds(min(per, 32), N) * per / min(per, 32)
    # rows x rows block measured; the full matrix has (nS/rows)^2 such blocks
    s_img, s_gemm, s_pick = (nS / n_done) * t_img, (nS / rows) ** 2 * t_gemm, (nS / per) * t_pick
    t_pd = s_img + s_gemm + s_pick
    # Pool / MPI-schedule emulation: a worker runs its PD alone (per images in t_img, one BLAS thread), `workers` at once
    workers = min(cores, len(jobs))
    t_pd_worker = (nS / per) * t_img + (nS / rows1) ** 2 * t_gemm1 + (nS / per) * t_pick
    v_serial, v_pool = nS * nS / t_pd / 1e9, workers * nS * nS / t_pd_worker / 1e9
    scale = (nS / per)
    detail = dict(images=n_done, t_images_s=round(t_img, 3), gemm_rows=rows, t_gemm_s=round(t_gemm, 3),
                  per_pd_s=round(t_pd, 2),
                  stage_split_per_pd_s=dict(per_image_stages=round(s_img, 2), gemm=round(s_gemm, 2), pickle=round(s_pick, 2)),
                  per_image_stage_split_one_worker_per_pd_s={k: round(v * scale, 2) for k, v in stages.items()},
                  serial_mode_gpairs_s=v_serial,
                  pool_and_mpi_schedule_emulation=dict(gpairs_s=v_pool, workers=workers, blas_threads_per_worker=1,
                                                       per_pd_per_worker_s=round(t_pd_worker, 1), gemm_rows=rows1,
                                                       t_gemm_s=round(t_gemm1, 3)))
    return max(v_serial, v_pool), detail, t_pd


def cpu_split_other_configs(cores):
    """serial / Pool(MPI-schedule) split of the reference algorithm for config 2 (1,000 x 128^2) and for a
    median demo PD of config 1 (206 x 256^2), on bounded samples (2 images per core)."""
    out = {}
    for name, nS, N in (('config_2', 1000, 128), ('config_1_median_pd', 206, 256)):
        try:
            v, d, _ = cpu_reference_sample(nS, N, 2 * cores, cores, contraction_rows=min(nS, 400))
            out[name] = dict(nS=nS, N=N, gpairs_s=v, per_pd_s=d['per_pd_s'], stage_split_per_pd_s=d['stage_split_per_pd_s'],
                             per_image_stage_split_one_worker_per_pd_s=d['per_image_stage_split_one_worker_per_pd_s'],
                             serial_mode_gpairs_s=d['serial_mode_gpairs_s'],
                             pool_and_mpi_schedule_gpairs_s=d['pool_and_mpi_schedule_emulation']['gpairs_s'])
        except Exception as e:
            out[name] = dict(error=repr(e))
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nS, N = args.nS, args.N
    vals, walls = [], []
    detail = None
    n_img = 4 * cores
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, detail, t_pd = cpu_reference_sample(nS, N, n_img, cores, contraction_rows=min(nS, 500))
        if i >= args.warmup:
            vals.append((v, nS * nS / (v * 1e9)))     # seconds per PD at the reported rate
            walls.append(time.perf_counter() - t0)
    v = float(np.mean([a for a, _ in vals]))
    sample = ('%d of %d particles through the per-image stages on %d processes + a %dx%d block of the fp64 '
              'dgemm/zgemm + the pickle of those particles; per-PD time extrapolated linearly in images and quadratically '
              'in the block; value = the faster of the serial mode (all BLAS threads) and the Pool / MPI-schedule emulation '
              '(one PD per worker); ms_per_step = wall time of one such sample'
              % (detail['images'], nS, cores, detail['gemm_rows'], detail['gemm_rows']))
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean(walls)) * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=workload_config(nS, N, args.pds),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind='port', sample=sample, detail=detail,
                                  extrapolated_step_ms=float(np.mean([b for _, b in vals])) * 1e3 * args.pds * max(1, args.gpus),
                                  unmodified_reference_vs_port=reference_vs_port()),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- B200 arm: side workloads
def _timed_pds(ctx, lib, _lib, prms, ios, warm, reps):
    for k in range(warm + reps):
        if k == warm:
            ctx.sync()
            ctx.timer_start()
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prms[k % len(prms)]), C.byref(ios[k % len(ios)]), None))
    return ctx.timer_stop() / reps


def config1_block(ctx, lib, _lib, local, N):
    """BASELINE config 1: the demo's 53 projection directions (117..450 particles; bookkeeping from the star through the
    reference's reader and tessellation, tests/golden/demo_config1.npz) with synthetic images at box N, inputs
    resident; one stream, and 2 PDs in flight on 2 streams."""
    from manifoldem_python_b200 import workloads
    pds, em = workloads.demo_config1()
    rng = np.random.default_rng(3)
    dev = []
    for pd in pds:
        n = len(pd['ind'])
        prm = pd_params(_lib, n, N, pd['psi_p'])
        prm.pix_size, prm.Cs, prm.EkV, prm.AmpContrast = em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
        d = dict(raw=_lib.DeviceArray(ctx, (n, N * N), np.float32, rng.standard_normal((n, N * N), dtype=np.float32)),
                 flip=_lib.DeviceArray(ctx, (n,), np.uint8, pd['flip']), psi=_lib.DeviceArray(ctx, (n,), np.float64, pd['psi_deg']),
                 df=_lib.DeviceArray(ctx, (n,), np.float64, pd['df']), D=_lib.DeviceArray(ctx, (n, n), np.float32), prm=prm)
        io = _lib.PdIO()
        io.raw, io.flip, io.psi_deg, io.df, io.D = d['raw'].ptr, d['flip'].ptr, d['psi'].ptr, d['df'].ptr, d['D'].ptr
        d['io'] = io
        dev.append(d)
    pairs = float(sum(len(pd['ind']) ** 2 for pd in pds))
    images = sum(len(pd['ind']) for pd in pds)
    order = sorted(range(len(pds)), key=lambda i: -len(pds[i]['ind']))
    out = dict(workload='BASELINE config 1: the demo star\'s 53 PDs (117..450 particles, %d in all, %d pairs), synthetic '
                        'images at %d^2, inputs resident' % (images, int(pairs), N), N=N, pds=len(pds))
    for S in (1, 2):
        ctxs = [ctx] + [_lib.Context(local) for _ in range(S - 1)]
        best = 1e9
        for rep in range(4):
            for c in ctxs:
                c.sync()
            t0 = time.perf_counter()
            for k, i in enumerate(order):
                _lib.check(lib.mem_pd_distance_device(ctxs[k % S].handle, C.byref(dev[i]['prm']), C.byref(dev[i]['io']), None))
            for c in ctxs:
                c.sync()
            if rep:
                best = min(best, time.perf_counter() - t0)
        out['streams_%d' % S] = dict(ms=best * 1e3, gpairs_s=pairs / best / 1e9, images_per_s=images / best)
        for c in ctxs[1:]:
            c.close()
    # all 53 PDs as ONE group: per-image stages once over the concatenated stack, one grouped tcgen05 launch
    try:
        sizes = np.array([len(pd['ind']) for pd in pds], dtype=np.int64)
        start = np.zeros(len(pds) + 1, dtype=np.int32)
        start[1:] = np.cumsum(sizes)
        nall = int(start[-1])
        b_raw = _lib.DeviceArray(ctx, (nall, N * N), np.float32, rng.standard_normal((nall, N * N), dtype=np.float32))
        b_flip = _lib.DeviceArray(ctx, (nall,), np.uint8, np.concatenate([pd['flip'] for pd in pds]))
        b_psi = _lib.DeviceArray(ctx, (nall,), np.float64, np.concatenate([pd['psi_deg'] for pd in pds]))
        b_df = _lib.DeviceArray(ctx, (nall,), np.float64, np.concatenate([pd['df'] for pd in pds]))
        b_D = _lib.DeviceArray(ctx, (int((sizes ** 2).sum()),), np.float32)
        bprm = pd_params(_lib, nall, N, 0.0)
        bprm.pix_size, bprm.Cs, bprm.EkV, bprm.AmpContrast = em['pix_size'], em['Cs'], em['EkV'], em['AmpContrast']
        bio = _lib.PdIO()
        bio.raw, bio.flip, bio.psi_deg, bio.df, bio.D = b_raw.ptr, b_flip.ptr, b_psi.ptr, b_df.ptr, b_D.ptr
        pp = np.ascontiguousarray([pd['psi_p'] for pd in pds], dtype=np.float64)
        best = 1e9
        for rep in range(4):
            ctx.sync()
            t0 = time.perf_counter()
            _lib.check(lib.mem_pd_distance_batch_device(ctx.handle, C.byref(bprm), C.byref(bio), len(pds), start.ctypes.data,
                                                        pp.ctypes.data, None))
            ctx.sync()
            if rep:
                best = min(best, time.perf_counter() - t0)
        out['batched'] = dict(ms=best * 1e3, gpairs_s=pairs / best / 1e9, images_per_s=images / best, stage_ms=ctx.timings(),
                              how='mem_pd_distance_batch_device: one launch sequence over the concatenated stack + one grouped tcgen05 launch')
        for a in (b_raw, b_flip, b_psi, b_df, b_D):
            a.free()
    except Exception as e:
        out['batched'] = dict(error=repr(e))
    # what bounds it: the per-image pre-processing (20 N^2 algorithmic bytes per image), not the pairs — a demo PD has
    # ~230 pairs per image where config 4 has 2,000
    best = min([out['streams_1']['ms'], out['streams_2']['ms']] + ([out['batched']['ms']] if 'ms' in out.get('batched', {}) else [])) * 1e-3
    out['gpairs_s'] = pairs / best / 1e9
    out['hbm_alg_gbs'] = 20.0 * N * N * images / best / 1e9
    for d in dev:
        for k in ('raw', 'flip', 'psi', 'df', 'D'):
            d[k].free()
    return out


def nlsa_block(ctx):
    """One psi of NLSA.op (modules/NLSA.py:23-158) through the device path: a synthetic PD with a 1-D conformational
    coordinate, 600 snapshots at 128^2, ConOrder = 12, psiTrunc = 8.  Wall clock per step (device sync after each)."""
    from manifoldem_python_b200 import NLSA, DMembeddingII, synthetic, p
    p.init()
    n, Nn = 600, 128
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 1, n))
    g = (np.arange(Nn) - Nn / 2) / Nn
    yy, xx = np.meshgrid(g, g, indexing='ij')
    img = np.stack([np.exp(-((xx - 0.2 * (ti - 0.5)) ** 2 + yy ** 2) / 0.02) for ti in t]) + 0.3 * rng.standard_normal((n, Nn, Nn))
    ctf = np.stack([synthetic.ctf_2d(Nn, df) for df in rng.uniform(10000, 30000, n)])
    flat = img.reshape(n, -1)
    sq = (flat ** 2).sum(1)
    D = np.maximum(sq[:, None] + sq[None, :] - 2 * flat @ flat.T, 0)
    np.random.seed(1)
    psi = DMembeddingII.embed(D.copy(), n, 3.0)[1]
    con = n // 50
    par = dict(num=n, ConOrder=con, k=n - con, tune=3.0, nS=n, save=False, psiTrunc=8)
    t0 = time.perf_counter()
    state = NLSA.PdState(D, img, ctf, ctx=ctx)
    t1 = time.perf_counter()
    sel = np.argsort(psi[:, 0])
    tm = {}
    for rep in range(2):
        tm = {}
        t2 = time.perf_counter()
        out = NLSA.analyse(state, sel, sel, par, 1, keep_IMGT_on_device=True, timings=tm)
        t3 = time.perf_counter()
        out[0].free()
    state.free()
    return dict(workload='NLSA.op, one psi: %d snapshots x %d^2, ConOrder %d, psiTrunc 8 (the reference: %d fft2 / ifft2 pairs + '
                         'a %d^2 x %d float64 Gram + up to 101 x %d np.roots calls)' % (n, Nn, con, con * (n - con), n - 2 * con, Nn * Nn, n - 2 * con),
                upload_and_forward_transforms_ms=(t1 - t0) * 1e3, analyse_ms=(t3 - t2) * 1e3, step_ms={k: round(v, 2) for k, v in tm.items()},
                tau_range=[float(np.min(out[7])), float(np.max(out[7]))])


def config5_block(ctx, lib, _lib, local):
    """BASELINE config 5, one PD: 20,000 particles at 320^2 (73 GB of workspace on one B200), D requested."""
    import torch
    n5, N5 = 20000, 320
    free_b, _tot = torch.cuda.mem_get_info(local)
    if free_b < 110e9:
        return dict(skipped='needs ~85 GB of free HBM, %.0f GB free' % (free_b / 1e9))
    raw = torch.randn(n5, N5 * N5, device='cuda:%d' % local, dtype=torch.float32)
    pds5, _ = make_inputs(n5, N5, 1, seed=79)
    aux = [_lib.DeviceArray(ctx, (n5,), np.uint8, pds5[0]['flip']), _lib.DeviceArray(ctx, (n5,), np.float64, pds5[0]['psi_deg']),
           _lib.DeviceArray(ctx, (n5,), np.float64, pds5[0]['df'])]
    D5 = _lib.DeviceArray(ctx, (n5, n5), np.float32)
    prm = pd_params(_lib, n5, N5, pds5[0]['psi_p'])
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df, io.D = raw.data_ptr(), aux[0].ptr, aux[1].ptr, aux[2].ptr, D5.ptr
    torch.cuda.synchronize()
    ms = _timed_pds(ctx, lib, _lib, [prm], [io], 1, 2)
    st = ctx.timings()
    out = dict(workload='BASELINE config 5, one PD: 20,000 particles x 320^2, full D', per_pd_ms=ms,
               gpairs_s=n5 * n5 / (ms * 1e-3) / 1e9, stage_ms=st, pds_in_config=200,
               whole_config_estimate_s=200 * ms * 1e-3)
    for a in aux + [D5]:
        a.free()
    del raw
    torch.cuda.empty_cache()
    return out


def h2d_ceiling(ctx, _lib, lib, local, barrier, sources=None):
    """Concurrent pinned-host -> device copy rate of this rank while every other rank does the same (the roofline of the
    e2e leg: 524 MB of raw particles per PD).  Returns (rate over the e2e leg's own pinned source buffers, each copied
    once in turn — what the leg streams from host DRAM —, rate of 8 copies of ONE 512 MB pinned buffer — the figure a bare
    copy loop reports, part of which a large host cache can serve)."""
    n = 512 << 20
    h = _lib.PinnedArray((n,), np.uint8)
    h.array[::4096] = 1
    d = _lib.DeviceArray(ctx, (n,), np.uint8)
    _lib.check(lib.mem_copy_h2d(ctx.handle, d.ptr, h.ptr, n))
    barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        _lib.check(lib.mem_copy_h2d(ctx.handle, d.ptr, h.ptr, n))
    dt = time.perf_counter() - t0
    barrier()
    single = 8 * n / dt / 1e9
    pool = single
    if sources:
        m = min(n, min(int(a.array.nbytes) for a in sources))
        barrier()
        t0 = time.perf_counter()
        for a in sources:
            _lib.check(lib.mem_copy_h2d(ctx.handle, d.ptr, a.ptr, m))
        dt = time.perf_counter() - t0
        barrier()
        pool = len(sources) * m / dt / 1e9
    d.free()
    h.free()
    return pool, single


def contraction_ncu():
    """dram bytes / tensor-pipe figures of the contraction from the committed ncu capture of this round."""
    try:
        with open(CONTRACT_NCU) as f:
            d = json.load(f)
        d['file'] = os.path.relpath(CONTRACT_NCU, ROOT)
        return d
    except Exception:
        return None


# ----------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from manifoldem_python_b200 import _lib
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout at the first collective: keep stdout = the one JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = _lib.load()
    nS, N, P = args.nS, args.N, args.pds
    numa = bind_to_gpu_numa_node(local)       # pinned staging buffers are first-touched on the GPU's own socket
    ctx = _lib.Context(local)
    pds, rng = make_inputs(nS, N, POOL, seed=1000 + rank)
    NN = N * N
    # ---- inputs resident in HBM
    d_raw, d_flip, d_psi, d_df = [], [], [], []
    h_raw = []
    for j, pd in enumerate(pds):
        hr = _lib.PinnedArray((nS, NN), np.float32)
        hr.array[...] = rng.standard_normal((nS, NN), dtype=np.float32)
        h_raw.append(hr)
        d_raw.append(_lib.DeviceArray(ctx, (nS, NN), np.float32, hr.array))
        d_flip.append(_lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip']))
        d_psi.append(_lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg']))
        d_df.append(_lib.DeviceArray(ctx, (nS,), np.float64, pd['df']))
    d_D = _lib.DeviceArray(ctx, (nS, nS), np.float32)
    prms = [pd_params(_lib, nS, N, pd['psi_p']) for pd in pds]
    ios = []
    for j in range(POOL):
        io = _lib.PdIO()
        io.raw, io.flip, io.psi_deg, io.df, io.D = d_raw[j].ptr, d_flip[j].ptr, d_psi[j].ptr, d_df[j].ptr, d_D.ptr
        ios.append(io)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        for k in range(P):
            j = k % POOL
            _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prms[j]), C.byref(ios[j]), None))

    for _ in range(args.warmup):
        step()
    barrier()
    ctx.kernel_time(reset=True)
    ctx.launches(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches(reset=True)
    k_ms, k_n, k_items, k_kb = ctx.kernel_time(reset=True)
    k_mhz, k_probe_ms = ctx.kernel_clock()
    stage = ctx.timings()
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device='cuda:%d' % local)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, launches = float(tm[0]), int(ts[1])
    pairs_step = float(nS) * nS * P * world
    value = pairs_step * args.steps / (ms * 1e-3) / 1e9

    details = {}
    side = rank == 0 and not args.no_e2e and not args.no_details
    # ---- the same PDs with every array of the reference's per-PD record produced on the device as well (imgAll,
    # imgAllFlip, float64 CTF, Wiener / flip averages, intensity): reported beside the headline, which asks for D only
    if side:
        outs = [_lib.DeviceArray(ctx, (nS, NN), np.float32), _lib.DeviceArray(ctx, (nS, NN), np.float32),
                _lib.DeviceArray(ctx, (nS, NN), np.float64), _lib.DeviceArray(ctx, (NN,), np.float32),
                _lib.DeviceArray(ctx, (NN,), np.float32), _lib.DeviceArray(ctx, (NN,), np.float32)]
        fio = []
        for j in range(POOL):
            io = _lib.PdIO()
            io.raw, io.flip, io.psi_deg, io.df, io.D = d_raw[j].ptr, d_flip[j].ptr, d_psi[j].ptr, d_df[j].ptr, d_D.ptr
            (io.imgAll, io.imgAllFlip, io.CTF, io.imgAvg, io.imgAvgFlip, io.imgAllIntensity) = [o.ptr for o in outs]
            fio.append(io)
        details['all_record_fields_per_pd_ms'] = _timed_pds(ctx, lib, _lib, prms, fio, 2, 8)
        for o in outs:
            o.free()
    # ---- BASELINE config 2 (one synthetic PD of 1,000 particles at 128^2, full D), device-resident
    if side and (nS, N) == (NS, NPIX):
        n2, N2 = 1000, 128
        pds2, rng2 = make_inputs(n2, N2, 2, seed=77)
        raws2 = [_lib.DeviceArray(ctx, (n2, N2 * N2), np.float32, rng2.standard_normal((n2, N2 * N2), dtype=np.float32))
                 for _ in pds2]
        aux2 = [(_lib.DeviceArray(ctx, (n2,), np.uint8, pd['flip']), _lib.DeviceArray(ctx, (n2,), np.float64, pd['psi_deg']),
                 _lib.DeviceArray(ctx, (n2,), np.float64, pd['df'])) for pd in pds2]
        D2 = _lib.DeviceArray(ctx, (n2, n2), np.float32)
        io2, prm2 = [], []
        for j, pd in enumerate(pds2):
            io = _lib.PdIO()
            io.raw, io.flip, io.psi_deg, io.df, io.D = raws2[j].ptr, aux2[j][0].ptr, aux2[j][1].ptr, aux2[j][2].ptr, D2.ptr
            io2.append(io)
            prm2.append(pd_params(_lib, n2, N2, pd['psi_p']))
        ms2 = _timed_pds(ctx, lib, _lib, prm2, io2, 4, 40)
        details['config_2'] = dict(workload='BASELINE config 2: one PD of 1,000 particles x 128^2, full D', per_pd_ms=ms2,
                                   gpairs_s=n2 * n2 / (ms2 * 1e-3) / 1e9, stage_ms=ctx.timings())
        for a in raws2 + [x for t3 in aux2 for x in t3] + [D2]:
            a.free()
    # ---- BASELINE config 3 (one PD of 5,000 particles at 256^2, k = 100 neighbour lists only): the lists are selected
    # from the contraction's partial tiles, D is never assembled.  Particles: rows of three of the pool stacks.
    if side and (nS, N) == (NS, NPIX) and POOL * nS >= 5000:
        n3, k3 = 5000, 100
        pds3, _ = make_inputs(n3, N, 1, seed=78)
        raw3 = _lib.DeviceArray(ctx, (n3, NN), np.float32, np.concatenate([h.array for h in h_raw])[:n3])
        aux3 = [_lib.DeviceArray(ctx, (n3,), np.uint8, pds3[0]['flip']), _lib.DeviceArray(ctx, (n3,), np.float64, pds3[0]['psi_deg']),
                _lib.DeviceArray(ctx, (n3,), np.float64, pds3[0]['df'])]
        idx3, val3 = _lib.DeviceArray(ctx, (n3, k3), np.int32), _lib.DeviceArray(ctx, (n3, k3), np.float64)
        prm3 = pd_params(_lib, n3, N, pds3[0]['psi_p'])
        prm3.knn_k = k3
        io3 = _lib.PdIO()
        io3.raw, io3.flip, io3.psi_deg, io3.df = raw3.ptr, aux3[0].ptr, aux3[1].ptr, aux3[2].ptr
        io3.knn_idx, io3.knn_val = idx3.ptr, val3.ptr
        ms3 = _timed_pds(ctx, lib, _lib, [prm3], [io3], 2, 5)
        nb = idx3.download()
        assert np.array_equal(nb[:, 0], np.arange(n3)) and nb.min() >= 0 and nb.max() < n3
        details['config_3'] = dict(workload='BASELINE config 3: one PD of 5,000 particles x 256^2, k = 100 neighbour lists only (no D)',
                                   per_pd_ms=ms3, gpairs_s=n3 * n3 / (ms3 * 1e-3) / 1e9, knn_k=k3, stage_ms=ctx.timings())
        for a in [raw3, idx3, val3] + aux3:
            a.free()
    # ---- BASELINE config 1 (the demo's 53 PDs) at 256^2 and 128^2
    if side and (nS, N) == (NS, NPIX):
        try:
            details['config_1'] = dict(N256=config1_block(ctx, lib, _lib, local, 256), N128=config1_block(ctx, lib, _lib, local, 128))
        except Exception as e:
            details['config_1'] = dict(error=repr(e))
    # ---- rows a15-a18 on the D the last PD left on the device (k = nS, the way manifoldTrimmingAuto calls
    # DMembeddingII.op): kNN, graph, Ferguson sweep, Gaussian-kernel Laplacian; wall clock around the device chain
    if side:
        try:
            from manifoldem_python_b200 import DMembeddingII
            _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prms[0]), C.byref(ios[0]), None))
            ctx.sync()
            sig = 3.0 * float(np.sqrt(np.median(d_D.download()[0, 1:])))
            for rep in range(2):                     # second pass = warm workspaces
                t0 = time.perf_counter()
                M_dev, _le, _ls, _i, _v = DMembeddingII.graph_and_sweep(d_D, nS, ctx=ctx, want_lists=False)
                t1 = time.perf_counter()
                L_dev = DMembeddingII.laplacian(M_dev, nS, sig, ctx=ctx, resident=True)
                ctx.sync()
                t2 = time.perf_counter()
                ev, _vecs, einfo = DMembeddingII.eigsh_device(L_dev, nS, 16, ctx=ctx)      # a19 on the device (Lanczos)
                t3 = time.perf_counter()
                M_dev.free()
                L_dev.free()
            details['embedding_front_end'] = dict(
                workload='kNN (k = nS) + graph + Ferguson sweep on the resident D of one PD, then the Laplacian '
                         '(host wall clock around the device chain, allocations included)',
                nS=nS, k=nS, knn_graph_sweep_ms=(t1 - t0) * 1e3, laplacian_ms=(t2 - t1) * 1e3,
                eigsh_device_ms=(t3 - t2) * 1e3, lanczos_steps=einfo['steps'], lanczos_converged=bool(einfo['converged']),
                leading_eigenvalue=float(np.max(ev)), logSumWij_finite=bool(np.isfinite(_ls).all()))
        except Exception as e:                       # the embedding front end is reported beside the headline, never instead of it
            details['embedding_front_end'] = dict(error=repr(e))
    # ---- one psi of the NLSA / psi-analysis stage on a synthetic PD (600 snapshots at 128^2)
    if side and (nS, N) == (NS, NPIX):
        try:
            details['nlsa_one_psi'] = nlsa_block(ctx)
        except Exception as e:
            details['nlsa_one_psi'] = dict(error=repr(e))
    ctx.kernel_time(reset=True)
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers through mem_pd_distance_host, 3 PDs in flight; and the bare concurrent H2D ceiling
    e2e = None
    if not args.no_e2e:
        ceil_gbs, ceil_single = h2d_ceiling(ctx, _lib, lib, local, barrier, sources=h_raw)
        nthreads = max(1, int(os.environ.get('MANIFOLDEM_B200_BENCH_INFLIGHT', '3')))
        ctxs = [ctx] + [_lib.Context(local) for _ in range(nthreads - 1)]
        h_D = [_lib.PinnedArray((nS, nS), np.float32) for _ in range(nthreads)]

        def e2e_step(Pn):
            def worker(w):
                for k in range(w, Pn, nthreads):
                    j = k % POOL
                    io = _lib.PdIO()
                    io.raw, io.flip = h_raw[j].ptr, pds[j]['flip'].ctypes.data
                    io.psi_deg, io.df, io.D = pds[j]['psi_deg'].ctypes.data, pds[j]['df'].ctypes.data, h_D[w].ptr
                    _lib.check(lib.mem_pd_distance_host(ctxs[w].handle, C.byref(prms[j]), C.byref(io)))
            th = [threading.Thread(target=worker, args=(w,)) for w in range(nthreads)]
            [x.start() for x in th]
            [x.join() for x in th]

        P_e = max(nthreads, min(P, 48))
        # static partition of the world x P_e PDs of an e2e step by the measured feed rate of every rank: the box does not
        # serve its GPUs equally when all copy at once (profiles/r02_h2d_concurrent_8gpu.txt: 23 GB/s on four, 35 GB/s on
        # the other four), and the e2e leg is bound by exactly that copy
        from manifoldem_python_b200 import partition
        rates = [ceil_gbs]
        if world > 1:
            tr = torch.zeros(world, dtype=torch.float64, device='cuda:%d' % local)
            tr[rank] = ceil_gbs
            dist.all_reduce(tr, op=dist.ReduceOp.SUM)
            rates = [float(x) for x in tr]
        counts = partition.counts_by_speed(world * P_e, rates)
        P_mine = max(1, counts[rank])
        e2e_step(min(P_e, 2 * nthreads))           # warm-up (plans / workspaces of the extra contexts)
        e_steps = max(1, min(args.steps, 2))
        P_all = sum(max(1, c) for c in counts)
        pd_h2d = nS * NN * 4 + nS * 17

        def all_ranks(x):                            # one float per rank -> list over ranks
            if world == 1:
                return [float(x)]
            tv = torch.zeros(world, dtype=torch.float64, device='cuda:%d' % local)
            tv[rank] = x
            dist.all_reduce(tv, op=dist.ReduceOp.SUM)
            return [float(v) for v in tv]

        # (a) static partition: rank r runs counts[r] PDs per step
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step(P_mine)
        for cx in ctxs:
            cx.sync()
        dt_mine = time.perf_counter() - t0
        sec_static = all_ranks(dt_mine)
        dt = max(sec_static)
        sched = dict(static=dict(value=float(nS) * nS * P_all * e_steps / dt / 1e9, seconds_per_rank=sec_static,
                                 pds_per_rank_per_step=counts))
        done_mine = P_mine * e_steps
        done_all = [max(1, c) * e_steps for c in counts]
        chosen = 'static'
        # (b) one job queue for the box, as the reference's Pool.imap_unordered (GetDistancesS2.py:110-113) and the drop-in
        # driver's job queue hand PDs out: a shared counter in the rendezvous store; a rank takes the next PD when one of its
        # in-flight slots frees up, so every GPU is fed at the rate its host link delivers at that moment
        store = None
        if world > 1 and os.environ.get('MANIFOLDEM_B200_BENCH_E2E_SCHED', 'queue') != 'static':
            try:
                store = dist.distributed_c10d._get_default_store()
                store.add('e2e_queue_probe', 1)
            except Exception as e:                   # no usable store: the static shares stand
                sched['queue'] = dict(error=repr(e))
                store = None
        if world > 1 and min(all_ranks(1.0 if store is not None else 0.0)) < 1.0:
            store = None                             # every rank or none
        if store is not None:
            total = P_all * e_steps
            taken = [0] * nthreads

            def qworker(w):
                while True:
                    k = int(store.add('e2e_queue', 1)) - 1
                    if k >= total:
                        return
                    j = k % POOL
                    io = _lib.PdIO()
                    io.raw, io.flip = h_raw[j].ptr, pds[j]['flip'].ctypes.data
                    io.psi_deg, io.df, io.D = pds[j]['psi_deg'].ctypes.data, pds[j]['df'].ctypes.data, h_D[w].ptr
                    _lib.check(lib.mem_pd_distance_host(ctxs[w].handle, C.byref(prms[j]), C.byref(io)))
                    taken[w] += 1
            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=qworker, args=(w,)) for w in range(nthreads)]
            [x.start() for x in th]
            [x.join() for x in th]
            for cx in ctxs:
                cx.sync()
            dq_mine = time.perf_counter() - t0
            sec_q = all_ranks(dq_mine)
            took = [int(round(v)) for v in all_ranks(float(sum(taken)))]
            sched['queue'] = dict(value=float(nS) * nS * total / max(sec_q) / 1e9, seconds_per_rank=sec_q, pds_per_rank=took)
            if sched['queue']['value'] > sched['static']['value']:
                dt, done_mine, done_all, chosen = max(sec_q), sum(taken), took, 'queue'
        ceil_min = min(rates)
        n_done = sum(done_all)
        h2d_b, d2h_b = int(done_mine * pd_h2d / e_steps), int(done_mine * nS * nS * 4 / e_steps)
        e2e_val = float(nS) * nS * n_done / dt / 1e9
        per_gpu_h2d = done_mine * pd_h2d / dt / 1e9
        agg_h2d = float(n_done) * pd_h2d / dt / 1e9
        e2e = dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=h2d_b, d2h_bytes_per_step=d2h_b,
                   pds_per_step=done_mine // e_steps, pds_per_step_all_ranks=n_done // e_steps,
                   pds_per_rank=[c // e_steps for c in done_all], steps=e_steps, in_flight=nthreads,
                   h2d_gbs_per_gpu=per_gpu_h2d, h2d_gbs_all_ranks=agg_h2d,
                   h2d_ceiling=dict(gbs_per_rank=rates, gbs_all_ranks=float(sum(rates)), gbs_per_gpu_slowest_rank=ceil_min,
                                    gbs_this_rank=ceil_gbs, gbs_this_rank_one_buffer_repeated=ceil_single,
                                    ranks_copying_at_once=world,
                                    how='one 512 MB cudaMemcpyAsync from each of the leg\'s %d pinned source stacks in turn, all ranks '
                                        'between the same barriers (one_buffer_repeated: 8 copies of a single 512 MB buffer)' % len(h_raw)),
                   frac_of_h2d_ceiling=agg_h2d / float(sum(rates)) if sum(rates) > 0 else None,
                   schedule=chosen, schedules=sched,
                   partition=('one PD queue for the box (shared counter in the rendezvous store): a rank takes the next PD when an '
                              'in-flight slot frees up' if chosen == 'queue' else
                              'PDs per rank proportional to the rank\'s measured concurrent H2D rate (static)'),
                   note='wall clock bracketed by stream syncs + barrier, max over ranks; H2D of each raw stack from pinned memory and '
                        'D2H of D inside; h2d / d2h bytes per step are this rank\'s (rank 0)')
        for cx in ctxs[1:]:
            cx.close()

    # ---- e2e through the reference's own API: stack on disk -> GetDistancesS2.op -> per-PD records
    dropin = None
    if side and (nS, N) == (NS, NPIX):
        try:
            dropin = dropin_block(h_raw, pds, nS, N)
        except Exception as e:
            dropin = dict(error=repr(e))
    c5 = None
    if side and (nS, N) == (NS, NPIX):
        for a in d_raw + [d_D]:
            a.free()
        try:
            c5 = config5_block(ctx, lib, _lib, local)
        except Exception as e:
            c5 = dict(error=repr(e))
        details['config_5_one_pd'] = c5

    if rank == 0:
        peaks, peak_src = load_peaks()
        tf32_cublas = measure_cublas_tf32(local)
        # executed TF32 tensor flops of one contraction launch: CTAs x (128x256 tile) x K per CTA x 2 x 3 passes
        flops_launch = float(k_items) * 128 * 256 * 2 * 3 * 32.0 * k_kb
        k_avg_ms = k_ms / max(1, k_n)
        achieved = flops_launch / (k_avg_ms * 1e-3) / 1e12 if k_n else None
        # MEASURED_PEAKS.json has no TF32 entry: TF32 runs at half the bf16 rate, so the denominator is half its measured
        # bf16 burst (the kernel is timed alone per launch); half the sustained figure is reported beside it
        peak = 0.5 * float(peaks.get('bf16_tflops', 1590.0))
        peak_sus = 0.5 * float(peaks.get('bf16_tflops_sustained', 1400.0))
        ncu = contraction_ncu()
        alg = 6.0 * NN * nS * nS                        # 6 N^2 fp32-equivalent flop per ordered pair
        hw_at_clock = 148 * 2048 * 2 * k_mhz * 1e6 / 1e12 if k_mhz else None   # 148 SMs x 2048 TF32 MAC/clk x 2
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        pre_ms = stage['ingest_lowpass'] + stage['align'] + stage['fft_ctf_operands'] + stage['flip_avg']
        ldz = 32 * (2 * 186 + 2048) if N == 256 else None
        stage_alg = dict(ingest_lowpass=8.0 * NN * nS, align=8.0 * NN * nS,
                         fft_ctf_operands=(4.0 * NN + (8.0 * ldz if ldz else 16.0 * NN)) * nS)
        hbm_roof = dict(bound='hbm', peak=hbm_peak, unit='GB/s', peak_source=peak_src,
                        whole_preprocessing=dict(algorithmic_bytes=20.0 * NN * nS, ms=pre_ms,
                                                 achieved=20.0 * NN * nS / (pre_ms * 1e-3) / 1e9,
                                                 frac=20.0 * NN * nS / (pre_ms * 1e-3) / 1e9 / hbm_peak),
                        stages={k: dict(algorithmic_bytes=b, ms=stage[k], achieved=b / (stage[k] * 1e-3) / 1e9,
                                        frac=b / (stage[k] * 1e-3) / 1e9 / hbm_peak) for k, b in stage_alg.items()},
                        note='stage times = CUDA events of the last timed PD; algorithmic bytes: SURVEY §8d (raw fp32 in, '
                             'four fp32 operand planes out) split per stage as read-once / write-once of each stage')
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32 (3xTF32 tensor-core products, fp32 accumulate)', data='synthetic',
                    config=workload_config(nS, N, P),
                    clocks=clocks, gpu_launches=launches, e2e=e2e,
                    roofline=dict(bound='tensor', kernel='k_contract_tc2 (tcgen05 cta_group::2 kind::tf32, 3 passes)',
                                  achieved=achieved, peak=peak, unit='TFLOP/s', frac=(achieved / peak) if achieved else None,
                                  peak_source='0.5 x bf16_tflops (burst) of %s: TF32 runs at half the bf16 rate' % peak_src,
                                  half_bf16_sustained=peak_sus, frac_vs_half_bf16_sustained=(achieved / peak_sus) if achieved else None,
                                  nominal_tf32_dense=1100.0, frac_vs_nominal=(achieved / 1100.0) if achieved else None,
                                  traffic=(ncu or {}).get('dram_bytes_per_launch'), ncu_capture=ncu,
                                  avg_launch_ms=k_avg_ms, launches=k_n,
                                  sm_mhz_in_kernel=k_mhz, sm_mhz_probe='clock64 / globaltimer inside CTA 0 of the last launch',
                                  tensor_pipe_rate_at_kernel_clock=hw_at_clock,
                                  frac_of_tensor_pipe_at_kernel_clock=(achieved / hw_at_clock) if (achieved and hw_at_clock) else None,
                                  cublas_tf32_tflops_measured_in_run=tf32_cublas,
                                  executed_flops_per_launch=flops_launch,
                                  algorithmic_flops_per_launch=alg,
                                  algorithmic_tflops=alg / (k_avg_ms * 1e-3) / 1e12 if k_n else None,
                                  share_of_step=k_ms / ms if ms else None),
                    hbm_roofline=hbm_roof,
                    details=dict(per_pd_ms=ms / args.steps / P, stage_ms_last_pd=stage, numa_cpus=numa, **details),
                    e2e_dropin=dropin)
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            try:
                os.sched_setaffinity(0, range(cores))       # the CPU arm gets every host core back
            except Exception:
                pass
            v, detail, _ = cpu_reference_sample(nS, N, 4 * cores, cores, contraction_rows=min(nS, 500))
            line['cpu_baseline'] = dict(value=v, unit=UNIT, cores=cores, kind='port', detail=detail,
                                        unmodified_reference_vs_port=reference_vs_port(),
                                        other_configs=None if args.no_details else cpu_split_other_configs(cores),
                                        sample='%d of %d particles through the per-image stages + a %dx%d block of the '
                                               'fp64 GEMMs + the pickle of those particles, extrapolated to one PD; value = the '
                                               'faster of the serial mode (all BLAS threads) and the Pool / MPI-schedule '
                                               'emulation (one PD per worker process)'
                                               % (detail['images'], nS, detail['gemm_rows'], detail['gemm_rows']))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dropin_block(h_raw, pds, nS, N):
    """Disk -> GetDistancesS2.op (the reference's stage driver signature) -> per-PD records, 8 PDs of the bench shape:
    a SPIDER stack + tessellation pickle on local disk, the drop-in modules, record layout 'sidecar' (float32 arrays,
    virtual CTF) and the reference's single float64 pickle."""
    import shutil
    import tempfile
    from manifoldem_python_b200 import GetDistancesS2, myio, p, synthetic
    n_pd = 8
    os.environ['MANIFOLDEM_B200_GPUS'] = '1'                  # this block measures ONE GPU (rank 0's)
    os.environ['MANIFOLDEM_B200_DEVICE'] = os.environ.get('LOCAL_RANK', '0')
    base = tempfile.mkdtemp(prefix='mem_b200_bench_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    out = dict(workload='%d PDs of %d x %d^2: stack on disk -> GetDistancesS2.op -> records' % (n_pd, nS, N), where=base)
    try:
        stack = os.path.join(base, 'stack.dat')
        with open(stack, 'wb') as f:
            for k in range(n_pd):
                f.write(memoryview(h_raw[k % len(h_raw)].array).cast('B'))
        n_half = n_pd * nS
        rng = np.random.default_rng(11)
        q = np.concatenate([synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS), 1.1 + 0.03 * rng.standard_normal(nS),
                                                    rng.uniform(0, 2 * np.pi, nS)) for _ in range(n_pd)], axis=1)
        q = np.concatenate([q, q], axis=1)                      # augmented set (conjugates appended); only first halves used
        df = np.tile(rng.uniform(10000.0, 30000.0, n_half), 2)
        CG = [np.arange(k * nS, (k + 1) * nS) for k in range(n_pd)]
        for layout in ('sidecar', 'sidecar_recipe', 'pickle'):
            p.init()
            p.nPix, p.pix_size, p.Cs, p.EkV, p.AmpContrast = N, EM['pix_size'], EM['Cs'], EM['EkV'], EM['AmpContrast']
            p.mask_vol_file, p.relion_data, p.num_part, p.ncpu = '', False, n_half, 1
            p.img_stack_file = stack
            p.dist_dir = os.path.join(base, 'dist_' + layout) + os.sep
            p.dist_prog = os.path.join(p.dist_dir, 'progress') + os.sep
            os.makedirs(p.dist_prog)
            p.dist_file = os.path.join(p.dist_dir, 'IMGs_')
            p.tess_file = os.path.join(base, 'tess_' + layout)
            p.numberofJobs = n_pd
            p.record_layout = layout.split('_')[0]
            p.record_virtual_images = (layout == 'sidecar_recipe')    # imgAll / imgAllFlip kept as a recipe, rebuilt on read
            myio.fout1(p.tess_file, ['CG', 'q', 'df', 'sh'], [CG, q, df, (np.zeros(n_half), np.zeros(n_half))], layout='pickle')
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):        # the stage driver prints like the reference's; stdout = ONE JSON line
                GetDistancesS2.op()
            dt = time.perf_counter() - t0
            done = len(os.listdir(p.dist_prog))
            size = sum(os.path.getsize(os.path.join(p.dist_dir, f)) for f in os.listdir(p.dist_dir)
                       if os.path.isfile(os.path.join(p.dist_dir, f)))
            out[layout] = dict(s_per_pd=dt / n_pd, pds_done=done, record_gb_per_pd=size / n_pd / 1e9,
                               gpairs_s=n_pd * float(nS) * nS / dt / 1e9)
            if layout == 'sidecar_recipe':                 # what a consumer pays to get the image arrays back
                t1 = time.perf_counter()
                rec0 = myio.fin1(p.dist_file + 'prD_0')
                img0 = rec0['imgAll']
                out[layout]['read_back_imgAll_and_imgAllFlip_s'] = time.perf_counter() - t1
                out[layout]['note'] = ('imgAll / imgAllFlip are kept as a recipe (stack path + the ind / q / df already in the record) and '
                                       'rebuilt by the same kernels when read: bit-identical values, shape %s' % (img0.shape,))
                del rec0, img0
            shutil.rmtree(p.dist_dir, ignore_errors=True)
            for attr in ('record_layout', 'record_virtual_images'):
                if hasattr(p, attr):
                    delattr(p, attr)
    finally:
        shutil.rmtree(base, ignore_errors=True)
    return out


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
