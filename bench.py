#!/usr/bin/env python
"""bench.py — CTF pairwise distances/s of the per-PD distance stage on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W]                 one process per GPU (torchrun for N>1)
  python bench.py --impl reference [--gpus N --steps K --warmup W]   the reference's CPU algorithm (oracle port)

Workload (weak scaling): every rank processes PDS_PER_GPU synthetic projection directions of
BASELINE config 4's shape (2,000 particles x 256^2) per step; at 8 GPUs one step is the whole
1,000-PD S2 run.  A "pair" is one entry of the nS x nS matrix D the reference materialises.
`value`  : whole-job G pairs/s with the raw particle stacks already resident in HBM.
`e2e`    : the same through the host-buffer C-ABI call (pinned host stacks -> D on the host),
           H2D and D2H inside the timed region, three PDs in flight per GPU.
`roofline`: executed TF32 tensor flops of the tcgen05 contraction per launch / its CUDA-event time.
"""
import argparse
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
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- helpers
def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback'


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
    """cuBLAS TF32 GEMM (8192^3, best of 5 after warm-up) measured in this run, next to the roofline's official
    denominator (MEASURED_PEAKS.json has no TF32 entry): context for the reader, not the `peak` field."""
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
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned host
    stacks of the e2e leg live on that socket (8 ranks otherwise share one socket's memory controllers)."""
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
    """Per-image part of the reference algorithm (ingest .. FFT) for a slice of particles."""
    nS_s, N, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from manifoldem_python_b200 import synthetic
    from oracle import pd_distance as opd
    rng = np.random.default_rng(seed)
    stack = rng.standard_normal(nS_s * N * N).astype(np.float32)
    q = synthetic.euler_to_quat(0.7 + 0.03 * rng.standard_normal(nS_s), 1.1 + 0.03 * rng.standard_normal(nS_s),
                                rng.uniform(0, 2 * np.pi, nS_s))
    df = rng.uniform(10000.0, 30000.0, nS_s)
    t0 = time.perf_counter()
    opd.pd_distance(np.arange(nS_s), q, df, stack, 2 * nS_s, N, EM['pix_size'], EM['Cs'], EM['EkV'], EM['AmpContrast'],
                    avg_only=True, rotate_impl='tile', keep=('imgAvg',))
    return time.perf_counter() - t0


def _ref_gemm_worker(args):
    """The two GEMMs of :391-397 on a rows x rows block with ONE BLAS thread (a Pool / MPI worker's share)."""
    rows, N, seed = args
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    rng = np.random.default_rng(seed)
    K = N * N
    CTF = rng.standard_normal((rows, K))
    fy = rng.standard_normal((rows, K)) + 1j * rng.standard_normal((rows, K))
    t0 = time.perf_counter()
    CTFfy = CTF.conj() * fy
    D = np.dot(np.abs(CTF) ** 2, (np.abs(fy) ** 2).T)
    D = D + D.T - 2 * np.real(np.dot(CTFfy, CTFfy.conj().T))
    return time.perf_counter() - t0


def cpu_reference_sample(nS, N, n_img, cores, contraction_rows):
    """Bounded sample of the reference's CPU algorithm on one PD of shape (nS, N):
    (1) the per-image stages (ingest, low-pass, 2x rotatefill on the 3x3 tile, CTF, 3 FFTs) on n_img particles,
        spread over `cores` worker processes — these stages are linear in nS;
    (2) the two GEMMs of :391-397 (float64 / complex128, all BLAS threads) on `contraction_rows` x nS pairs.
    Per-PD time = nS/n_img * t_images + nS/contraction_rows * t_gemm: the reference's serial mode (p.ncpu = 1) with
    the per-image loop ideally parallel.  Also timed: the Pool mode (GetDistancesS2.py:110-113) and its MPI twin's
    static round-robin (GetDistancesS2_mpi.py:14-15) — one whole PD per worker process, one BLAS thread each, `cores`
    PDs at once; for equal PDs the two schedules coincide.  Returns (pairs/s of the faster mode, detail dict)."""
    import multiprocessing as mp
    per = max(1, n_img // cores)
    jobs = [(per, N, 100 + i) for i in range(max(1, n_img // per))]
    rows1 = min(contraction_rows, 200)
    with mp.get_context('spawn').Pool(min(cores, len(jobs))) as pool:
        t_img = max(pool.map(_ref_images_worker, jobs))     # slowest worker's compute time (spawn/import excluded)
        t_gemm1 = max(pool.map(_ref_gemm_worker, [(rows1, N, 7 + i) for i in range(min(cores, len(jobs)))]))
    n_done = per * len(jobs)
    rng = np.random.default_rng(0)
    rows = contraction_rows
    K = N * N
    CTF = rng.standard_normal((rows, K))
    fy = rng.standard_normal((rows, K)) + 1j * rng.standard_normal((rows, K))
    try:                                  # torchrun exports OMP_NUM_THREADS=1: give the GEMMs every core back
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=cores)
    except Exception:
        limiter = None
    t0 = time.perf_counter()
    CTFfy = CTF.conj() * fy
    D = np.dot(np.abs(CTF) ** 2, (np.abs(fy) ** 2).T)
    D = D + D.T - 2 * np.real(np.dot(CTFfy, CTFfy.conj().T))
    t_gemm = time.perf_counter() - t0
    if limiter is not None:
        limiter.restore_original_limits()
    # rows x rows block measured; the full matrix has (nS/rows)^2 such blocks
    t_pd = (nS / n_done) * t_img + (nS / rows) ** 2 * t_gemm
    # Pool / MPI-schedule emulation: a worker runs its PD alone (per images in t_img, one BLAS thread), `workers` at once
    workers = min(cores, len(jobs))
    t_pd_worker = (nS / per) * t_img + (nS / rows1) ** 2 * t_gemm1
    v_serial, v_pool = nS * nS / t_pd / 1e9, workers * nS * nS / t_pd_worker / 1e9
    detail = dict(images=n_done, t_images_s=round(t_img, 3), gemm_rows=rows, t_gemm_s=round(t_gemm, 3),
                  per_pd_s=round(t_pd, 2),
                  stage_split_per_pd_s=dict(per_image_stages=round((nS / n_done) * t_img, 2),
                                            gemm=round((nS / rows) ** 2 * t_gemm, 2)),
                  serial_mode_gpairs_s=v_serial,
                  pool_and_mpi_schedule_emulation=dict(gpairs_s=v_pool, workers=workers, blas_threads_per_worker=1,
                                                       per_pd_per_worker_s=round(t_pd_worker, 1), gemm_rows=rows1,
                                                       t_gemm_s=round(t_gemm1, 3)))
    return max(v_serial, v_pool), detail


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nS, N = args.nS, args.N
    vals = []
    detail = None
    n_img = 4 * cores
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(nS, N, n_img, cores, contraction_rows=min(nS, 500))
        if i >= args.warmup:
            vals.append((v, nS * nS / (v * 1e9)))     # seconds per PD at the reported rate
    v = float(np.mean([a for a, _ in vals]))
    sample = ('%d of %d particles through the per-image stages on %d processes + a %dx%d block of the fp64 '
              'dgemm/zgemm; per-PD time extrapolated linearly in images and quadratically in the block; value = the '
              'faster of the serial mode (all BLAS threads) and the Pool / MPI-schedule emulation (one PD per worker)'
              % (detail['images'], nS, cores, detail['gemm_rows'], detail['gemm_rows']))
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean([b for _, b in vals])) * 1e3 * args.pds * max(1, args.gpus), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=dict(workload='BASELINE config 4 shape: PDs of %d particles x %d^2, %d PDs per GPU per step '
                                     '(1000 PDs at 8 GPUs)' % (nS, N, args.pds), pds_per_gpu=args.pds, nS=nS, N=N),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind='port', sample=sample, detail=detail),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


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

    # ---- the same PDs with every array of the reference's per-PD record produced on the device as well (imgAll,
    # imgAllFlip, float64 CTF, Wiener / flip averages, intensity): reported beside the headline, which asks for D only
    full_ms = None
    if rank == 0 and not args.no_e2e:
        outs = [_lib.DeviceArray(ctx, (nS, NN), np.float32), _lib.DeviceArray(ctx, (nS, NN), np.float32),
                _lib.DeviceArray(ctx, (nS, NN), np.float64), _lib.DeviceArray(ctx, (NN,), np.float32),
                _lib.DeviceArray(ctx, (NN,), np.float32), _lib.DeviceArray(ctx, (NN,), np.float32)]
        fio = []
        for j in range(POOL):
            io = _lib.PdIO()
            io.raw, io.flip, io.psi_deg, io.df, io.D = d_raw[j].ptr, d_flip[j].ptr, d_psi[j].ptr, d_df[j].ptr, d_D.ptr
            (io.imgAll, io.imgAllFlip, io.CTF, io.imgAvg, io.imgAvgFlip, io.imgAllIntensity) = [o.ptr for o in outs]
            fio.append(io)
        n_full = 8
        for k in range(2 + n_full):
            if k == 2:
                ctx.sync()
                ctx.timer_start()
            _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prms[k % POOL]), C.byref(fio[k % POOL]), None))
        full_ms = ctx.timer_stop() / n_full
        for o in outs:
            o.free()
    # ---- BASELINE config 2 (one synthetic PD of 1,000 particles at 128^2, full D), device-resident, for reference
    c2 = None
    if rank == 0 and not args.no_e2e and (nS, N) == (NS, NPIX):
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
        reps2 = 40
        for k in range(4 + reps2):
            if k == 4:
                ctx.sync()
                ctx.timer_start()
            _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm2[k % 2]), C.byref(io2[k % 2]), None))
        ms2 = ctx.timer_stop() / reps2
        c2 = dict(workload='BASELINE config 2: one PD of 1,000 particles x 128^2, full D', per_pd_ms=ms2,
                  gpairs_s=n2 * n2 / (ms2 * 1e-3) / 1e9)
        for a in raws2 + [x for t3 in aux2 for x in t3] + [D2]:
            a.free()
    # ---- BASELINE config 3 (one PD of 5,000 particles at 256^2, k = 100 neighbour lists only): the lists are selected
    # from the contraction's partial tiles, D is never assembled.  Particles: rows of three of the pool stacks.
    c3 = None
    if rank == 0 and not args.no_e2e and (nS, N) == (NS, NPIX) and POOL * nS >= 5000:
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
        reps3 = 5
        for k in range(2 + reps3):
            if k == 2:
                ctx.sync()
                ctx.timer_start()
            _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm3), C.byref(io3), None))
        ms3 = ctx.timer_stop() / reps3
        nb = idx3.download()
        assert np.array_equal(nb[:, 0], np.arange(n3)) and nb.min() >= 0 and nb.max() < n3
        c3 = dict(workload='BASELINE config 3: one PD of 5,000 particles x 256^2, k = 100 neighbour lists only (no D)',
                  per_pd_ms=ms3, gpairs_s=n3 * n3 / (ms3 * 1e-3) / 1e9, knn_k=k3, stage_ms=ctx.timings())
        for a in [raw3, idx3, val3] + aux3:
            a.free()
        ctx.kernel_time(reset=True)
    # ---- rows a15-a18 on the D the last PD left on the device (k = nS, the way manifoldTrimmingAuto calls
    # DMembeddingII.op): kNN, graph, Ferguson sweep, Gaussian-kernel Laplacian; wall clock around the device chain
    dm = None
    if rank == 0 and not args.no_e2e:
        try:
            from manifoldem_python_b200 import DMembeddingII
            ctx.sync()
            sig = 3.0 * float(np.sqrt(np.median(d_D.download()[0, 1:])))
            for rep in range(2):                     # second pass = warm workspaces
                t0 = time.perf_counter()
                M_dev, _le, _ls, _i, _v = DMembeddingII.graph_and_sweep(d_D, nS, ctx=ctx, want_lists=False)
                t1 = time.perf_counter()
                L_dev = DMembeddingII.laplacian(M_dev, nS, sig, ctx=ctx, resident=True)
                ctx.sync()
                t2 = time.perf_counter()
                M_dev.free()
                L_dev.free()
            dm = dict(workload='kNN (k = nS) + graph + Ferguson sweep on the resident D of one PD, then the Laplacian '
                               '(host wall clock around the device chain, allocations included)',
                      nS=nS, k=nS, knn_graph_sweep_ms=(t1 - t0) * 1e3, laplacian_ms=(t2 - t1) * 1e3,
                      logSumWij_finite=bool(np.isfinite(_ls).all()))
        except Exception as e:                       # the embedding front end is reported beside the headline, never instead of it
            dm = dict(error=repr(e))
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers through mem_pd_distance_host, 3 PDs in flight
    e2e = None
    if not args.no_e2e:
        nthreads = 3
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
        e2e_step(min(P_e, 2 * nthreads))           # warm-up (plans / workspaces of the extra contexts)
        barrier()
        t0 = time.perf_counter()
        e_steps = max(1, min(args.steps, 2))
        for _ in range(e_steps):
            e2e_step(P_e)
        for cx in ctxs:
            cx.sync()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device='cuda:%d' % local)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te[0])
        e2e = dict(value=float(nS) * nS * P_e * world * e_steps / dt / 1e9, unit=UNIT,
                   h2d_bytes_per_step=int(P_e * (nS * NN * 4 + nS * 17)), d2h_bytes_per_step=int(P_e * nS * nS * 4),
                   pds_per_step=P_e, steps=e_steps, in_flight=nthreads,
                   note='wall clock bracketed by stream syncs + barrier; H2D of each raw stack from pinned memory and D2H of D inside')

    if rank == 0:
        peaks, peak_src = load_peaks()
        tf32_cublas = measure_cublas_tf32(local)
        # executed TF32 tensor flops of one contraction launch: items x (128x256 tile) x K x 2 x 3 passes
        flops_launch = float(k_items) * 128 * 256 * 2 * 3 * 32.0 * k_kb      # CTAs x tile x K per CTA x 3 passes
        k_avg_ms = k_ms / max(1, k_n)
        achieved = flops_launch / (k_avg_ms * 1e-3) / 1e12 if k_n else None
        # MEASURED_PEAKS.json has no TF32 entry.  Half its measured bf16 burst (cuBLAS; TF32 runs at half the bf16 rate)
        # was the planning value, but this kernel sustains more than that, so it is not a ceiling: the denominator is
        # the TF32 dense figure of B200_PROFILING.md's table, 1.1 PFLOP/s; the half-bf16 ratio is reported beside it.
        half_bf16 = 0.5 * float(peaks.get('bf16_tflops', 1590.0))
        peak = 1100.0
        # dram__bytes_read+write of one launch from the committed ncu capture (profiles/r01_top_kernels_ncu_full.txt)
        traffic = 1.6318e9 if (nS, N) == (2000, 256) else None
        alg = 6.0 * NN * nS * nS                        # 6 N^2 fp32-equivalent flop per ordered pair
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32 (3xTF32 tensor-core products, fp32 accumulate)', data='synthetic',
                    config=dict(workload='BASELINE config 4 shape: PDs of %d particles x %d^2, %d PDs per GPU per step '
                                         '(1000 PDs at 8 GPUs)' % (nS, N, P), pds_per_gpu=P, nS=nS, N=N,
                                l2='inputs larger than L2: %d distinct 524 MB stacks cycled' % POOL,
                                per_pd_ms=ms / args.steps / P, stage_ms_last_pd=stage,
                                all_record_fields_per_pd_ms=full_ms, single_pd_config_2=c2, single_pd_config_3=c3,
                                embedding_front_end=dm),
                    clocks=clocks, gpu_launches=launches, e2e=e2e,
                    roofline=dict(bound='tensor', kernel='k_contract_tc2 (tcgen05 cta_group::2 kind::tf32, 3 passes)',
                                  achieved=achieved, peak=peak, unit='TFLOP/s', frac=(achieved / peak) if achieved else None,
                                  traffic=traffic, avg_launch_ms=k_avg_ms, launches=k_n,
                                  cublas_tf32_tflops_measured_in_run=tf32_cublas,
                                  # hardware ceiling at the clock sampled during the run: 148 SMs x 2048 TF32 MAC/clk x 2
                                  frac_of_hw_rate_at_sampled_clock=(achieved / (148 * 2048 * 2 * clocks['sm_mhz'] * 1e6 / 1e12))
                                  if (achieved and clocks and clocks.get('sm_mhz')) else None,
                                  ncu_tensor_pipe_active_pct=90.8 if traffic else None,   # sm__pipe_tensor_cycles_active, % of elapsed, same capture
                                  executed_flops_per_launch=flops_launch,
                                  algorithmic_tflops=alg / (k_avg_ms * 1e-3) / 1e12 if k_n else None,
                                  peak_source='TF32 dense 1.1 PFLOP/s (B200_PROFILING.md table); MEASURED_PEAKS.json (%s) has no TF32 '
                                              'entry and half its bf16 burst is below what this kernel sustains' % peak_src,
                                  half_measured_bf16_tflops=half_bf16,
                                  frac_vs_half_measured_bf16=(achieved / half_bf16) if achieved else None,
                                  share_of_step=k_ms / ms if ms else None))
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            try:
                os.sched_setaffinity(0, range(cores))       # the CPU arm gets every host core back
            except Exception:
                pass
            v, detail = cpu_reference_sample(nS, N, 4 * cores, cores, contraction_rows=min(nS, 500))
            line['cpu_baseline'] = dict(value=v, unit=UNIT, cores=cores, kind='port', detail=detail,
                                        sample='%d of %d particles through the per-image stages + a %dx%d block of the '
                                               'fp64 GEMMs, extrapolated to one PD; value = the faster of the serial mode (all BLAS '
                                               'threads) and the Pool / MPI-schedule emulation (one PD per worker process)'
                                               % (detail['images'], nS, detail['gemm_rows'], detail['gemm_rows']))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
