"""Host side of one PD of the distance stage: scalar bookkeeping on the CPU exactly as the
reference does it (a1, a6), image gather (a2), then ONE call through the C ABI for everything else.

Reference: modules/getDistanceCTF_local_Conj9combinedS2.py:216-420.
"""
import ctypes as C
import math
import os

import numpy as np

from . import _lib
from . import q2Spider

VERSION = 'getDistanceCTF_local9, V 1.0'       # same tag the reference stores (:41)
FILTERS = {'Butter': 0, 'Gauss': 1}


# ----------------------------------------------------------------------------- a6 (host scalars)
def calc_avg_pd(q):
    """(:169-185) per-image projection directions, (3,nS)."""
    if q.shape[0] <= 3:
        raise AssertionError('quaternion has wrong dimensions')
    return 2 * np.vstack((q[1] * q[3] - q[0] * q[2], q[0] * q[1] + q[2] * q[3], q[0] ** 2 + q[3] ** 2 - 0.5))


def psi_ang(PD):
    """(:206-213) in-plane angle of the PD itself in degrees, via q2Spider."""
    Qr = np.array([1 + PD[2], PD[1], -PD[0], 0.0])
    Qr = Qr / np.sqrt(np.sum(Qr ** 2))
    _, _, psi = q2Spider.op(Qr)
    return np.mod(psi, 2 * np.pi) * (180 / np.pi)


def host_angles(q):
    """PDs, PD, psi_p [deg], Psi [rad], s (Nom), c (Dnom) — (:297-323)."""
    PDs = calc_avg_pd(q)
    PD = np.sum(PDs, 1)
    PD = PD / np.linalg.norm(PD)
    psi_p = psi_ang(PD)
    s = -(1 + PD[2]) * q[3] - PD[0] * q[1] - PD[1] * q[2]
    c = (1 + PD[2]) * q[0] + PD[1] * q[1] - PD[0] * q[2]
    with np.errstate(divide='ignore', invalid='ignore'):
        Psi = 2 * np.arctan(s / c)
    Psi = np.where(np.isnan(Psi), 0.0, Psi)
    return PDs, PD, psi_p, Psi, s, c


# ----------------------------------------------------------------------------- a2 (gather)
def gather(stack, ind, nStot, N, out=None):
    """Member images in PD order as stored on disk + conjugate flags (:246-262).
    `stack`: flat float32 array / np.memmap (SPIDER raw) or (n,N,N) array (MRC stack data)."""
    ind = np.asarray(ind)
    nS = ind.shape[0]
    half = nStot / 2
    conj = ~(ind < half)
    base = np.where(conj, ind - half, ind).astype(np.int64)
    if out is None:
        out = np.empty((nS, N * N), dtype=np.float32)
    flat = stack.reshape(-1, N * N) if stack.ndim != 2 else stack
    if flat.dtype == np.float32 and flat.flags.c_contiguous and out.flags.c_contiguous and nS >= 64:
        # a few host threads memcpy the rows straight from the mapped file into the (pinned) destination; the
        # reference re-opens the memmap for every particle (:254-256)
        if base.min() < 0 or base.max() >= flat.shape[0]:
            raise IndexError('particle index outside the stack')
        rows = np.ascontiguousarray(base, dtype=np.int64)
        _lib.check(_lib.load().mem_gather_rows_host(out.ctypes.data, flat.ctypes.data, rows.ctypes.data, nS,
                                                    N * N * 4, int(os.environ.get('MANIFOLDEM_B200_GATHER_THREADS', '4'))))
    else:
        order = np.argsort(base, kind='stable')            # read the file front to back
        out[order] = flat[base[order]]
    return out, conj.astype(np.uint8), base


def open_stack(imgFileName, N, relion):
    """SPIDER: raw float32, image i at byte 4*N*N*i (:254-256).  RELION .mrcs: MRC2014 header
    (1024 bytes + NSYMBT extended header), mode 2 float32, (n,N,N) (:259-262)."""
    if not relion:
        return np.memmap(imgFileName, dtype='float32', mode='r')
    hdr = np.fromfile(imgFileName, dtype='<i4', count=256)
    nx, ny, nz, mode, nsymbt = int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[23])
    if mode != 2 or nx != N or ny != N:
        raise ValueError('unsupported MRC stack: mode %d, %dx%d (need float32 %dx%d)' % (mode, nx, ny, N, N))
    return np.memmap(imgFileName, dtype='<f4', mode='r', offset=1024 + nsymbt, shape=(nz, ny, nx))


class HostArena:
    """Pinned host buffers reused from one PD to the next by the thread that owns them: a fresh 0.5 GB NumPy array per
    PD costs more in page faults than the copy it receives, and pinned memory lets the H2D / D2H copies of the
    host-buffer entry point run at PCIe speed.  Arrays handed out stay valid until the same name is asked for again."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        buf = self._bufs.get(name)
        if buf is None or buf.nbytes < n * dtype.itemsize:
            if buf is not None and hasattr(buf, 'free'):
                buf.free()
            try:
                buf = _lib.PinnedArray((max(n * dtype.itemsize, 1),), np.uint8)
            except RuntimeError:                       # no pinned memory left: pageable, still reused
                buf = np.empty(max(n * dtype.itemsize, 1), dtype=np.uint8)
            self._bufs[name] = buf
        flat = buf.array if hasattr(buf, 'array') else buf
        return flat[:n * dtype.itemsize].view(dtype).reshape(shape)

    def close(self):
        for b in self._bufs.values():
            if hasattr(b, 'free'):
                b.free()
        self._bufs = {}


# ----------------------------------------------------------------------------- the C-ABI call
def run_pd(ind, q, df, stack, nStot, N, pix_size, Cs, EkV, AmpContrast, gaussEnv=np.inf, filterPar=None,
           msk2=None, relion=False, sh=None, avg_only=False, ctx=None, fields=('D', 'imgAll', 'imgAllFlip', 'CTF'),
           contraction=0, k_chunk_blocks=0, split_k=0, float64=True, angles=None, knn_k=0, arena=None,
           intensity=None):
    """Returns the dict the reference pickles (same keys / shapes; float64 unless float64=False).
    `fields` selects which of the heavy per-image outputs are materialised.  knn_k > 0 adds `knn_idx` (nS,k) int32
    and `knn_val` (nS,k) float64 — the lists DMembeddingII.initialize (:43-57) would take from D — selected on the
    device straight behind the contraction; without 'D' in `fields` the nS x nS matrix is then never assembled.
    `intensity`: produce imgAllIntensity (:400); default = whenever any per-image record array is asked for (the
    kernel does not need imgAllFlip exported to compute it).
    `arena` (a HostArena owned by the calling thread): the gathered stack and the outputs live in its reused pinned
    buffers — the arrays of the result are then only valid until the next call with the same arena."""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    if filterPar is None:
        filterPar = dict(type='Butter', Qc=0.5, N=8)
    if filterPar['type'] not in FILTERS:
        raise ValueError('%s filter is unsupported' % filterPar['type'])
    ind = np.asarray(ind)
    q = np.asarray(q, dtype=np.float64)
    df = np.ascontiguousarray(df, dtype=np.float64)
    nS = ind.shape[0]
    PDs, PD, psi_p, Psi, s, c = angles if angles is not None else host_angles(q)
    raw, flip, base = gather(stack, ind, nStot, N, out=arena.get('raw', (nS, N * N), np.float32) if arena else None)
    shift = None
    if relion:                                         # (:263) shi = (sh[1][idx] - 0.5, sh[0][idx] - 0.5)
        shift = np.ascontiguousarray(np.stack((np.asarray(sh[1])[base] - 0.5, np.asarray(sh[0])[base] - 0.5), axis=1),
                                     dtype=np.float64)
    psi_deg = np.ascontiguousarray(-(180 / math.pi) * Psi, dtype=np.float64)

    prm = _lib.PdParams(nS=nS, N=N, transposed=0 if relion else 1, relion_shift=1 if relion else 0,
                        filter_type=FILTERS[filterPar['type']], filter_order=int(filterPar['N']),
                        filter_Qc=float(filterPar['Qc']), pix_size=float(pix_size), Cs=float(Cs), EkV=float(EkV),
                        gaussEnv=float(gaussEnv), AmpContrast=float(AmpContrast), psi_p_deg=float(psi_p),
                        avg_only=1 if avg_only else 0, contraction=int(contraction),
                        k_chunk_blocks=int(k_chunk_blocks), split_k=int(split_k), knn_k=int(knn_k))
    outs = {}

    def want(name, shape, dtype):
        a = arena.get(name, shape, dtype) if arena else np.empty(shape, dtype=dtype)
        outs[name] = a
        return a.ctypes.data

    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df = raw.ctypes.data, flip.ctypes.data, psi_deg.ctypes.data, df.ctypes.data
    if shift is not None:
        io.shift = shift.ctypes.data
    m2 = None
    if msk2 is not None and not np.isscalar(msk2):
        m2 = np.ascontiguousarray(np.asarray(msk2) != 0, dtype=np.uint8)
        io.msk2 = m2.ctypes.data
    if 'D' in fields and not avg_only:
        io.D = want('D', (nS, nS), np.float32)
    if knn_k and not avg_only:
        io.knn_idx = want('knn_idx', (nS, int(knn_k)), np.int32)
        io.knn_val = want('knn_val', (nS, int(knn_k)), np.float64)
    if 'imgAll' in fields:
        io.imgAll = want('imgAll', (nS, N, N), np.float32)
    if 'imgAllFlip' in fields:
        io.imgAllFlip = want('imgAllFlip', (nS, N, N), np.float32)
    if 'CTF' in fields:
        io.CTF = want('CTF', (nS, N * N), np.float64)
    io.imgAvg = want('imgAvg', (N, N), np.float32)
    io.imgAvgFlip = want('imgAvgFlip', (N, N), np.float32)
    if intensity is None:
        intensity = any(f in fields for f in ('imgAll', 'imgAllFlip', 'CTF'))
    if intensity:
        io.imgAllIntensity = want('imgAllIntensity', (N, N), np.float32)
    _lib.check(lib.mem_pd_distance_host(ctx.handle, C.byref(prm), C.byref(io)))

    cast = (lambda a: a.astype(np.float64)) if float64 else (lambda a: a)
    res = dict(D=cast(outs['D']) if 'D' in outs else np.zeros((nS, nS)), ind=ind, q=q, df=df,
               CTF=outs.get('CTF'), imgAll=cast(outs['imgAll']) if 'imgAll' in outs else None,
               msk2=1 if m2 is None else (m2 != 0), PD=PD, PDs=PDs, Psis=Psi.reshape(nS, 1),
               imgAvg=cast(outs['imgAvg']), imgAvgFlip=cast(outs['imgAvgFlip']),
               imgAllFlip=cast(outs['imgAllFlip']) if 'imgAllFlip' in outs else None,
               imgLabels=np.where(flip != 0, -1, 1).astype(int), Dnom=c.reshape(nS, 1), Nom=s.reshape(nS, 1),
               imgAllIntensity=cast(outs['imgAllIntensity']) if 'imgAllIntensity' in outs else None,
               version=VERSION)
    res['_psi_p'] = psi_p
    if 'knn_idx' in outs:
        res['knn_idx'], res['knn_val'] = outs['knn_idx'], outs['knn_val']
    return res


def run_pd_resident(ind, q, df, stack, nStot, N, pix_size, Cs, EkV, AmpContrast, gaussEnv=np.inf, filterPar=None,
                    ctx=None, angles=None, knn_k=0, keep_D=True):
    """Distance stage with the result kept ON THE DEVICE: returns a `_lib.DeviceArray` (nS,nS) float32 holding D,
    for consumers that run there too (DMembeddingII.graph_and_sweep accepts it) — the N x N matrix never
    round-trips through the host (BASELINE config 3: kNN / kernel construction only).
    knn_k > 0: also (or, with keep_D=False, only) the kNN lists as device arrays — returns (D | None, idx, val);
    with keep_D=False the lists come from the contraction's partial tiles and D is never assembled."""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    if filterPar is None:
        filterPar = dict(type='Butter', Qc=0.5, N=8)
    ind = np.asarray(ind)
    q = np.asarray(q, dtype=np.float64)
    nS = ind.shape[0]
    PDs, PD, psi_p, Psi, s, c = angles if angles is not None else host_angles(q)
    raw, flip, base = gather(stack, ind, nStot, N)
    bufs = [_lib.DeviceArray(ctx, raw.shape, np.float32, raw), _lib.DeviceArray(ctx, (nS,), np.uint8, flip),
            _lib.DeviceArray(ctx, (nS,), np.float64, -(180 / math.pi) * Psi),
            _lib.DeviceArray(ctx, (nS,), np.float64, np.ascontiguousarray(df, dtype=np.float64))]
    D = _lib.DeviceArray(ctx, (nS, nS), np.float32) if (keep_D or not knn_k) else None
    prm = _lib.PdParams(nS=nS, N=N, transposed=1, relion_shift=0, filter_type=FILTERS[filterPar['type']],
                        filter_order=int(filterPar['N']), filter_Qc=float(filterPar['Qc']), pix_size=float(pix_size),
                        Cs=float(Cs), EkV=float(EkV), gaussEnv=float(gaussEnv), AmpContrast=float(AmpContrast),
                        psi_p_deg=float(psi_p), knn_k=int(knn_k))
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df = bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr
    if D is not None:
        io.D = D.ptr
    idx = val = None
    if knn_k:
        idx = _lib.DeviceArray(ctx, (nS, int(knn_k)), np.int32)
        val = _lib.DeviceArray(ctx, (nS, int(knn_k)), np.float64)
        io.knn_idx, io.knn_val = idx.ptr, val.ptr
    _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
    ctx.sync()
    for b in bufs:
        b.free()
    return (D, idx, val) if knn_k else D


def run_pd_batch(jobs, stack, nStot, N, pix_size, Cs, EkV, AmpContrast, gaussEnv=np.inf, filterPar=None, ctx=None,
                 want_imgAll=False):
    """A GROUP of PDs in one device call (`mem_pd_distance_batch_device`): jobs = [(ind, q, df), ...] — the job tuples
    GetDistancesS2.divide builds (:37-47), one per PD.  The members of all PDs are gathered into one stack, the per-image
    stages run once over it, one grouped tcgen05 launch contracts every PD.  SPIDER stacks, msk2 = 1.
    Returns a list of dicts (D (nS,nS) float32, PD, PDs, Psis, _psi_p[, imgAll]) in job order."""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    if filterPar is None:
        filterPar = dict(type='Butter', Qc=0.5, N=8)
    raws, flips, psis, dfs, psi_ps, host = [], [], [], [], [], []
    for ind, q, df in jobs:
        ind = np.asarray(ind)
        PDs, PD, psi_p, Psi, s, c = host_angles(np.asarray(q, dtype=np.float64))
        raw, flip, _base = gather(stack, ind, nStot, N)
        raws.append(raw)
        flips.append(flip)
        psis.append(-(180 / math.pi) * Psi)
        dfs.append(np.asarray(df, dtype=np.float64))
        psi_ps.append(float(psi_p))
        host.append(dict(PD=PD, PDs=PDs, Psis=Psi.reshape(-1, 1), _psi_p=psi_p))
    sizes = np.array([r.shape[0] for r in raws], dtype=np.int64)
    start = np.zeros(len(jobs) + 1, dtype=np.int32)
    start[1:] = np.cumsum(sizes)
    nS = int(start[-1])
    d_raw = _lib.DeviceArray(ctx, (nS, N * N), np.float32, np.concatenate(raws))
    d_flip = _lib.DeviceArray(ctx, (nS,), np.uint8, np.concatenate(flips))
    d_psi = _lib.DeviceArray(ctx, (nS,), np.float64, np.concatenate(psis))
    d_df = _lib.DeviceArray(ctx, (nS,), np.float64, np.concatenate(dfs))
    d_D = _lib.DeviceArray(ctx, (int((sizes ** 2).sum()),), np.float32)
    d_img = _lib.DeviceArray(ctx, (nS, N, N), np.float32) if want_imgAll else None
    prm = _lib.PdParams(nS=nS, N=N, transposed=1, relion_shift=0, filter_type=FILTERS[filterPar['type']],
                        filter_order=int(filterPar['N']), filter_Qc=float(filterPar['Qc']), pix_size=float(pix_size),
                        Cs=float(Cs), EkV=float(EkV), gaussEnv=float(gaussEnv), AmpContrast=float(AmpContrast), psi_p_deg=0.0,
                        avg_only=0, contraction=0, k_chunk_blocks=0, split_k=0, knn_k=0)
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df, io.D = d_raw.ptr, d_flip.ptr, d_psi.ptr, d_df.ptr, d_D.ptr
    if d_img is not None:
        io.imgAll = d_img.ptr
    pp = np.ascontiguousarray(psi_ps, dtype=np.float64)
    try:
        _lib.check(lib.mem_pd_distance_batch_device(ctx.handle, C.byref(prm), C.byref(io), len(jobs), start.ctypes.data,
                                                    pp.ctypes.data, None))
        Dall = d_D.download()
        img = d_img.download() if d_img is not None else None
    finally:
        for a in (d_raw, d_flip, d_psi, d_df, d_D, d_img):
            if a is not None:
                a.free()
    off = 0
    for g, h in enumerate(host):
        n = int(sizes[g])
        h['D'] = Dall[off:off + n * n].reshape(n, n)
        off += n * n
        if img is not None:
            h['imgAll'] = img[start[g]:start[g + 1]]
    return host
