"""Drop-in for the reference's per-PD worker (modules/getDistanceCTF_local_Conj9combinedS2.py:216-420).

Same call signature, same implicit inputs from the module-global config `p`, same side effects: a pickle at
`outFile` with the reference's 19 keys (float64 arrays, same shapes), then an empty marker file
`p.dist_prog/<prD>` written only AFTER the pickle dump returned (resume protocol, :415-419).
All arithmetic between "images gathered" and "arrays ready to pickle" runs on the GPU through the C ABI.
"""
import os
import threading

import numpy as np

from . import myio, pd_stage
from . import _lib

version = pd_stage.VERSION

_KEYS = ['D', 'ind', 'q', 'df', 'CTF', 'imgAll', 'msk2', 'PD', 'PDs', 'Psis', 'imgAvg', 'imgAvgFlip',
         'imgAllFlip', 'imgLabels', 'Dnom', 'Nom', 'imgAllIntensity', 'version', 'options']

_tls = threading.local()
_slots = {}                      # device -> list of idle (Context, HostArena) pairs; they live as long as the process
_slots_lock = threading.Lock()


def _cfg():
    """The reference's module-global config when running inside ManifoldEM, else this package's p."""
    try:
        import p                      # noqa: WPS433  (modules/p.py when modules/ is on sys.path)
        if hasattr(p, 'nPix'):
            return p
    except ImportError:
        pass
    from . import p as own
    return own


def _device():
    return int(os.environ.get('MANIFOLDEM_B200_DEVICE', '0'))


class _Slot:
    """A CUDA context (stream + device workspace + FFT plans) and a pinned host arena, checked out by one host thread
    for the duration of one op() call and handed back afterwards.  The reference calls op() from short-lived worker
    threads / processes (GetDistancesS2.py:110-113); keeping the slots at module level is what lets the 5 GB device
    workspace, the cuFFT plans and the pinned staging buffers of PD k serve PD k + 1 (with
    thread-local contexts every GetDistancesS2.op call paid for them again)."""

    def __init__(self, dev):
        self.ctx = _lib.Context(dev)
        self.arena = pd_stage.HostArena() if os.environ.get('MANIFOLDEM_B200_PINNED_ARENA', '1') != '0' else None


class _checkout:
    def __enter__(self):
        dev = _device()
        with _slots_lock:
            free = _slots.setdefault(dev, [])
            self.slot = free.pop() if free else None
        if self.slot is None:
            self.slot = _Slot(dev)
        self.dev = dev
        return self.slot

    def __exit__(self, *exc):
        with _slots_lock:
            _slots[self.dev].append(self.slot)
        return False


def _ctx():
    """A context for callers outside op() (myio's virtual CTF field): one per host thread and device."""
    dev = _device()
    ctx = getattr(_tls, 'ctx', None)
    if ctx is None or ctx.device != dev:
        ctx = _lib.Context(dev)
        _tls.ctx = ctx
    return ctx


def _read_mrc_volume(path):
    """3-D mask volume (:305-306): MRC2014, modes 0/1/2/6, as mrcfile's .data (z,y,x)."""
    hdr = np.fromfile(path, dtype='<i4', count=256)
    nx, ny, nz, mode, nsymbt = int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[23])
    dt = {0: np.int8, 1: '<i2', 2: '<f4', 6: '<u2'}[mode]
    return np.fromfile(path, dtype=dt, offset=1024 + nsymbt, count=nx * ny * nz).reshape(nz, ny, nx)


def op(input_data, filterPar, imgFileName, sh, nStot, options, fields=None):
    """[ind, q(4,nS), df(nS), outFile, prD], filterPar{'type','Qc','N'}, stack path, (shx, shy), augmented
    particle count, options{'verbose','avgOnly','visual','parallel','relion_data','thres'} -> None.

    `fields` (extension, default None = everything the reference stores) may name a subset of the heavy
    per-image arrays ('D','imgAll','imgAllFlip','CTF') to materialise; the others are stored as None."""
    p = _cfg()
    ind, q, df, outFile, prD = input_data[0], input_data[1], input_data[2], input_data[3], input_data[4]
    N = int(p.nPix)
    relion = bool(options.get('relion_data', False))
    stack = pd_stage.open_stack(imgFileName, N, relion)
    angles = pd_stage.host_angles(np.asarray(q, dtype=np.float64))
    msk2 = None
    if getattr(p, 'mask_vol_file', ''):                                   # :303-310
        from . import projectMask
        msk2 = projectMask.op(_read_mrc_volume(p.mask_vol_file), angles[1])
    # 'sidecar' records (myio.py) keep the float32 the GPU produced and promote on read; 'pickle' records hold the
    # reference's float64 arrays
    layout = myio.default_layout()
    want = tuple(fields or ('D', 'imgAll', 'imgAllFlip', 'CTF'))
    virtual = {}
    if layout == 'sidecar':
        # p.record_skip: heavy arrays not to store at all (e.g. ('imgAllFlip',): no consumer reads it);
        # p.record_virtual_ctf (default on): keep df and the microscope constants, rebuild the CTF field when it is read
        want = tuple(f for f in want if f not in tuple(getattr(p, 'record_skip', ())))
        if 'CTF' in want and getattr(p, 'record_virtual_ctf', True) and not options.get('avgOnly', False):
            nS_ = len(np.asarray(ind))
            virtual['CTF'] = dict(virtual='ctf', df_key='df', N=N, pix_size=float(p.pix_size), Cs=float(p.Cs),
                                  EkV=float(p.EkV), gaussEnv=float(getattr(p, 'gaussEnv', np.inf)),
                                  AmpContrast=float(p.AmpContrast),
                                  shape=(nS_, N, N) if options.get('parallel') else (nS_, N * N))
            want = tuple(f for f in want if f != 'CTF')
        # p.record_virtual_images (opt-in): keep the recipe instead of imgAll / imgAllFlip — the stack's path; ind, q, df are in
        # the record anyway — and let myio rebuild them with the same kernels when a consumer reads the key (myio._image_fields)
        if getattr(p, 'record_virtual_images', False) and not options.get('avgOnly', False):
            sh_pd = None
            if relion:                                  # the shifts of this PD's members only (sh is indexed by particle)
                base_idx = np.where(np.asarray(ind) >= nStot // 2, np.asarray(ind) - nStot // 2, np.asarray(ind))
                sh_pd = (np.asarray(sh[0])[base_idx].copy(), np.asarray(sh[1])[base_idx].copy())
            for name in ('imgAll', 'imgAllFlip'):
                if name in want:
                    virtual[name] = dict(virtual='images', stack=os.path.abspath(imgFileName), relion=relion, nStot=int(nStot), N=N,
                                         pix_size=float(p.pix_size), Cs=float(p.Cs), EkV=float(p.EkV),
                                         gaussEnv=float(getattr(p, 'gaussEnv', np.inf)), AmpContrast=float(p.AmpContrast),
                                         filterPar=dict(filterPar), sh_members=sh_pd, shape=(len(np.asarray(ind)), N, N))
            want = tuple(f for f in want if f not in virtual)
    with _checkout() as slot:      # the arrays of `res` live in the slot's pinned arena until the record is on disk
        res = pd_stage.run_pd(ind, q, df, stack, nStot, N, p.pix_size, p.Cs, p.EkV, p.AmpContrast,
                              gaussEnv=getattr(p, 'gaussEnv', np.inf), filterPar=filterPar, msk2=msk2, relion=relion,
                              sh=sh, avg_only=bool(options.get('avgOnly', False)), ctx=slot.ctx, angles=angles,
                              fields=want, float64=(layout != 'sidecar'), arena=slot.arena,
                              intensity=True if virtual else None)
        if (options.get('parallel') or options.get('avgOnly')) and res['CTF'] is not None:
            res['CTF'] = res['CTF'].reshape(-1, N, N)     # only the default non-avgOnly branch flattens CTF (:392-393)
        res['options'] = options
        promote = {k: np.float64 for k in _KEYS if isinstance(res[k], np.ndarray) and res[k].dtype == np.float32}
        myio.fout1(outFile, _KEYS, [res[k] for k in _KEYS], layout=layout, promote=promote, virtual=virtual)
    # marker AFTER the dump: signifies a non-corrupted pickle (:415-419)
    open(os.path.join(p.dist_prog, '%s' % (prD)), 'a').close()
