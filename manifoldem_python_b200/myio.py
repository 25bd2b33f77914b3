"""Record in/out with the reference's semantics (modules/myio.py:21-59) and, on request, a lighter on-disk layout.

fin1(filename) -> dict or None on ANY failure; fout1(filename, keys, values); fout2(filename, dict) — the three
functions all 57 call sites of the reference use.

Two layouts, chosen per write (argument `layout`, else `p.record_layout`, else $MANIFOLDEM_B200_RECORD, else
'pickle'):

'pickle'   one pickle.HIGHEST_PROTOCOL file of {key: value}: byte-compatible with the reference's own myio, so
           un-replaced reference modules keep reading it.  A C4-sized PD record is 3.2 GB of float64 that every
           consumer unpickles whole, although manifoldTrimmingAuto reads only D and ind.

'sidecar'  SURVEY.md §8f rank 3.  `filename` stays a small pickle (scalars, small arrays and a manifest); every
           array of >= SIDECAR_MIN_BYTES goes to `filename.<key>.npy` in the dtype it was computed in (float32
           from the GPU: half the bytes, no host-side float64 copy before the dump).  fin1 returns a `Record`
           (a dict) that opens a sidecar only when its key is read and promotes it to the dtype the reference
           stores (float64) at that moment — the values are identical to the 'pickle' layout (float32 -> float64
           is exact), the consumer of D no longer pays for imgAll / imgAllFlip / CTF.  Needs this module in place
           of the reference's (manifoldem_python_b200/dropin/myio.py); the reference's own fin1 would return the
           manifest dict and fail loudly on data['D'].

Write order keeps the resume protocol (getDistanceCTF...py:415-419): sidecars, then the manifest through a
temporary file + rename, and only then does the caller touch its marker.
"""
import os
import pickle

import numpy as np

MAGIC = '__manifoldem_b200_record__'
SIDECAR_MIN_BYTES = 1 << 16
_LAYOUTS = ('pickle', 'sidecar')


def default_layout():
    cfgs = []
    try:
        import p                                   # the reference's config module when running inside ManifoldEM
        if hasattr(p, 'nPix'):
            cfgs.append(p)
    except ImportError:
        pass
    from . import p as own
    cfgs.append(own)
    for cfg in cfgs:
        lay = getattr(cfg, 'record_layout', None)
        if lay:
            return lay
    return os.environ.get('MANIFOLDEM_B200_RECORD', 'pickle')


class Record(dict):
    """dict of a 'sidecar' record: heavy arrays are read (np.load, memory-mapped) and promoted on first access."""

    def __init__(self, small, arrays, base, keys=None):
        super().__init__()
        self._lazy = dict(arrays)
        self._base = base
        for k in (keys or list(small) + list(arrays)):   # the writer's key order; keys(), len(), `in` as a plain dict
            super().__setitem__(k, small[k] if k in small else None)

    def _load(self, k):
        spec = self._lazy.pop(k)
        if spec.get('virtual') == 'ctf':
            a = _ctf_field(self[spec['df_key']], spec)
            super().__setitem__(k, a)
            return a
        if spec.get('virtual') == 'images':
            # one pass of the distance stage's own kernels rebuilds every image array of the record that is still virtual
            names = [k] + [o for o, sp in self._lazy.items() if sp.get('virtual') == 'images']
            out = _image_fields(self, spec, names)
            for o in names:
                self._lazy.pop(o, None)
                super().__setitem__(o, out[o])
            return out[k]
        a = np.load(os.path.join(os.path.dirname(self._base), spec['file']), mmap_mode='r')
        if tuple(a.shape) != tuple(spec['shape']):
            raise IOError('sidecar %s has shape %s, manifest says %s' % (spec['file'], a.shape, spec['shape']))
        a = np.array(a, dtype=spec.get('promote') or a.dtype)     # private, writeable (consumers mutate D in place)
        super().__setitem__(k, a)
        return a

    def __getitem__(self, k):
        if k in self._lazy:
            return self._load(k)
        return super().__getitem__(k)

    def get(self, k, default=None):
        return self[k] if k in self else default

    # --- writes: a key that is assigned, deleted or popped stops being lazy (the stored value must win over the sidecar)
    def __setitem__(self, k, v):
        self._lazy.pop(k, None)
        super().__setitem__(k, v)

    def __delitem__(self, k):
        self._lazy.pop(k, None)
        super().__delitem__(k)

    def pop(self, k, *default):
        if k in self._lazy:
            self._load(k)
        return super().pop(k, *default)

    def popitem(self):
        self.materialise()
        return super().popitem()

    def setdefault(self, k, default=None):
        if k in self:
            return self[k]
        self[k] = default
        return default

    def update(self, *args, **kw):
        for k, v in dict(*args, **kw).items():
            self[k] = v

    def __eq__(self, other):
        return dict.__eq__(self.materialise(), other.materialise() if isinstance(other, Record) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return 'Record(%s)' % ', '.join('%r: %s' % (k, '<lazy>' if k in self._lazy else repr(dict.__getitem__(self, k)))
                                         for k in dict.keys(self))

    def raw(self, k):
        """The array as stored (no promotion, read-only memory map) — for consumers that keep float32."""
        if k in self._lazy and not self._lazy[k].get('virtual'):
            spec = self._lazy[k]
            return np.load(os.path.join(os.path.dirname(self._base), spec['file']), mmap_mode='r')
        return super().__getitem__(k)

    def materialise(self):
        for k in list(self._lazy):
            self._load(k)
        return self

    def items(self):
        self.materialise()
        return super().items()

    def values(self):
        self.materialise()
        return super().values()

    def __iter__(self):                            # defined so that dict(record) / {**record} go through __getitem__
        return super().__iter__()                  # (CPython copies a dict subclass's raw slots unless __iter__ is overridden)

    def copy(self):
        return dict(self.materialise())

    def __reduce__(self):                          # pickling a Record stores the plain dict
        return (dict, (dict(self.materialise()),))


def _ctf_field(df, spec):
    """A record written with virtual={'CTF': ...} keeps df and the microscope constants instead of the 8 N^2 bytes per
    particle of the CTF field (ctemh_cryoFrank.op; 1 GB of a C4-sized record): the field is produced when the key is
    read, by the kernel that produced it inside the distance stage (C ABI mem_ctf_host) — bit-identical values."""
    import ctypes as C
    from . import _lib
    from .getDistanceCTF_local_Conj9combinedS2 import _ctx
    lib = _lib.load()
    df = np.ascontiguousarray(df, dtype=np.float64)
    nS, N = df.shape[0], int(spec['N'])
    prm = _lib.PdParams(nS=nS, N=N, pix_size=float(spec['pix_size']), Cs=float(spec['Cs']), EkV=float(spec['EkV']),
                        gaussEnv=float(spec['gaussEnv']), AmpContrast=float(spec['AmpContrast']))
    out = np.empty((nS, N * N), dtype=np.float64)
    _lib.check(lib.mem_ctf_host(_ctx().handle, C.byref(prm), df.ctypes.data, out.ctypes.data))
    return out.reshape(spec['shape'])


def _image_fields(rec, spec, names):
    """A record written with virtual={'imgAll': ..., 'imgAllFlip': ...} does not store the aligned / phase-flipped image
    stacks (2 x 4 N^2 bytes per particle — 1 GB of a C4-sized record, and what bounds the distance stage from disk to disk:
    the box writes 2.3 GB/s) but the recipe: the particle stack's path and the PD's ind / q / df are in the record already.
    The arrays are produced when a key is read, by the kernels that produced them inside the distance stage
    (pd_stage.run_pd on the same inputs: deterministic, bit-identical values, ~0.1 s for a C4-sized PD — about what reading
    1 GB back from disk costs).  The stack file must still be where it was."""
    from . import pd_stage
    from .getDistanceCTF_local_Conj9combinedS2 import _ctx
    N = int(spec['N'])
    stack = pd_stage.open_stack(spec['stack'], N, bool(spec['relion']))
    msk2 = rec['msk2']
    msk2 = None if np.ndim(msk2) == 0 else np.asarray(msk2)
    sh = None
    if spec.get('sh_members') is not None:              # back to arrays indexed by particle, as run_pd takes them
        ind = np.asarray(rec['ind'])
        half = int(spec['nStot']) // 2
        base = np.where(ind >= half, ind - half, ind)
        sh = (np.zeros(half), np.zeros(half))
        sh[0][base], sh[1][base] = spec['sh_members'][0], spec['sh_members'][1]
    res = pd_stage.run_pd(rec['ind'], rec['q'], rec['df'], stack, int(spec['nStot']), N, spec['pix_size'], spec['Cs'], spec['EkV'],
                          spec['AmpContrast'], gaussEnv=spec['gaussEnv'], filterPar=spec['filterPar'], msk2=msk2,
                          relion=bool(spec['relion']), sh=sh, fields=tuple(names), float64=True, ctx=_ctx(), intensity=False)
    return {o: res[o] for o in names}


def _sidecars_ok(arrays, filename):
    """Every stored sidecar exists and has at least the bytes its manifest shape implies."""
    for spec in arrays.values():
        if spec.get('virtual'):
            continue
        path = os.path.join(os.path.dirname(filename), spec['file'])
        need = int(np.prod(spec['shape'])) * np.dtype(spec['dtype']).itemsize
        if not os.path.isfile(path) or os.path.getsize(path) < need:
            return False
    return True


def fin1(filename):
    """modules/myio.py:21-30: the file is opened OUTSIDE the try (a missing file raises, as in the reference); any
    failure while reading it — a truncated pickle, or for 'sidecar' records a missing / short sidecar — gives None."""
    with open(filename, 'rb') as f:
        try:
            data = pickle.load(f)
            if isinstance(data, dict) and data.get(MAGIC) == 2:
                if not _sidecars_ok(data['arrays'], filename):
                    return None
                return Record(data['small'], data['arrays'], filename, data.get('keys'))
            return data
        except Exception:
            return None


def _write_npy(path, a):
    a = np.ascontiguousarray(a)
    with open(path, 'wb') as f:
        np.lib.format.write_array_header_1_0(f, np.lib.format.header_data_from_array_1_0(a))
        f.write(memoryview(a).cast('B'))


def fout1(filename, key_list, v_list, layout=None, promote=None, virtual=None):
    """`promote`: {key: dtype} the reader restores for that key (e.g. float32 on disk -> float64 like the reference).
    `virtual` ('sidecar' only): {key: spec} for fields that are not stored but produced when read; known spec:
    dict(virtual='ctf', df_key, N, pix_size, Cs, EkV, gaussEnv, AmpContrast, shape)."""
    layout = layout or default_layout()
    if layout not in _LAYOUTS:
        raise ValueError('record layout %r (known: %s)' % (layout, ', '.join(_LAYOUTS)))
    if layout == 'pickle':
        with open(filename, 'wb') as f:
            pickle.dump(dict(zip(key_list, v_list)), f, protocol=pickle.HIGHEST_PROTOCOL)
        return
    promote = promote or {}
    virtual = virtual or {}
    small, arrays = {}, {}
    base = os.path.basename(filename)
    for k, v in zip(key_list, v_list):
        if k in virtual:
            arrays[k] = dict(virtual[k])
        elif isinstance(v, np.ndarray) and v.dtype != object and v.nbytes >= SIDECAR_MIN_BYTES:
            name = '%s.%s.npy' % (base, k)
            _write_npy(os.path.join(os.path.dirname(filename), name), v)
            arrays[k] = dict(file=name, shape=tuple(v.shape), dtype=str(v.dtype),
                             promote=str(np.dtype(promote[k])) if k in promote else None)
        elif isinstance(v, np.ndarray) and k in promote:
            small[k] = v.astype(promote[k])
        else:
            small[k] = v
    # sidecars left by an earlier write of this record whose keys are no longer stored as files
    keep = {spec['file'] for spec in arrays.values() if 'file' in spec}
    d = os.path.dirname(filename) or '.'
    for name in os.listdir(d):
        if name.startswith(base + '.') and name.endswith('.npy') and name not in keep:
            os.remove(os.path.join(d, name))
    tmp = filename + '.tmp'
    with open(tmp, 'wb') as f:
        pickle.dump({MAGIC: 2, 'keys': list(key_list), 'small': small, 'arrays': arrays}, f, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, filename)


def fout2(filename, d, layout=None):
    fout1(filename, list(d.keys()), list(d.values()), layout=layout)
