"""Host-side measurement of the per-PD record (no GPU involved): the reference's layout — one
pickle of float64 arrays (modules/myio.py:38-51, keys of getDistanceCTF...py:409-412) — against the 'sidecar' layout of
manifoldem_python_b200/myio.py, for (i) the writer after a PD of the distance stage and (ii) the reader
manifoldTrimmingAuto.py:44-46, which needs D and ind only.

    python scripts/record_io_bench.py [nS] [N] [dir]        default 2000 256 (BASELINE config 4 PD: 3.2 GB record)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import myio                                   # noqa: E402
from manifoldem_python_b200.getDistanceCTF_local_Conj9combinedS2 import _KEYS   # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out = sys.argv[3] if len(sys.argv) > 3 else '/tmp/record_io_bench'
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(0)
f32 = lambda *s: rng.random(s, dtype=np.float32)
gpu = dict(D=f32(nS, nS), imgAll=f32(nS, N, N), imgAllFlip=f32(nS, N, N), imgAvg=f32(N, N), imgAvgFlip=f32(N, N),
           imgAllIntensity=f32(N, N))                                     # what the C ABI hands back (float32)
rest = dict(ind=np.arange(nS), q=rng.random((4, nS)), df=rng.random(nS), CTF=rng.random((nS, N * N)), msk2=1,
            PD=rng.random(3), PDs=rng.random((3, nS)), Psis=rng.random((nS, 1)), imgLabels=np.ones(nS, int),
            Dnom=rng.random((nS, 1)), Nom=rng.random((nS, 1)), version='v', options={})


def drop_cache(path):
    fd = os.open(path, os.O_RDONLY)
    try:
        os.fsync(fd)
        os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
    finally:
        os.close(fd)


def files_of(base):
    d = os.path.dirname(base)
    return [os.path.join(d, f) for f in os.listdir(d) if f.startswith(os.path.basename(base))]


res = {}
for layout in ('pickle', 'sidecar'):
    base = os.path.join(out, 'IMGs_%s_prD_0' % layout)
    t0 = time.time()
    if layout == 'pickle':                                                # float64 copies, then one pickle
        rec = dict(rest)
        rec.update({k: v.astype(np.float64) for k, v in gpu.items()})
        myio.fout1(base, _KEYS, [rec[k] for k in _KEYS], layout='pickle')
    else:                                                                 # float32 straight to .npy, promoted on read
        rec = dict(rest)
        rec.update(gpu)
        myio.fout1(base, _KEYS, [rec[k] for k in _KEYS], layout='sidecar', promote={k: np.float64 for k in gpu})
    t_write = time.time() - t0
    del rec
    size = sum(os.path.getsize(f) for f in files_of(base))
    t_read = {}
    for cache in ('warm', 'cold'):
        if cache == 'cold':
            for f in files_of(base):
                drop_cache(f)
        t0 = time.time()
        data = myio.fin1(base)
        D, ind = data['D'], data['ind']                                   # manifoldTrimmingAuto.py:45-46
        assert D.dtype == np.float64 and D.shape == (nS, nS) and np.array_equal(D, gpu['D'].astype(np.float64))
        t_read[cache] = time.time() - t0
        del data, D
    res[layout] = (t_write, size, t_read)
    print('%-8s write %6.2f s (%5.2f GB on disk, %5.2f GB/s)   read D + ind: page cache %6.3f s, after fadvise(DONTNEED) %6.3f s'
          % (layout, t_write, size / 1e9, size / 1e9 / t_write, t_read['warm'], t_read['cold']))
    for f in files_of(base):
        os.remove(f)
p_, s_ = res['pickle'], res['sidecar']
print('PD of %d x %d^2: record write %.1fx faster, %.1fx smaller; the consumer of D reads %.0fx faster (cold %.0fx)'
      % (nS, N, p_[0] / s_[0], p_[1] / s_[1], p_[2]['warm'] / s_[2]['warm'], p_[2]['cold'] / s_[2]['cold']))
