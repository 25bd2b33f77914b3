"""Stage timings of one PD (C4 shape unless NS / NPIX are set in the environment) under a context option's values
(kernel-variant experiments).    [NS=1000 NPIX=128] python scripts/variant_check.py <option> <v0> <v1> ..."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib   # noqa: E402
import bench                              # noqa: E402

opt = sys.argv[1]
vals = [int(a) for a in sys.argv[2:]]
nS, N = int(os.environ.get('NS', 2000)), int(os.environ.get('NPIX', 256))
lib = _lib.load()
ctx = _lib.Context(0)
pds, rng = bench.make_inputs(nS, N, 1, seed=0)
pd = pds[0]
raw = _lib.DeviceArray(ctx, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32))
flip = _lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip'])
psi = _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg'])
df = _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])
D = _lib.DeviceArray(ctx, (nS, nS), np.float32)
prm = bench.pd_params(_lib, nS, N, float(os.environ.get('PSI_P', pd['psi_p'])))
print('psi_p', prm.psi_p_deg)
io = _lib.PdIO()
io.raw, io.flip, io.psi_deg, io.df, io.D = raw.ptr, flip.ptr, psi.ptr, df.ptr, D.ptr
ref = None
for v in vals:
    ctx.set_option(opt, v)
    best = None
    for r in range(6):
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
        ctx.sync()
        t = ctx.timings()
        if r >= 2 and (best is None or t['device_total'] < best['device_total']):
            best = t
    d = D.download()
    off = ~np.eye(nS, dtype=bool)
    same = 'first' if ref is None else ('identical D' if np.array_equal(d, ref) else 'max off-diagonal rel diff %.2e' % (np.abs(d - ref)[off] / ref[off]).max())
    if ref is None:
        ref = d
    print('%s=%d' % (opt, v), {k: round(x, 3) for k, x in best.items() if k in ('ingest_lowpass', 'align', 'fft_ctf_operands', 'contraction', 'device_total')}, same)
