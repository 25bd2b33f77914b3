"""New rotation kernels (rotate.cu) against the generic k_rotate on one PD: imgAll and D of both paths, stage timings.
    python scripts/rotate_check.py [nS] [N]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib   # noqa: E402
import bench                              # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lib = _lib.load()
ctx = _lib.Context(0)
pds, rng = bench.make_inputs(nS, N, 1, seed=0)
pd = pds[0]
raw = _lib.DeviceArray(ctx, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32))
flip = _lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip'])
psi = _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg'])
df = _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])
D = _lib.DeviceArray(ctx, (nS, nS), np.float32)
img = _lib.DeviceArray(ctx, (nS, N * N), np.float32)
prm = bench.pd_params(_lib, nS, N, pd['psi_p'])
io = _lib.PdIO()
io.raw, io.flip, io.psi_deg, io.df, io.D, io.imgAll = raw.ptr, flip.ptr, psi.ptr, df.ptr, D.ptr, img.ptr
out = {}
for legacy in (1, 0):
    ctx.set_option('legacy_rotate', legacy)
    for r in range(3):
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
        ctx.sync()
    print('legacy_rotate=%d' % legacy, {k: round(v, 3) for k, v in ctx.timings().items()})
    out[legacy] = (img.download(), D.download())
a, b = out[1][0], out[0][0]
print('imgAll: identical' if np.array_equal(a, b) else 'imgAll: max |diff| / max = %.3e' % (np.abs(a - b).max() / np.abs(a).max()))
a, b = out[1][1], out[0][1]
print('D: identical' if np.array_equal(a, b) else 'D: max rel diff = %.3e' % (np.abs(a - b)[a > 0] / a[a > 0]).max())
