"""Run a few PDs of a given shape through the device API (inputs resident) — the command ncu wraps.
    python scripts/one_pd.py [nS] [N] [reps] [knn_k]     knn_k > 0: neighbour lists instead of D (never assembled)
"""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
knn_k = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = _lib.load()
ctx = _lib.Context(0)
pds, rng = bench.make_inputs(nS, N, 1, seed=0)
pd = pds[0]
raw = _lib.DeviceArray(ctx, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32))
flip = _lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip'])
psi = _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg'])
df = _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])
D = _lib.DeviceArray(ctx, (nS, nS), np.float32)
prm = bench.pd_params(_lib, nS, N, pd['psi_p'])
io = _lib.PdIO()
io.raw, io.flip, io.psi_deg, io.df, io.D = raw.ptr, flip.ptr, psi.ptr, df.ptr, D.ptr
if knn_k:
    idx = _lib.DeviceArray(ctx, (nS, knn_k), np.int32)
    val = _lib.DeviceArray(ctx, (nS, knn_k), np.float64)
    prm.knn_k, io.D, io.knn_idx, io.knn_val = knn_k, None, idx.ptr, val.ptr
for r in range(reps):
    _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
    ctx.sync()
    print('rep', r, {k: round(v, 3) for k, v in ctx.timings().items()})
