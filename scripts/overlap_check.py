"""Does running S PDs concurrently (S contexts = S streams, one host thread each) on ONE GPU raise the
device-resident throughput?  The tensor-bound contraction of one PD can overlap the HBM-bound passes of another
if the block scheduler co-schedules them.   python scripts/overlap_check.py [nS] [N] [pds]"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib   # noqa: E402
import bench                              # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
PDS = int(sys.argv[3]) if len(sys.argv) > 3 else 48
lib = _lib.load()
pds, rng = bench.make_inputs(nS, N, 2, seed=0)
ctx0 = _lib.Context(0)
raw = [_lib.DeviceArray(ctx0, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32)) for _ in pds]
flip = [_lib.DeviceArray(ctx0, (nS,), np.uint8, pd['flip']) for pd in pds]
psi = [_lib.DeviceArray(ctx0, (nS,), np.float64, pd['psi_deg']) for pd in pds]
df = [_lib.DeviceArray(ctx0, (nS,), np.float64, pd['df']) for pd in pds]
prms = [bench.pd_params(_lib, nS, N, pd['psi_p']) for pd in pds]

for S in (1, 2, 3, 4):
    ctxs = [ctx0] + [_lib.Context(0) for _ in range(S - 1)]
    Ds = [_lib.DeviceArray(ctx0, (nS, nS), np.float32) for _ in range(S)]

    def worker(w, count):
        for k in range(count):
            j = (w + k) % len(pds)
            io = _lib.PdIO()
            io.raw, io.flip, io.psi_deg, io.df, io.D = raw[j].ptr, flip[j].ptr, psi[j].ptr, df[j].ptr, Ds[w].ptr
            _lib.check(lib.mem_pd_distance_device(ctxs[w].handle, C.byref(prms[j]), C.byref(io), None))
        ctxs[w].sync()

    for rep in range(2):
        th = [threading.Thread(target=worker, args=(w, PDS // S)) for w in range(S)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        dt = time.perf_counter() - t0
    n = (PDS // S) * S
    print('streams=%d: %d PDs in %.1f ms -> %.3f ms/PD, %.3f Gpairs/s' % (S, n, dt * 1e3, dt * 1e3 / n, n * nS * nS / dt / 1e9))
    for c in ctxs[1:]:
        c.close()
    for d in Ds:
        d.free()
