"""BASELINE config 1 (the demo's 53 PDs, 117..450 particles, synthetic images at box N) on one GPU, inputs resident:
serial, S PDs in flight on S streams, and (when built) the batched entry point.
    python scripts/config1_check.py [N] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib, workloads   # noqa: E402
import bench                                          # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
pds, em = workloads.demo_config1()
bench.EM.update(em)
ctx0 = _lib.Context(0)
rng = np.random.default_rng(0)
dev = []
for pd in pds:
    nS = len(pd['ind'])
    dev.append(dict(raw=_lib.DeviceArray(ctx0, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32)),
                    flip=_lib.DeviceArray(ctx0, (nS,), np.uint8, pd['flip']),
                    psi=_lib.DeviceArray(ctx0, (nS,), np.float64, pd['psi_deg']),
                    df=_lib.DeviceArray(ctx0, (nS,), np.float64, pd['df']),
                    D=_lib.DeviceArray(ctx0, (nS, nS), np.float32), prm=bench.pd_params(_lib, nS, N, pd['psi_p'])))
pairs = sum(len(pd['ind']) ** 2 for pd in pds)
order = sorted(range(len(pds)), key=lambda i: -len(pds[i]['ind']))
for S in (1, 2, 4, 8):
    ctxs = [ctx0] + [_lib.Context(0) for _ in range(S - 1)]
    best = 1e9
    for rep in range(reps + 1):
        for c in ctxs:
            c.sync()
        t0 = time.perf_counter()
        for k, i in enumerate(order):
            d = dev[i]
            io = _lib.PdIO()
            io.raw, io.flip, io.psi_deg, io.df, io.D = d['raw'].ptr, d['flip'].ptr, d['psi'].ptr, d['df'].ptr, d['D'].ptr
            _lib.check(lib.mem_pd_distance_device(ctxs[k % S].handle, C.byref(d['prm']), C.byref(io), None))
        for c in ctxs:
            c.sync()
        if rep:
            best = min(best, time.perf_counter() - t0)
    print('N=%d streams=%d: 53 PDs (%d particles, %d pairs) in %.2f ms -> %.3f Gpairs/s'
          % (N, S, sum(len(p['ind']) for p in pds), pairs, best * 1e3, pairs / best / 1e9))
    for c in ctxs[1:]:
        c.close()
