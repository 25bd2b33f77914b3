"""BASELINE.json configs beyond the bench shape, on one B200 (timings + property checks, no oracle):
  C2  single PD 1,000 x 128^2, full D
  C3  single PD 5,000 x 256^2, kNN k=100 + Ferguson sweep + Laplacian from the resident D
  C5' large box 6,000 x 320^2 (N = 320 path: E = 10 prefilter segments, 102,400-pixel spectra)
  demo-like  53 PDs with the occupancies of the RyR1 demo (117..450, median 206) at N = 128 and 256
python scripts/configs_check.py [c2] [c3] [c5] [demo]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import _lib, DMembeddingII   # noqa: E402
import bench                                              # noqa: E402

lib = _lib.load()
ctx = _lib.Context(0)


def run_shape(nS, N, reps=3, keepD=False):
    pds, rng = bench.make_inputs(nS, N, 1, seed=nS + N)
    pd = pds[0]
    raw = _lib.DeviceArray(ctx, (nS, N * N), np.float32, rng.standard_normal((nS, N * N), dtype=np.float32))
    flip = _lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip'])
    psi = _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg'])
    df = _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])
    D = _lib.DeviceArray(ctx, (nS, nS), np.float32)
    prm = bench.pd_params(_lib, nS, N, pd['psi_p'])
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df, io.D = raw.ptr, flip.ptr, psi.ptr, df.ptr, D.ptr
    t = []
    for r in range(reps + 1):
        ctx.timer_start()
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
        t.append(ctx.timer_stop())
    stage = ctx.timings()
    for a in (raw, flip, psi, df):
        a.free()
    ms = float(np.median(t[1:]))
    print('PD %5d x %d^2: %.3f ms  -> %.3f Gpairs/s   stages %s' %
          (nS, N, ms, nS * nS / ms / 1e6, {k: round(v, 2) for k, v in stage.items() if v > 0}))
    if keepD:
        return D, ms
    D.free()
    return None, ms


if 'c2' in sys.argv:
    run_shape(1000, 128)

if 'c3' in sys.argv:
    nS, k = 5000, 100
    Dd, ms = run_shape(nS, 256, reps=2, keepD=True)
    D = Dd.download().astype(np.float64)
    assert np.array_equal(D, D.T) and np.abs(np.diag(D)).max() < 1e-5 * D.max()
    for _ in range(2):                      # second pass = warm
        t0 = time.time()
        M, logEps, logSumWij, idx, val = DMembeddingII.graph_and_sweep(Dd, k)    # D resident on the device
        t1 = time.time()
        if _ == 0:
            M.free()
    Dd.free()
    # kNN property checks: sorted ascending, self first, every listed neighbour no farther than the k-th
    assert np.array_equal(idx[:, 0], np.arange(nS)) and (np.diff(val[:, 1:], axis=1) >= 0).all()
    rows = np.random.default_rng(0).integers(0, nS, 50)
    for i in rows:
        d = D[i].copy()
        d[i] = -np.inf
        assert set(np.argsort(d, kind='stable')[:k]) == set(idx[i])
    # a15 alone, CUDA events: full sort of every row vs selection on the resident D vs lists taken from the
    # contraction's partial tiles (D never assembled)
    Dd, _ = run_shape(nS, 256, reps=1, keepD=True)
    idx_d = _lib.DeviceArray(ctx, (nS, k), np.int32)
    val_d = _lib.DeviceArray(ctx, (nS, k), np.float64)
    t_knn = {}
    for name, mode in (('sort', 1), ('selection', 2)):
        _lib.check(lib.mem_knn_mode(mode))
        for r in range(3):
            ctx.timer_start()
            _lib.check(lib.mem_knn_device_f32(ctx.handle, Dd.ptr, nS, k, idx_d.ptr, val_d.ptr, None))
            t_knn[name] = ctx.timer_stop()
        if mode == 1:
            idx_sort, val_sort = idx_d.download(), val_d.download()
    _lib.check(lib.mem_knn_mode(0))
    assert np.array_equal(idx_sort, idx_d.download()) and np.array_equal(val_sort, val_d.download())
    Dd.free()
    pds, rng = bench.make_inputs(nS, 256, 1, seed=nS + 256)
    pd = pds[0]
    raw = _lib.DeviceArray(ctx, (nS, 256 * 256), np.float32, rng.standard_normal((nS, 256 * 256), dtype=np.float32))
    small = [_lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip']), _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg']),
             _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])]
    prm = bench.pd_params(_lib, nS, 256, pd['psi_p'])
    prm.knn_k = k
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df = raw.ptr, small[0].ptr, small[1].ptr, small[2].ptr
    io.knn_idx, io.knn_val = idx_d.ptr, val_d.ptr
    for r in range(3):
        ctx.timer_start()
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
        t_fused = ctx.timer_stop()
    assert np.array_equal(idx_sort, idx_d.download()) and np.array_equal(val_sort, val_d.download())
    print('C3 a15 (k=%d) on the resident D: bitonic sort %.2f ms, radix selection %.2f ms; whole PD with the lists '
          'selected from the partial tiles, no D: %.2f ms (PD with D: %.2f ms); lists identical' %
          (k, t_knn['sort'], t_knn['selection'], t_fused, ms))
    for a in [raw, idx_d, val_d] + small:
        a.free()
    sigma = 3.0 * np.sqrt(np.median(val[:, 1:]))
    t1b = time.time()
    L = DMembeddingII.laplacian(M, nS, sigma)
    t2 = time.time()
    M.free()
    assert np.allclose(L, L.T) and np.isfinite(L).all()
    print('C3 kNN(k=%d)+graph+compaction+Ferguson sweep on the resident D: %.1f ms; Laplacian + 200 MB D2H of L: %.1f ms' %
          (k, (t1 - t0) * 1e3, (t2 - t1b) * 1e3))

if 'c5' in sys.argv:
    run_shape(6000, 320, reps=1)

if 'c5full' in sys.argv:
    # BASELINE config 5, one PD at full size: 20,000 particles x 320^2 (8.2 GB of raw particles, ~60 GB of workspace).
    # The particle stack is generated on the device with torch (host RNG would take minutes); torch only allocates.
    import torch
    nS, N = 20000, 320
    pds, rng = bench.make_inputs(nS, N, 1, seed=5)
    pd = pds[0]
    g = torch.Generator(device='cuda:0').manual_seed(5)
    raw_t = torch.randn((nS, N * N), dtype=torch.float32, device='cuda:0', generator=g)
    D_t = torch.empty((nS, nS), dtype=torch.float32, device='cuda:0')
    torch.cuda.synchronize()
    flip = _lib.DeviceArray(ctx, (nS,), np.uint8, pd['flip'])
    psi = _lib.DeviceArray(ctx, (nS,), np.float64, pd['psi_deg'])
    df = _lib.DeviceArray(ctx, (nS,), np.float64, pd['df'])
    prm = bench.pd_params(_lib, nS, N, pd['psi_p'])
    io = _lib.PdIO()
    io.raw, io.flip, io.psi_deg, io.df, io.D = raw_t.data_ptr(), flip.ptr, psi.ptr, df.ptr, D_t.data_ptr()
    for r in range(2):
        ctx.timer_start()
        _lib.check(lib.mem_pd_distance_device(ctx.handle, C.byref(prm), C.byref(io), None))
        ms = ctx.timer_stop()
    free, total = torch.cuda.mem_get_info(0)
    sub = D_t[:2000, :2000]
    assert torch.equal(sub, sub.T) and float(D_t.diagonal().abs().max()) < 1e-5 * float(D_t.max())
    assert bool(torch.isfinite(D_t).all()) and float(D_t.min()) > -1e-5 * float(D_t.max())
    print('C5 PD %d x %d^2: %.1f ms -> %.3f Gpairs/s; stages %s; device memory in use %.1f GB of %.1f GB' %
          (nS, N, ms, nS * nS / ms / 1e6, {k: round(v, 1) for k, v in ctx.timings().items() if v > 0},
           (total - free) / 1e9, total / 1e9))

if 'demo' in sys.argv:
    for N in (128, 256):
        for nS in (117, 206, 450):      # min / median / max occupancy of the demo's 53 PDs
            run_shape(nS, N, reps=5)
