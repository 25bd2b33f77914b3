"""GPU bring-up report: per-field errors of the CUDA path against the oracle, SIMT vs tcgen05
contraction, tensor-core accumulation error vs K, quick timings.  Run on the GPU box:
    python scripts/gpu_debug.py [quick]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifoldem_python_b200 import _lib, synthetic, pd_stage   # noqa: E402
from oracle import pd_distance as opd                            # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def report_pd(nS, N, seed, snr, contraction, chunk=0, split=0, msk=False, impl='periodic'):
    pd = synthetic.make_pd(nS, N, seed=seed, snr=snr)
    em = pd['em']
    msk2 = None
    if msk:
        yy, xx = np.mgrid[:N, :N]
        msk2 = ((yy - N / 2) ** 2 / (0.4 * N) ** 2 + (xx - N / 2) ** 2 / (0.3 * N) ** 2) < 1
    t0 = time.time()
    res = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                          em['EkV'], em['AmpContrast'], msk2=msk2, contraction=contraction, k_chunk_blocks=chunk,
                          split_k=split)
    t1 = time.time()
    ref = opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                          em['EkV'], em['AmpContrast'], msk2=1 if msk2 is None else msk2, rotate_impl=impl)
    t2 = time.time()
    D, Dr = res['D'], ref['D']
    off = ~np.eye(nS, dtype=bool)
    relD = np.abs(D - Dr)[off] / Dr[off]
    print('PD nS=%d N=%d snr=%g contraction=%d chunk=%d split=%d msk=%d: gpu %.2fs oracle %.2fs' %
          (nS, N, snr, contraction, chunk, split, msk, t1 - t0, t2 - t1))
    print('   D offdiag rel err: max %.3e  median %.3e   diag abs/maxD %.3e   min D/(a+b)~ %.3f' %
          (relD.max(), np.median(relD), np.abs(np.diag(D) - np.diag(Dr)).max() / Dr.max(),
           Dr[off].min() / Dr.max()))
    for k in ('imgAll', 'imgAllFlip', 'CTF', 'imgAvg', 'imgAvgFlip', 'imgAllIntensity'):
        print('   %-16s rel-to-max err %.3e' % (k, rel(res[k].reshape(ref[k].shape), ref[k])))
    print('   timings(ms):', {k: round(v, 3) for k, v in _lib.default_context().timings().items()})
    return res, ref


def accumulation_probe():
    """Tensor-core accumulation error vs the number of accumulate steps: integer-valued operands
    (exact in TF32, Zlo = 0) so every product is exact and only the FP32 accumulation rounds."""
    lib = _lib.load()
    ctx = _lib.default_context()
    rng = np.random.default_rng(0)
    nS = 128
    print('accumulation probe (S3-only, exact products):')
    for n3 in (8, 32, 128, 512, 2048):
        K = 32 * n3
        Z = rng.integers(1, 1024, size=(nS, K)).astype(np.float32)      # positive -> monotone growth
        exact = -4.0 * (Z.astype(np.float64) @ Z.astype(np.float64).T)
        zhi = _lib.DeviceArray(ctx, Z.shape, np.float32, Z)
        zlo = _lib.DeviceArray(ctx, Z.shape, np.float32, np.zeros_like(Z))
        Dd = _lib.DeviceArray(ctx, (nS, nS), np.float32)
        shp = _lib.ContractShape(nS=nS, n1_blocks=0, n3_blocks=n3, ldz=K)
        for chunk in (n3, 64, 16, 4):
            if chunk > n3:
                continue
            _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, 0, chunk, 1, None))
            ctx.sync()
            D = Dd.download().astype(np.float64)
            e = (D - exact) / np.abs(exact)
            print('   K=%6d chunk=%5d blocks: rel err mean %+.3e  max|.| %.3e' % (K, chunk, e.mean(), np.abs(e).max()))
        _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, 1, 0, 1, None))
        ctx.sync()
        e = (Dd.download().astype(np.float64) - exact) / np.abs(exact)
        print('   K=%6d SIMT fp64 checker     : rel err mean %+.3e  max|.| %.3e' % (K, e.mean(), np.abs(e).max()))
        for a in (zhi, zlo, Dd):
            a.free()


def contraction_timing(nS, N, reps=5):
    lib = _lib.load()
    ctx = _lib.default_context()
    shp = _lib.ContractShape()
    _lib.check(lib.mem_operand_shape(ctx.handle, N, C.byref(shp)))
    shp.nS = nS
    K = int(shp.ldz)
    rng = np.random.default_rng(1)
    Z = rng.standard_normal((nS, K)).astype(np.float32)
    hi = (Z.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = ((Z - hi).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    zhi = _lib.DeviceArray(ctx, Z.shape, np.float32, hi)
    zlo = _lib.DeviceArray(ctx, Z.shape, np.float32, lo)
    Dd = _lib.DeviceArray(ctx, (nS, nS), np.float32)
    for chunk in (16, 64):
        for _ in range(2):
            _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, 0, chunk, 0, None))
        ctx.sync()
        t0 = time.time()
        for _ in range(reps):
            _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, 0, chunk, 0, None))
        ctx.sync()
        dt = (time.time() - t0) / reps
        flops = 3 * 2.0 * K * (nS * nS / 2)     # executed TF32 flops (upper triangle, 3 passes), approx
        print('contraction nS=%d N=%d K=%d chunk=%d: %.3f ms  -> %.1f TF/s executed (approx), %.3f Gpairs/s' %
              (nS, N, K, chunk, dt * 1e3, flops / dt / 1e12, nS * nS / dt / 1e9))
    for a in (zhi, zlo, Dd):
        a.free()


if __name__ == '__main__':
    quick = 'quick' in sys.argv
    report_pd(40, 32, 0, 0.1, contraction=1, impl='tile')
    report_pd(40, 32, 0, 0.1, contraction=0, impl='tile')
    report_pd(37, 25, 1, 10.0, contraction=1, impl='tile')
    report_pd(37, 25, 1, 10.0, contraction=0, impl='tile')
    report_pd(150, 64, 2, 0.1, contraction=1)
    report_pd(150, 64, 2, 0.1, contraction=0)
    report_pd(150, 64, 2, 0.1, contraction=0, msk=True)
    report_pd(300, 128, 3, 10.0, contraction=0)
    report_pd(300, 128, 3, 10.0, contraction=0, chunk=4)
    report_pd(300, 128, 3, 10.0, contraction=0, chunk=64, split=1)
    accumulation_probe()
    if not quick:
        contraction_timing(1000, 128)
        contraction_timing(2000, 256)
