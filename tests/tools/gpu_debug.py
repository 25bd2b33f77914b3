"""GPU bring-up report: per-field errors of the CUDA path against the oracle, SIMT vs tcgen05
contraction, chunk sweeps, quick timings.  Run on the GPU box:
    python tests/tools/gpu_debug.py [acc] [time]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from manifoldem_python_b200 import _lib, synthetic, pd_stage   # noqa: E402
from oracle import pd_distance as opd                            # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


_cache = {}


def report_pd(nS, N, seed, snr, contraction, chunk=0, split=0, msk=False, impl='periodic', fields=True):
    key = (nS, N, seed, snr, msk, impl)
    if key not in _cache:
        pd = synthetic.make_pd(nS, N, seed=seed, snr=snr)
        em = pd['em']
        msk2 = None
        if msk:
            yy, xx = np.mgrid[:N, :N]
            msk2 = ((yy - N / 2) ** 2 / (0.4 * N) ** 2 + (xx - N / 2) ** 2 / (0.3 * N) ** 2) < 1
        ref = opd.pd_distance(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                              em['EkV'], em['AmpContrast'], msk2=1 if msk2 is None else msk2, rotate_impl=impl)
        _cache[key] = (pd, msk2, ref)
    pd, msk2, ref = _cache[key]
    em = pd['em']
    t0 = time.time()
    res = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'],
                          em['EkV'], em['AmpContrast'], msk2=msk2, contraction=contraction, k_chunk_blocks=chunk,
                          split_k=split)
    t1 = time.time()
    D, Dr = res['D'], ref['D']
    off = ~np.eye(nS, dtype=bool)
    relD = np.abs(D - Dr)[off] / Dr[off]
    print('PD nS=%d N=%d snr=%g contraction=%d chunk=%d split=%d msk=%d: gpu %.2fs' %
          (nS, N, snr, contraction, chunk, split, msk, t1 - t0))
    print('   D offdiag rel err: max %.3e  p99 %.3e  median %.3e   diag abs/maxD %.3e   minD/maxD %.3f' %
          (relD.max(), np.quantile(relD, 0.99), np.median(relD), np.abs(np.diag(D) - np.diag(Dr)).max() / Dr.max(),
           Dr[off].min() / Dr.max()))
    if fields:
        print('   ' + '  '.join('%s %.2e' % (k, rel(res[k].reshape(ref[k].shape), ref[k]))
                                for k in ('imgAll', 'imgAllFlip', 'CTF', 'imgAvg', 'imgAvgFlip', 'imgAllIntensity')))
        print('   timings(ms):', {k: round(v, 3) for k, v in _lib.default_context().timings().items()})
    return res, ref


def contraction_timing(nS, N, reps=5, kinds=(0, 2), chunks=(1, 2, 3, 4, 8)):
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
    for kind in kinds:
        for chunk in chunks:
            for _ in range(2):
                _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, kind, chunk, 0, None))
            ctx.sync()
            t0 = time.time()
            for _ in range(reps):
                _lib.check(lib.mem_contract_device(ctx.handle, C.byref(shp), zhi.ptr, zlo.ptr, Dd.ptr, kind, chunk, 0, None))
            ctx.sync()
            dt = (time.time() - t0) / reps
            print('contraction kind=%d nS=%d N=%d K=%d chunk=%d: %.3f ms  %.3f Gpairs/s' %
                  (kind, nS, N, K, chunk, dt * 1e3, nS * nS / dt / 1e9))
    for a in (zhi, zlo, Dd):
        a.free()


if __name__ == '__main__':
    if 'chunk3' in sys.argv:      # promotion period per column group: code = c12 | c3 << 8
        codes = (2, 2 | 4 << 8, 2 | 8 << 8, 2 | 16 << 8, 1 | 8 << 8, 1 | 16 << 8, 8)
        for case in ((300, 128, 3, 10.0), (300, 128, 3, 0.1), (200, 256, 7, 10.0), (600, 64, 8, 100.0)):
            for code in codes:
                report_pd(*case, contraction=0, chunk=code, fields=False)
        contraction_timing(2000, 256, kinds=(0,), chunks=codes)
        sys.exit(0)
    if 'tc2' in sys.argv:
        report_pd(40, 32, 0, 0.1, contraction=2, impl='tile', fields=False)
        report_pd(40, 32, 0, 0.1, contraction=0, impl='tile', fields=False)
        report_pd(300, 128, 3, 10.0, contraction=2, fields=False)
        report_pd(300, 128, 3, 10.0, contraction=0, fields=False)
        report_pd(300, 128, 3, 10.0, contraction=0, fields=False, split=1)
        report_pd(300, 128, 3, 10.0, contraction=0, fields=False, split=3, chunk=2)
        contraction_timing(1000, 128)
        contraction_timing(2000, 256)
        sys.exit(0)
    report_pd(40, 32, 0, 0.1, contraction=1, impl='tile')
    report_pd(40, 32, 0, 0.1, contraction=0, impl='tile')
    report_pd(37, 25, 1, 10.0, contraction=0, impl='tile')
    report_pd(150, 64, 2, 0.1, contraction=0)
    report_pd(150, 64, 2, 0.1, contraction=0, msk=True)
    report_pd(100, 96, 4, 1.0, contraction=0)
    report_pd(60, 160, 5, 1.0, contraction=0)
    report_pd(300, 128, 3, 10.0, contraction=1)
    for ch in (1, 2, 4, 8):
        report_pd(300, 128, 3, 10.0, contraction=0, chunk=ch, fields=False)
    for ch in (1, 2, 4):
        report_pd(257, 64, 5, 10.0, contraction=0, chunk=ch, fields=False)
    if 'time' in sys.argv:
        contraction_timing(1000, 128)
        contraction_timing(2000, 256)
