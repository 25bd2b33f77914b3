"""Pinned-memory flavours for the H2D leg of the e2e path: default, write-combined, portable|mapped — all 55.6 GB/s on the round-2 box
(the link, not the host-memory flavour, is the limit).   python scripts/h2d_pinned_flags.py"""
import ctypes as C, time, numpy as np
rt = C.CDLL('libcudart.so.12')
def chk(r):
    assert r == 0, r
n = 512 << 20
d = C.c_void_p(); chk(rt.cudaMalloc(C.byref(d), C.c_size_t(n)))
st = C.c_void_p(); chk(rt.cudaStreamCreate(C.byref(st)))
for name, flags in (('default pinned', 0), ('write-combined', 4), ('portable|mapped', 3)):
    h = C.c_void_p(); chk(rt.cudaHostAlloc(C.byref(h), C.c_size_t(n), C.c_uint(flags)))
    C.memset(h, 1, n)
    chk(rt.cudaMemcpyAsync(d, h, C.c_size_t(n), 1, st)); chk(rt.cudaStreamSynchronize(st))
    best = 0
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(8):
            chk(rt.cudaMemcpyAsync(d, h, C.c_size_t(n), 1, st))
        chk(rt.cudaStreamSynchronize(st))
        best = max(best, 8 * n / (time.perf_counter() - t0) / 1e9)
    print('%-18s H2D %.1f GB/s' % (name, best))
    chk(rt.cudaFreeHost(h))
