"""CPU model of k_knn_select (csrc/dm.cu): the same digit-by-digit radix selection, threshold ties in index order and
final (value, index) sort, written with the kernel's per-thread structure (256 'threads' as vector lanes), checked
against a plain lexsort.  Lets the selection logic be tested where there is no GPU (tests/test_host.py)."""
import numpy as np

T = 256


def ord_key(x):
    """order-preserving unsigned key of a float32 / float64 array (-0 folded onto +0)."""
    x = np.where(x == 0, 0.0, x).astype(x.dtype)
    if x.dtype == np.float32:
        u = x.view(np.uint32)
        return np.where(u >> np.uint32(31), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    u = x.view(np.uint64)
    return np.where(u >> np.uint64(63), ~u, u | np.uint64(1 << 63)).astype(np.uint64)


def ord_val(u):
    if u.dtype == np.uint32:
        v = np.where(u >> np.uint32(31), u ^ np.uint32(0x80000000), ~u).astype(np.uint32)
        return v.view(np.float32).astype(np.float64)
    v = np.where(u >> np.uint64(63), u ^ np.uint64(1 << 63), ~u).astype(np.uint64)
    return v.view(np.float64)


def select_row(row, i, k):
    """row: 1-D float32/float64; returns (idx[k], val[k]) like the kernel for point i."""
    nS = row.shape[0]
    r = row.copy()
    r[i] = -np.inf
    ukey = ord_key(r)
    U = ukey.dtype.type
    bits = ukey.dtype.itemsize * 8
    prefix, mask, rem = U(0), U(0), k
    tid = np.arange(T)
    lane, wrp = tid & 31, tid >> 5
    for shift in range(bits - 8, -1, -8):
        hist = np.zeros(T, dtype=np.int64)
        sel = (ukey & mask) == prefix
        np.add.at(hist, ((ukey[sel] >> U(shift)) & U(0xff)).astype(np.int64), 1)
        h = hist.copy()
        inc = h.copy()
        o = 1
        while o < 32:                                       # warp-inclusive scan by shuffles
            t = np.zeros(T, dtype=np.int64)
            t[o:] = inc[:-o]
            inc = np.where(lane >= o, inc + t, inc)
            o <<= 1
        wtot = inc[31::32].copy()
        inc = inc + np.array([wtot[:w].sum() for w in wrp])
        exc = inc - h
        hit = (h > 0) & (exc < rem) & (rem <= inc)
        assert hit.sum() == 1
        s_bin = int(np.nonzero(hit)[0][0])
        rem = int(rem - exc[s_bin])
        prefix = U(prefix | (U(s_bin) << U(shift)))
        mask = U(mask | (U(0xff) << U(shift)))
    thr = prefix
    n_less = k - rem
    P2 = 32
    while P2 < k:
        P2 <<= 1
    selk = np.full(P2, ~U(0), dtype=ukey.dtype)
    seli = np.full(P2, 0x7fffffff, dtype=np.int64)
    less = np.nonzero(ukey < thr)[0]
    assert less.shape[0] == n_less
    perm = np.random.default_rng(0).permutation(n_less)      # atomics hand out the slots in any order
    selk[:n_less] = ukey[less][perm]
    seli[:n_less] = less[perm]
    running = 0
    for c0 in range(0, nS, T):
        if running >= rem:
            break
        j = c0 + tid
        f = (j < nS) & (ukey[np.minimum(j, nS - 1)] == thr)
        cnt = np.array([f[w * 32:(w + 1) * 32].sum() for w in range(T // 32)])
        off = running + np.array([cnt[:w].sum() for w in wrp])
        rank = off + np.array([f[(t // 32) * 32:t].sum() for t in tid])
        for t in np.nonzero(f & (rank < rem))[0]:
            selk[n_less + rank[t]] = thr
            seli[n_less + rank[t]] = j[t]
        running += int(cnt.sum())
    order = np.lexsort((seli, selk))                          # the bitonic network sorts by (key, index)
    selk, seli = selk[order][:k], seli[order][:k]
    val = ord_val(selk)
    val[0] = 0.0
    return seli.astype(np.int32), val


def reference_row(row, i, k):
    r = row.astype(np.float64)
    r[i] = -np.inf
    order = np.lexsort((np.arange(r.shape[0]), r))[:k]
    val = r[order]
    val[0] = 0.0
    return order.astype(np.int32), val
