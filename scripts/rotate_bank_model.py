"""Shared-memory wavefront model of the 16-tap gather of k_rotate (align.cu), to size the next experiment.

A warp computes 32 consecutive pixels of an output row: lane l samples the source at (x0 + l sin t, x1 + l cos t), taps
at rows i0-1..i0+2, columns j0-1..j0+2 of the staged tile.  32 banks of 4 bytes; a load costs as many wavefronts as the
largest number of distinct words it needs from one bank (64-bit loads: per half-warp, the two halves add up).

  current  : tile[i][j] float, pitch 65 (sin*cos >= 0) or 63, 16 LDS.32 per pixel
  pairs    : tile[i][j] = (c[i][j], c[i][j+1]) float2, pitch 65 / 63 float2, 8 LDS.64 per pixel (any column parity)
  pairs-pad: the same with pitch 66 / 62 float2 (bank pair advances by 2 per row)

    python scripts/rotate_bank_model.py [--explore]      --explore: warp = W x H pixel patch, best 1..4 compile-time pitches
"""
import numpy as np


def wavefronts_32(addr_words):
    banks = addr_words % 32
    worst = 0
    for b in np.unique(banks):
        worst = max(worst, len(np.unique(addr_words[banks == b])))
    return worst


def wavefronts_64(addr_pairs):
    tot = 0
    for half in (addr_pairs[:16], addr_pairs[16:]):
        bp = half % 16                                   # bank pair of an 8-byte word
        worst = 0
        for b in np.unique(bp):
            worst = max(worst, len(np.unique(half[bp == b])))
        tot += worst
    return tot


def model(pitch_pos, pitch_neg, pairs, n_angles=720, n_off=6, seed=0):
    rng = np.random.default_rng(seed)
    lanes = np.arange(32)
    per_angle = []
    for t in np.linspace(0, 2 * np.pi, n_angles, endpoint=False):
        s, c = np.sin(t), np.cos(t)
        P = pitch_pos if s * c >= 0 else pitch_neg
        acc = 0.0
        for _ in range(n_off):
            x0, x1 = rng.random(2) + 40.0
            i0 = np.floor(x0 + lanes * s).astype(int)
            j0 = np.floor(x1 + lanes * c).astype(int)
            w = 0
            for a in range(4):
                if pairs:
                    for b in (0, 2):
                        w += wavefronts_64((i0 - 1 + a) * P + (j0 - 1 + b))
                else:
                    for b in range(4):
                        w += wavefronts_32((i0 - 1 + a) * P + (j0 - 1 + b))
            acc += w
        per_angle.append(acc / n_off)
    return np.array(per_angle)


def explore():
    """Warp as a W x H patch of output pixels (lane = lc + W lr), tile pitch from a small compile-time set chosen per
    image by its angle: mean wavefronts per warp-pixel for the best 1..4 pitches in 50..99."""
    rng = np.random.default_rng(0)
    angles = np.linspace(0, 2 * np.pi, 180, endpoint=False)
    lane = np.arange(32)
    for W, H in ((32, 1), (16, 2), (8, 4), (4, 8)):
        lc, lr = lane % W, lane // W
        table = np.zeros((len(angles), 50))
        for ia, t in enumerate(angles):
            s, c = np.sin(t), np.cos(t)
            offs = rng.random((4, 2)) + 60.0
            for ip, P in enumerate(range(50, 100)):
                acc = 0
                for x0, x1 in offs:
                    i0 = np.floor(x0 + lc * s + lr * c).astype(int)
                    j0 = np.floor(x1 + lc * c - lr * s).astype(int)
                    acc += sum(wavefronts_32((i0 - 1 + a) * P + (j0 - 1 + b)) for a in range(4) for b in range(4))
                table[ia, ip] = acc / len(offs)
        chosen = []
        for k in range(4):
            m, P = min((table[:, chosen + [P]].min(1).mean(), P) for P in range(50))
            chosen.append(P)
            print('patch %2d x %d, %d pitch(es) %-18s -> %.1f wavefronts per warp-pixel (%.2f per load)'
                  % (W, H, k + 1, [50 + q for q in chosen], m, m / 16))
        print('patch %2d x %d, best pitch per angle          -> %.1f' % (W, H, table.min(1).mean()))


if __name__ == '__main__':
    import sys
    if '--explore' in sys.argv:
        explore()
        sys.exit(0)
    cur = model(65, 63, False)
    print('current  (16 LDS.32): %.1f wavefronts per warp-pixel (%.2f per load), worst angle %.1f'
          % (cur.mean(), cur.mean() / 16, cur.max()))
    for name, pp, pn in (('pairs    ', 65, 63), ('pairs-pad', 66, 62), ('pairs-p67', 67, 61)):
        m = model(pp, pn, True)
        print('%s ( 8 LDS.64): %.1f wavefronts per warp-pixel (%.2f per load), worst angle %.1f  -> %.0f %% of current'
              % (name, m.mean(), m.mean() / 8, m.max(), 100 * m.mean() / cur.mean()))
