// align.cu — ingest/normalise (a2, a3) and the in-plane alignment (a7): two periodic cubic-B-spline
// resamplings per particle, each = separable recursive prefilter (rows, columns) + 4x4-tap gather.
// All HBM-bound streaming kernels: every pass reads and writes each image exactly once.
#include "common.cuh"
#include "spline.cuh"

namespace mem {

// ------------------------------------------------------------------------------------------------
// a2+a3 (getDistanceCTF...py:246-283): picture = raw^T (SPIDER) or raw; conjugates flipped upside down;
// (x - mean(b))/std(b) with b = x*(1-msk) over all N^2 pixels (population std).
// One CTA per particle.  Pass 1: moments in fp64 (rows over warps, float4 over lanes).  Pass 2: every
// warp transposes 32x32 tiles through its own shared-memory buffer (no block barriers).
// ------------------------------------------------------------------------------------------------
template <bool FULL>   // FULL: N is a multiple of 32 (no edge predicates, 32-bit offsets)
__global__ void __launch_bounds__(256) k_ingest(const float* __restrict__ raw, const uint8_t* __restrict__ flip,
                                                float* __restrict__ out, int N, int transposed) {
  __shared__ double red[16];
  __shared__ float tile[8][32][33];
  const int i = blockIdx.x;
  const float* src = raw + (size_t)i * N * N;
  float* dst = out + (size_t)i * N * N;
  const bool fl = flip[i] != 0;
  const float half = 0.5f * N, r2lim = half * half;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double s = 0, s2 = 0;
  // Pixels inside the disc contribute 0.  Per warp row the fp32 partial sums of <= N values are promoted to fp64 once.
  for (int a = warp; a < N; a += nw) {
    float rs = 0.0f, rs2 = 0.0f;
    const float* rowp = src + a * N;
    for (int b = lane; b < N; b += 32) {
      int r = transposed ? b : a;
      const int c = transposed ? a : b;
      if (fl) r = N - 1 - r;
      const float x = (float)r - half + 1.0f, y = (float)c - half;   // annularMask.py:24-30, centre (N/2-1, N/2)
      const float v = rowp[b];
      const float m = (x * x + y * y < r2lim) ? 0.0f : v;
      rs += m;
      rs2 = fmaf(m, m, rs2);
    }
    s += (double)rs;
    s2 += (double)rs2;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { red[warp] = s; red[8 + warp] = s2; }
  __syncthreads();
  s = 0; s2 = 0;
  for (int w = 0; w < nw; ++w) { s += red[w]; s2 += red[8 + w]; }
  const double n = (double)N * N;
  const double mean = s / n;
  const double var = s2 / n - mean * mean;
  const float fm = (float)mean;
  const float inv = (float)(1.0 / sqrt(var));
  const int nt = (N + 31) / 32;
  for (int t = warp; t < nt * nt; t += nw) {
    const int tr = (t / nt) * 32, tc = (t % nt) * 32;   // output tile origin (rows r', cols c)
    if (transposed) {
      float v[32];
      const int rp0 = tr + lane;
      const float* sp = src + tc * N + (fl ? N - 1 - rp0 : rp0);   // raw index = c*N + r : lanes run along r
#pragma unroll
      for (int k = 0; k < 32; ++k)
        v[k] = (FULL || (tc + k < N && rp0 < N)) ? sp[k * N] : 0.0f;
#pragma unroll
      for (int k = 0; k < 32; ++k) tile[warp][k][lane] = v[k];
      __syncwarp();
      float* dp = dst + tr * N + tc + lane;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (FULL || (tr + k < N && tc + lane < N)) dp[k * N] = (tile[warp][lane][k] - fm) * inv;
      __syncwarp();
    } else {
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const int rp = tr + k, c = tc + lane;
        if (rp < N && c < N) {
          const int r = fl ? N - 1 - rp : rp;
          dst[(size_t)rp * N + c] = (src[(size_t)r * N + c] - fm) * inv;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cubic B-spline prefilter with periodic boundary — what scipy.ndimage.rotate's spline_filter amounts to
// on the 3x3-tiled image of rotatefill.py:21-25 away from the tile border (SURVEY §7 hard part 2):
//   c+[k] = 6 s[k] + z c+[k-1],   c[k] = z (c[k+1] - c+[k]),   z = sqrt(3) - 2.
// Both recursions are linear, so a line is cut into segments of E samples: each segment runs the
// recursion from a zero carry, and the true carry is  sum_h z^(samples between) * (local end of the
// h-th previous segment) — z^20 < 4e-12, so a few hops suffice, wrapping around for periodicity.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float zpow(int n) {
  float r = 1.0f, b = SPL_Z;
  while (n > 0) {
    if (n & 1) r *= b;
    b *= b;
    n >>= 1;
  }
  return r;
}

// cubic B-spline weights at fractional offset t in [0,1): taps at -1, 0, +1, +2
__device__ __forceinline__ void bspline_w(float t, float (&w)[4]) {
  const float u = 1.0f - t;
  const float t2 = t * t, u2 = u * u;
  w[0] = u2 * u * (1.0f / 6.0f);
  w[3] = t2 * t * (1.0f / 6.0f);
  w[1] = fmaf(t2, fmaf(0.5f, t, -1.0f), 2.0f / 3.0f);
  w[2] = fmaf(u2, fmaf(0.5f, u, -1.0f), 2.0f / 3.0f);
}

// Segment geometry of one filtered line, computed once on the host.
struct SegGeom {
  int E;        // samples per segment
  int used;     // segments that own samples
  int H;        // hops of the carry sum (<= used)
  float zE;     // z^E
  float zLast;  // z^(samples of the last segment)
};

// rows: one warp per image row, lane l owns samples [l*E, l*E+E).  Global traffic is fully coalesced
// (lane l touches sample k*32+l); the regrouping to "E consecutive samples per lane" goes through a
// per-warp shared-memory line with a +1 skew every 32 floats (conflict-free both ways for E = 2^m).
// Carries travel through warp shuffles.  grid = (ceil(N/8), nS).
template <int EMAX>
__global__ void __launch_bounds__(256) k_prefilter_rows(const float* __restrict__ in, float* __restrict__ out, int N,
                                                        int L, SegGeom g, int apply_mask) {
  __shared__ float line[8][EMAX * 32 + EMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= N) return;
  const size_t off = ((size_t)blockIdx.y * N + r) * N;
  const float* src = in + off;
  float* dst = out + off;
  float* ln = line[warp];
  const float half = 0.5f * N, r2lim = half * half;
  const float xm = (float)r - half + 1.0f;
  const float xm2 = xm * xm;
#pragma unroll
  for (int k = 0; k < EMAX; ++k) {
    const int c = k * 32 + lane;
    if (c < L) {
      float t = src[c < N ? c : L - c];              // L == N: periodic line; L == 2N-2: mirror extension
      if (apply_mask) {                              // img * msk, :325
        const float y = (float)c - half;
        if (!(xm2 + y * y < r2lim)) t = 0.0f;
      }
      ln[c + k] = 6.0f * t;                          // c + (c >> 5)
    }
  }
  __syncwarp();
  const int E = g.E, used = g.used;
  const int c0 = lane * E;
  const int n = max(0, min(E, L - c0));              // samples of this lane
  float v[EMAX];
#pragma unroll
  for (int j = 0; j < EMAX; ++j) v[j] = (j < n) ? ln[c0 + j + ((c0 + j) >> 5)] : 0.0f;
  // causal pass, zero carry
  float run = 0.0f;
#pragma unroll
  for (int j = 0; j < EMAX; ++j)
    if (j < n) { run = fmaf(SPL_Z, run, v[j]); v[j] = run; }
  // carry from the previous lanes (periodic)
  float carry = 0.0f, f = 1.0f;
  for (int h = 1; h <= g.H; ++h) {
    int srcl = lane - h;
    if (srcl < 0) srcl += used;
    const float e = __shfl_sync(0xffffffffu, run, srcl & 31);
    carry = fmaf(f, e, carry);
    f *= (srcl == used - 1) ? g.zLast : g.zE;
  }
  float zp = SPL_Z * carry;
#pragma unroll
  for (int j = 0; j < EMAX; ++j)
    if (j < n) { v[j] += zp; zp *= SPL_Z; }
  // anti-causal pass on u[k] = -z c+[k], zero carry
  run = 0.0f;
#pragma unroll
  for (int j = EMAX - 1; j >= 0; --j)
    if (j < n) { run = SPL_Z * (run - v[j]); v[j] = run; }
  carry = 0.0f; f = 1.0f;
  for (int h = 1; h <= g.H; ++h) {
    int srcl = lane + h;
    if (srcl >= used) srcl -= used;
    const float e = __shfl_sync(0xffffffffu, run, srcl & 31);
    carry = fmaf(f, e, carry);
    f *= (srcl == used - 1) ? g.zLast : g.zE;
  }
  zp = SPL_Z * carry;                                // restarts at the LAST sample of the lane
#pragma unroll
  for (int j = EMAX - 1; j >= 0; --j)
    if (j < n) { v[j] += zp; zp *= SPL_Z; }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < EMAX; ++j)
    if (j < n) ln[c0 + j + ((c0 + j) >> 5)] = v[j];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < EMAX; ++k) {
    const int c = k * 32 + lane;
    if (c < N) dst[c] = ln[c + k];
  }
}

// columns: CTA = 32 columns x S row-segments of up to 16 rows; a thread keeps its segment of one column
// in registers (loads coalesced across the 32 columns), carries go through shared memory. In place.
template <int MAXT, int MINB, int COL_E>
__global__ void __launch_bounds__(MAXT, MINB) k_prefilter_cols(float* __restrict__ data, int N, int L, SegGeom g) {
  __shared__ float ends[32][33];
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * 32 + lane;
  const int seg = threadIdx.x >> 5;
  const int E = g.E, used = g.used;                  // used == blockDim.x / 32
  const int r0 = seg * E;
  float* base = data + (size_t)blockIdx.y * N * N + col;
  const int n = (col < N) ? max(0, min(E, L - r0)) : 0;
  float v[COL_E];
#pragma unroll
  for (int j = 0; j < COL_E; ++j) {
    const int rr = r0 + j;
    v[j] = (j < n) ? 6.0f * base[(size_t)(rr < N ? rr : L - rr) * N] : 0.0f;
  }
  if (L != N) __syncthreads();                       // mirror rows are read by two segments before any store
  float run = 0.0f;
#pragma unroll
  for (int j = 0; j < COL_E; ++j)
    if (j < n) { run = fmaf(SPL_Z, run, v[j]); v[j] = run; }
  ends[seg][lane] = run;
  __syncthreads();
  float carry = 0.0f, f = 1.0f;
  for (int h = 1; h <= g.H; ++h) {
    int s = seg - h;
    if (s < 0) s += used;
    carry = fmaf(f, ends[s][lane], carry);
    f *= (s == used - 1) ? g.zLast : g.zE;
  }
  float zp = SPL_Z * carry;
#pragma unroll
  for (int j = 0; j < COL_E; ++j)
    if (j < n) { v[j] += zp; zp *= SPL_Z; }
  run = 0.0f;
#pragma unroll
  for (int j = COL_E - 1; j >= 0; --j)
    if (j < n) { run = SPL_Z * (run - v[j]); v[j] = run; }
  __syncthreads();
  ends[seg][lane] = run;
  __syncthreads();
  carry = 0.0f; f = 1.0f;
  for (int h = 1; h <= g.H; ++h) {
    int s = seg + h;
    if (s >= used) s -= used;
    carry = fmaf(f, ends[s][lane], carry);
    f *= (s == used - 1) ? g.zLast : g.zE;
  }
  zp = SPL_Z * carry;
#pragma unroll
  for (int j = COL_E - 1; j >= 0; --j)
    if (j < n) { v[j] += zp; zp *= SPL_Z; }
#pragma unroll
  for (int j = 0; j < COL_E; ++j)
    if (j < n && r0 + j < N) base[(size_t)(r0 + j) * N] = v[j];
}

// Exact-fit specialisations for the common box sizes (periodic boundary): N == 32*E for the row pass and
// N == 16*S for the column pass, so every lane / segment owns a full segment, all guards fold away and the hop
// count is a compile-time constant (the generic kernels above are issue-bound on their predication).
template <int E, bool MASK>
__global__ void __launch_bounds__(256) k_prefilter_rows_x(const float* __restrict__ in, float* __restrict__ out, float zE) {
  constexpr int N = 32 * E;
  __shared__ float line[8][N + E];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  const size_t off = ((size_t)blockIdx.y * N + r) * N;
  const float* src = in + off;
  float* dst = out + off;
  float* ln = line[warp];
  constexpr float half = 0.5f * N, r2lim = half * half;
  const float xm = (float)r - half + 1.0f;
  const float lim = r2lim - xm * xm;                  // keep the pixel iff y*y < lim
#pragma unroll
  for (int k = 0; k < E; ++k) {
    const int c = k * 32 + lane;
    float t = src[c];
    if (MASK) {                                        // img * msk, :325
      const float y = (float)c - half;
      if (!(y * y < lim)) t = 0.0f;
    }
    ln[c + k] = 6.0f * t;
  }
  __syncwarp();
  const int c0 = lane * E;
  float v[E];
#pragma unroll
  for (int j = 0; j < E; ++j) v[j] = ln[c0 + j + ((c0 + j) >> 5)];
  spline_line_warp<E>(v, zE, lane);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < E; ++j) ln[c0 + j + ((c0 + j) >> 5)] = v[j];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < E; ++k) dst[k * 32 + lane] = ln[k * 32 + lane + k];
}

template <int S, int MB = (S == 16 ? 3 : (S < 16 ? 2 : 1))>   // N = 256: 40 registers -> three 512-thread CTAs per SM (176 -> 151 us per pass)
__global__ void __launch_bounds__(32 * S, MB) k_prefilter_cols_x(float* __restrict__ data, float zE) {
  constexpr int N = 16 * S, E = 16, H = SPL_REACH / E + 2;
  __shared__ float ends[S][33];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  float* base = data + ((size_t)blockIdx.y * N + seg * E) * N + blockIdx.x * 32 + lane;
  float v[E];
#pragma unroll
  for (int j = 0; j < E; ++j) v[j] = 6.0f * base[(size_t)j * N];
  float run = 0.0f;
#pragma unroll
  for (int j = 0; j < E; ++j) { run = fmaf(SPL_Z, run, v[j]); v[j] = run; }
  ends[seg][lane] = run;
  __syncthreads();
  float carry = 0.0f, f = 1.0f;
#pragma unroll
  for (int h = 1; h <= H; ++h) {
    carry = fmaf(f, ends[(seg - h + S) % S][lane], carry);
    f *= zE;
  }
  float zp = SPL_Z * carry;
#pragma unroll
  for (int j = 0; j < E; ++j) { v[j] += zp; zp *= SPL_Z; }
  run = 0.0f;
#pragma unroll
  for (int j = E - 1; j >= 0; --j) { run = SPL_Z * (run - v[j]); v[j] = run; }
  __syncthreads();
  ends[seg][lane] = run;
  __syncthreads();
  carry = 0.0f; f = 1.0f;
#pragma unroll
  for (int h = 1; h <= H; ++h) {
    carry = fmaf(f, ends[(seg + h) % S][lane], carry);
    f *= zE;
  }
  zp = SPL_Z * carry;
#pragma unroll
  for (int j = E - 1; j >= 0; --j) { v[j] += zp; zp *= SPL_Z; }
#pragma unroll
  for (int j = 0; j < E; ++j) base[(size_t)j * N] = v[j];
}

// a2, RELION branch (:263-264): scipy.ndimage.shift(img, (s0, s1), order=3, mode='wrap') on mirror-prefiltered
// coefficients.  SciPy semantics pinned in SURVEY §7(2) and tests/test_host.py: input coordinate o - s is
// wrapped with period N-1, the 4 support indices are mirrored (k<0 -> -k, k>N-1 -> 2(N-1)-k).
__device__ __forceinline__ void shift_taps(int o, double s, int N, int (&k)[4], float (&w)[4]) {
  double x = (double)o - s;
  const double Lp = (double)(N - 1);
  if (x < 0.0) x += Lp * (double)((int)(-x / Lp) + 1);
  else if (x > Lp) x -= Lp * (double)((int)(x / Lp));
  const double fl = floor(x);
  const int i0 = (int)fl;
  bspline_w((float)(x - fl), w);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int kk = i0 - 1 + a;
    if (kk < 0) kk = -kk;
    if (kk > N - 1) kk = 2 * (N - 1) - kk;
    k[a] = kk;
  }
}
__global__ void __launch_bounds__(256) k_shift(const float* __restrict__ coef, float* __restrict__ out,
                                               const double* __restrict__ shift, int N) {
  const int img = blockIdx.z;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (r >= N || c >= N) return;
  int ka[4], kb[4];
  float wa[4], wb[4];
  shift_taps(r, shift[2 * img], N, ka, wa);
  shift_taps(c, shift[2 * img + 1], N, kb, wb);
  const float* src = coef + (size_t)img * N * N;
  float acc = 0.0f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* rowp = src + (size_t)ka[a] * N;
    acc = fmaf(wa[a], wb[0] * __ldg(rowp + kb[0]) + wb[1] * __ldg(rowp + kb[1]) + wb[2] * __ldg(rowp + kb[2]) +
                          wb[3] * __ldg(rowp + kb[3]), acc);
  }
  out[(size_t)img * N * N + (size_t)r * N + c] = acc;
}

// per-image rotation cos/sin in fp64 (ndimage.rotate: matrix [[c, s], [-s, c]], angle in degrees);
// entry nS holds the common second rotation by -psi_p (:330)
__global__ void k_angles(const double* __restrict__ psi_deg, double psi_p_deg, double2* __restrict__ cs, int nS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nS) return;
  const double a = (i < nS ? psi_deg[i] : -psi_p_deg) * 0.017453292519943295769;
  double s, c;
  sincos(a, &s, &c);
  cs[i] = make_double2(c, s);
}

// a7 (rotatefill.py:21-41): out(o) = sum_{4x4} w * coef[(floor(x)-1+a) mod N], x = R (o - ctr) + ctr, ctr=(N-1)/2.
// CTA = 32x32 output tile; the (periodically wrapped) bounding box of its source footprint, at most
// 50x50 coefficients, is staged in shared memory.  Coordinates in fp64, weights in fp32.
// If msk2 != NULL a second, masked copy is written (img*msk2, :344).  FULL: N is a multiple of 32.
// Bank = (i*P + j) mod 32 with lanes stepping by (sin, cos) per output column: P = 65 gives bank = i + j, which
// advances by |sin + cos| >= 1 per lane when sin*cos >= 0; P = 63 gives j - i for the other two quadrants.
constexpr int ROT_T = 32, ROT_B = 50, ROT_PMAX = 65;
template <bool FULL, int ROT_P>   // the pitch is a compile-time constant so the 16 taps use immediate offsets
__device__ __forceinline__ void rotate_tile(float* __restrict__ tile, const double2 a, const float* __restrict__ coef,
                                            float* __restrict__ out, int N, const uint8_t* __restrict__ msk2,
                                            float* __restrict__ out_masked) {
  const int img = blockIdx.z;
  const double ctr = 0.5 * (N - 1);
  const int r0 = blockIdx.y * ROT_T, c0 = blockIdx.x * ROT_T;
  // source coordinates of the tile corners -> bounding box origin (floor(min) - 1); every thread computes it
  const double tr = r0 - ctr, tc = c0 - ctr, ext = ROT_T - 1;
  const double b0 = a.x * tr + a.y * tc + ctr + fmin(a.x * ext, 0.0) + fmin(a.y * ext, 0.0);
  const double b1 = -a.y * tr + a.x * tc + ctr + fmin(-a.y * ext, 0.0) + fmin(a.x * ext, 0.0);
  const int o0 = __double2int_rd(b0) - 1, o1 = __double2int_rd(b1) - 1;
  const float* src = coef + (size_t)img * N * N;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  {
    // Periodic wrap of the box origin.  Source coordinates stay within ctr +- N/sqrt(2), so -N < o < 2N, and every
    // index below exceeds N by less than 2N for N >= 16 (the smallest box the prefilter takes): conditional
    // adds instead of modulo arithmetic, and two running row pointers instead of one 64-bit address per load.
    int w0m = o0, w1m = o1;
    w0m += (w0m < 0) ? N : 0;
    w0m -= (w0m >= N) ? N : 0;
    w1m += (w1m < 0) ? N : 0;
    w1m -= (w1m >= N) ? N : 0;
    int gx0 = w1m + lane;
    gx0 -= (gx0 >= N) ? N : 0;
    gx0 -= (gx0 >= N) ? N : 0;
    int gx1 = gx0 + 32;
    gx1 -= (gx1 >= N) ? N : 0;
    gx1 -= (gx1 >= N) ? N : 0;
    int gy = w0m + wrp;
    gy -= (gy >= N) ? N : 0;
    const float* p0 = src + (gy * N + gx0);
    const float* p1 = src + (gy * N + gx1);
    const int fwd = 8 * N, back = 8 * N - N * N;
    float* tp = tile + wrp * ROT_P + lane;
    const bool second = lane + 32 < ROT_B;
#pragma unroll
    for (int it = 0; it < (ROT_B + 7) / 8; ++it) {
      if (it * 8 + wrp < ROT_B) {                      // warp-uniform
        tp[it * 8 * ROT_P] = __ldg(p0);
        if (second) tp[it * 8 * ROT_P + 32] = __ldg(p1);
      }
      gy += 8;
      const bool wrap = gy >= N;
      gy -= wrap ? N : 0;
      const int adv = wrap ? back : fwd;
      p0 += adv;
      p1 += adv;
    }
  }
  __syncthreads();
  const int c = c0 + lane;
  const double xb0 = a.x * ((r0 + wrp) - ctr) + a.y * (c - ctr) + ctr - (double)o0;   // row coordinate, bbox-relative
  const double xb1 = -a.y * ((r0 + wrp) - ctr) + a.x * (c - ctr) + ctr - (double)o1;  // column coordinate
  // 6.26 unsigned fixed point from here on (coordinates are in [1, 48)): high bits = tap origin, low 26 bits =
  // fraction (1.5e-8 pixel, below the fp32 weights' resolution).  The 4 outputs of a thread are 8 rows apart: one
  // 32-bit add per axis (modulo 2^32, so negative steps need no special case) instead of fp64 floor/convert chains.
  unsigned int x0 = (unsigned int)__double2ll_rn(xb0 * 67108864.0), x1 = (unsigned int)__double2ll_rn(xb1 * 67108864.0);
  const unsigned int DX0 = (unsigned int)__double2ll_rn(8.0 * a.x * 67108864.0);
  const unsigned int DX1 = (unsigned int)__double2ll_rn(-8.0 * a.y * 67108864.0);
  const int oidx = (r0 + wrp) * N + c;                  // pixel index inside the image (N*N < 2^31)
  float* dst = out + (size_t)img * N * N + oidx;
  float* dstm = out_masked ? out_masked + (size_t)img * N * N + oidx : nullptr;
  const uint8_t* mk = msk2 ? msk2 + oidx : nullptr;
  const int rstep = 8 * N;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (FULL || (r0 + wrp + 8 * q < N && c < N)) {
      const int i0 = (int)(x0 >> 26), j0 = (int)(x1 >> 26);
      const float t0 = (float)(x0 & 0x3ffffffu) * 1.4901161193847656e-08f;
      const float t1 = (float)(x1 & 0x3ffffffu) * 1.4901161193847656e-08f;
      float wa[4], wb[4];
      bspline_w(t0, wa);
      bspline_w(t1, wb);
      const float* p = tile + ((i0 - 1) * ROT_P + (j0 - 1));   // 0 <= i0-1, j0-1 and i0+2, j0+2 < ROT_B
      float acc = 0.0f;
#pragma unroll
      for (int ai = 0; ai < 4; ++ai) {
        float rs = wb[0] * p[0];
        rs = fmaf(wb[1], p[1], rs);
        rs = fmaf(wb[2], p[2], rs);
        rs = fmaf(wb[3], p[3], rs);
        acc = fmaf(wa[ai], rs, acc);
        p += ROT_P;
      }
      dst[q * rstep] = acc;
      if (dstm) dstm[q * rstep] = mk[q * rstep] ? acc : 0.0f;
    }
    x0 += DX0;
    x1 += DX1;
  }
}

template <bool FULL>
__global__ void __launch_bounds__(256) k_rotate(const float* __restrict__ coef, float* __restrict__ out,
                                                const double2* __restrict__ cs, int cs_stride, int N,
                                                const uint8_t* __restrict__ msk2, float* __restrict__ out_masked) {
  __shared__ float tile[ROT_B * ROT_PMAX];
  const double2 a = cs[(size_t)blockIdx.z * cs_stride];
  if (a.x * a.y >= 0.0) rotate_tile<FULL, 65>(tile, a, coef, out, N, msk2, out_masked);
  else rotate_tile<FULL, 63>(tile, a, coef, out, N, msk2, out_masked);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int ingest_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float* out, int nS, int N, int transposed,
               cudaStream_t st) {
  if (N % 32 == 0) MEM_LAUNCH(ctx, k_ingest<true>, nS, 256, 0, st, raw, flip, out, N, transposed);
  else MEM_LAUNCH(ctx, k_ingest<false>, nS, 256, 0, st, raw, flip, out, N, transposed);
  return 0;
}

static SegGeom make_geom(int L, int E) {
  const int N = L;
  SegGeom g;
  g.E = E;
  g.used = (N + E - 1) / E;
  g.H = SPL_REACH / E + 2;
  if (g.H > g.used) g.H = g.used;
  const double z = sqrt(3.0) - 2.0;
  g.zE = (float)pow(z, (double)E);
  g.zLast = (float)pow(z, (double)(N - (g.used - 1) * E));
  return g;
}

// mirror == 0: periodic boundary (rotatefill); mirror == 1: whole-sample mirror boundary, realised as the
// periodic filter on the (2N-2)-long mirror extension of every line (ndimage.shift's spline_filter).
// rows_done != 0: `out` already holds the row-filtered lines (fused upstream, lowpass.cu); only the column pass runs.
static int prefilter_run(mem_ctx* ctx, const float* in, float* out, int nS, int N, int apply_mask, int mirror,
                         cudaStream_t st, int rows_done = 0) {
  const int L = mirror ? 2 * N - 2 : N;
  if (L > 1024 || N > 512 || N < 16) {
    set_error("box size %d is not supported by the spline prefilter (16 <= N <= 512)", N);
    return 1;
  }
  const int E = (L + 31) / 32;
  const SegGeom gr = make_geom(L, E);
  const dim3 grid((N + 7) / 8, nS);
  const bool rows_exact = !mirror && (N == 64 || N == 128 || N == 256 || N == 320);
  if (rows_done) {
    // nothing to do for the rows
  } else if (rows_exact) {
    if (N == 64) {
      if (apply_mask) MEM_LAUNCH(ctx, (k_prefilter_rows_x<2, true>), grid, 256, 0, st, in, out, gr.zE);
      else MEM_LAUNCH(ctx, (k_prefilter_rows_x<2, false>), grid, 256, 0, st, in, out, gr.zE);
    } else if (N == 128) {
      if (apply_mask) MEM_LAUNCH(ctx, (k_prefilter_rows_x<4, true>), grid, 256, 0, st, in, out, gr.zE);
      else MEM_LAUNCH(ctx, (k_prefilter_rows_x<4, false>), grid, 256, 0, st, in, out, gr.zE);
    } else if (N == 256) {
      if (apply_mask) MEM_LAUNCH(ctx, (k_prefilter_rows_x<8, true>), grid, 256, 0, st, in, out, gr.zE);
      else MEM_LAUNCH(ctx, (k_prefilter_rows_x<8, false>), grid, 256, 0, st, in, out, gr.zE);
    } else {                                         // 320 = 32 x 10: BASELINE config 5
      if (apply_mask) MEM_LAUNCH(ctx, (k_prefilter_rows_x<10, true>), grid, 256, 0, st, in, out, gr.zE);
      else MEM_LAUNCH(ctx, (k_prefilter_rows_x<10, false>), grid, 256, 0, st, in, out, gr.zE);
    }
  } else if (E <= 4) MEM_LAUNCH(ctx, k_prefilter_rows<4>, grid, 256, 0, st, in, out, N, L, gr, apply_mask);
  else if (E <= 8) MEM_LAUNCH(ctx, k_prefilter_rows<8>, grid, 256, 0, st, in, out, N, L, gr, apply_mask);
  else if (E <= 16) MEM_LAUNCH(ctx, k_prefilter_rows<16>, grid, 256, 0, st, in, out, N, L, gr, apply_mask);
  else MEM_LAUNCH(ctx, k_prefilter_rows<32>, grid, 256, 0, st, in, out, N, L, gr, apply_mask);
  // columns: segments of <= 16 rows (<= 32 rows for lines longer than 512), at most 32 segments per column
  const int CE = L <= 512 ? 16 : 32;
  const int S = (L + CE - 1) / CE;
  const int Ec = (L + S - 1) / S;
  const SegGeom gc = make_geom(L, Ec);
  const dim3 gcols((N + 31) / 32, nS);
  if (!mirror && (N == 128 || N == 256 || N == 320)) {
    if (N == 128) MEM_LAUNCH(ctx, k_prefilter_cols_x<8>, gcols, 256, 0, st, out, gc.zE);
    else if (N == 256) MEM_LAUNCH(ctx, k_prefilter_cols_x<16>, gcols, 512, 0, st, out, gc.zE);
    else MEM_LAUNCH(ctx, k_prefilter_cols_x<20>, gcols, 640, 0, st, out, gc.zE);
    return 0;
  }
  auto kc_small = k_prefilter_cols<512, 2, 16>;
  auto kc_large = k_prefilter_cols<1024, 1, 16>;
  auto kc_long = k_prefilter_cols<1024, 1, 32>;
  if (CE == 32) MEM_LAUNCH(ctx, kc_long, gcols, 32 * gc.used, 0, st, out, N, L, gc);
  else if (gc.used <= 16) MEM_LAUNCH(ctx, kc_small, gcols, 32 * gc.used, 0, st, out, N, L, gc);
  else MEM_LAUNCH(ctx, kc_large, gcols, 32 * gc.used, 0, st, out, N, L, gc);
  return 0;
}

// RELION ingest: raw (picture orientation) -> cubic 'wrap' shift by shift[i] = (s0, s1) -> out.  tmp: scratch.
int shift_run(mem_ctx* ctx, const float* raw, const double* shift, float* tmp, float* out, int nS, int N,
              cudaStream_t st) {
  MEM_CHECK(prefilter_run(ctx, raw, tmp, nS, N, 0, 1, st));
  MEM_LAUNCH(ctx, k_shift, dim3((N + 31) / 32, (N + 7) / 8, nS), 256, 0, st, tmp, out, shift, N);
  return 0;
}

// B (low-passed images) -> imgAll (aligned), optional masked copy into B.  A, B: [nS][N][N] scratch.
// rows_done != 0: A already holds (img * msk) after the row pass of the first prefilter.
int align_run(mem_ctx* ctx, float* A, float* B, float* imgAll, const double* psi_deg, double psi_p_deg, double2* cs,
              const uint8_t* msk2, int nS, int N, cudaStream_t st, int rows_done) {
  if (rotate_fast_supported(N) && !ctx->legacy_rotate) {
    MEM_CHECK(ctx->rot_pid.ensure((size_t)nS));
    uint8_t* pid = ctx->rot_pid.as<uint8_t>();
    MEM_CHECK(rotate_angles_run(ctx, psi_deg, psi_p_deg, cs, pid, nS, st));
    MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 1, 0, st, rows_done));         // (img * msk) -> coefficients
    MEM_CHECK(rotate_img_run(ctx, A, B, cs, pid, nS, N, st));
    MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 0, 0, st));
    MEM_CHECK(rotate_common_run(ctx, A, imgAll, cs + nS, -psi_p_deg, nS, N, msk2, msk2 ? B : (float*)nullptr, st));
    return 0;
  }
  MEM_LAUNCH(ctx, k_angles, (nS + 1 + 127) / 128, 128, 0, st, psi_deg, psi_p_deg, cs, nS);
  const dim3 grot((N + ROT_T - 1) / ROT_T, (N + ROT_T - 1) / ROT_T, nS);
  const bool full = (N % ROT_T) == 0;
  MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 1, 0, st, rows_done));         // (img * msk) -> coefficients
  if (full) MEM_LAUNCH(ctx, k_rotate<true>, grot, 256, 0, st, A, B, cs, 1, N, (const uint8_t*)nullptr, (float*)nullptr);
  else MEM_LAUNCH(ctx, k_rotate<false>, grot, 256, 0, st, A, B, cs, 1, N, (const uint8_t*)nullptr, (float*)nullptr);
  MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 0, 0, st));
  if (full) MEM_LAUNCH(ctx, k_rotate<true>, grot, 256, 0, st, A, imgAll, cs + nS, 0, N, msk2, msk2 ? B : (float*)nullptr);
  else MEM_LAUNCH(ctx, k_rotate<false>, grot, 256, 0, st, A, imgAll, cs + nS, 0, N, msk2, msk2 ? B : (float*)nullptr);
  return 0;
}

// Batched PDs (pd_distance_batch_device): the same chain over the concatenated stack; cs2 / pid2 [nS] carry the second
// rotation's angle per image (its PD's -psi_p).  Boxes that are a multiple of 32 only; no msk2.
int align_batch_run(mem_ctx* ctx, float* A, float* B, float* imgAll, const double* psi_deg, double2* cs, const double2* cs2,
                    const uint8_t* pid2, int nS, int N, cudaStream_t st, int rows_done) {
  MEM_CHECK(ctx->rot_pid.ensure((size_t)nS));
  uint8_t* pid = ctx->rot_pid.as<uint8_t>();
  MEM_CHECK(rotate_angles_run(ctx, psi_deg, 0.0, cs, pid, nS, st));
  MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 1, 0, st, rows_done));
  MEM_CHECK(rotate_img_run(ctx, A, B, cs, pid, nS, N, st));
  MEM_CHECK(prefilter_run(ctx, B, A, nS, N, 0, 0, st));
  MEM_CHECK(rotate_img_run(ctx, A, imgAll, cs2, pid2, nS, N, st));
  return 0;
}

}  // namespace mem
