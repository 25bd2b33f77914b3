// lowpass.cu — ingest, a5 low-pass (getDistanceCTF_local_Conj9combinedS2.py:286-293),
//   img <- Re ifft2( fft2(img) * ifftshift(G) ),
// and the a10 forward transform (:344) with the library's own FFT kernels for N = 256 and N = 128.
// The transform is separable, the filter is not: the row passes are fused into their neighbours (k_ingest_rowfft*: ingest +
// moments + R2C rows; k_rowifft_prefilter*: C2R rows + annular mask + row pass of the first spline prefilter), and everything
// that happens along ky — forward FFT, normalisation, multiplication by G(ky,kx)/N^2, inverse FFT — is ONE kernel
// (k_colfilter*), so the half spectrum crosses HBM once (read + write) in the column pass instead of three times (2-D plan
// column pass, scale kernel, 2-D plan column pass).  HBM-bound: 16 * N * Nh bytes per image and pass.
//
// N = 256 = 16 x 16.  A CTA owns COLS adjacent kx columns of one image, 16 threads per column.
//   forward:  thread t loads rows t + 16 m (m = 0..15) of its column straight from global memory (lanes run along kx:
//             coalesced), FFT-16 over m in registers, twiddle W256^(t k1), exchange through shared memory,
//             FFT-16 over t: the thread acting as k1 now holds X[k1 + 16 k2], k2 = 0..15.
//   filter:   X *= G  (table row-major like the spectrum: coalesced).
//   inverse:  X[k1 + 16 k2] is again a "residue + 16 m" set, so the inverse starts from the registers:
//             inverse FFT-16 over k2, conjugate twiddle, exchange, inverse FFT-16, store rows t + 16 m.
// Shared memory is touched only by the two exchanges; all its accesses are unit-stride across lanes.
// N = 128 = 16 x 8 has the same three kernels with eight threads per transform, N = 320 = 20 x 16 with twenty threads per
// transform and CTA-level exchanges (second half of the file); other boxes take the generic ingest + 2-D cuFFT path of
// preprocess.cu.
#include "common.cuh"
#include "spline.cuh"

#include <math.h>
#include <algorithm>

namespace mem {

constexpr int CF_COLS = 22;                 // 6 slabs cover Nh = 129 (3 idle columns in the last slab)
constexpr int CF_THREADS = 16 * CF_COLS;

__constant__ float2 c_tw256[256];           // exp(-2 pi i j / 256)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// complex add / subtract as ONE packed instruction (FADD2 on sm_100: both halves of a 64-bit register pair)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long ua, ub, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
  return d;
}

// 4-point DFT, exponent sign S: (a, b, c, d) <- (X0, X1, X2, X3)
template <int S>
__device__ __forceinline__ void fft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c);
  const float2 s2 = cadd(b, d), s3 = csub(b, d);
  a = cadd(s0, s2);
  c = csub(s0, s2);
  if (S < 0) {   // X1 = s1 - i s3, X3 = s1 + i s3
    b = make_float2(s1.x + s3.y, s1.y - s3.x);
    d = make_float2(s1.x - s3.y, s1.y + s3.x);
  } else {
    b = make_float2(s1.x - s3.y, s1.y + s3.x);
    d = make_float2(s1.x + s3.y, s1.y - s3.x);
  }
}

// 16-point DFT in registers, natural order in and out: n = 4 n1 + n2, k = k1 + 4 k2
template <int S>
__device__ __forceinline__ void fft16(float2 (&x)[16]) {
  // cos / sin of 2 pi j / 16 for the exponents n2 * k1 that occur (j <= 9)
  constexpr float C[10] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                           -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f,
                           -0.92387953251128674f};
  constexpr float Sn[10] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f, 1.0f,
                            0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                            -0.38268343236508977f};
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    fft4<S>(x[n2], x[n2 + 4], x[n2 + 8], x[n2 + 12]);       // x[n2 + 4 k1] = A[n2][k1]
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) {
      if (n2 * k1 != 0) x[n2 + 4 * k1] = cmul(x[n2 + 4 * k1], make_float2(C[n2 * k1], (float)S * Sn[n2 * k1]));
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4<S>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);   // x[4 k1 + k2] = X[k1 + 4 k2]
  float2 y[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) y[k1 + 4 * k2] = x[4 * k1 + k2];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = y[k];
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// Persistent CTAs (two per SM): CTA b owns slab b % CF_SLABS and walks over the images b / CF_SLABS + j * stride.  The
// next image's slab is prefetched with cp.async into `stage` while the current one is transformed, so the global
// load latency is off the critical path; the 16 filter values of a thread do not depend on the image and are loaded
// once.
constexpr int CF_SLABS = 6;
// stats != NULL: the rows came from k_ingest_rowfft256, i.e. from images that are only shifted by a per-image offset,
// not yet normalised; (x - mean) / std is linear, so it is applied here: subtract mean * N^2 at (ky, kx) = (0, 0) and
// scale the filter by 1 / std.
// FWD: the column pass of the a10 transform — forward FFT along ky only, no filter, no way back.
template <bool FWD = false>
__global__ void __launch_bounds__(CF_THREADS, 2) k_colfilter256(float2* __restrict__ spec, const float* __restrict__ G,
                                                             const float2* __restrict__ stats, int Nh, int nS,
                                                             int img_stride) {
  extern __shared__ float2 cf_smem[];
  float2* ex = cf_smem;                           // [256][CF_COLS] exchange buffer
  float2* stage = cf_smem + 256 * CF_COLS;        // [256][CF_COLS] next image's slab
  const int t = threadIdx.x / CF_COLS, col = threadIdx.x - t * CF_COLS;
  const int slab = blockIdx.x % CF_SLABS;
  const int kx = slab * CF_COLS + col;
  const bool live = kx < Nh;
  const int kxc = live ? kx : 0;
  float gk[16];
  if (!FWD) {
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) gk[k2] = live ? G[(t + 16 * k2) * Nh + kxc] : 0.0f;
  }
  __shared__ float2 tws[256];                     // tws[k1][t] = W256^(t k1)
  if (threadIdx.x < 256) tws[threadIdx.x] = c_tw256[((threadIdx.x >> 4) * (threadIdx.x & 15)) & 255];
  __syncthreads();
  int img = blockIdx.x / CF_SLABS;
  if (img < nS) {
    const float2* src = spec + (size_t)img * 256 * Nh + kxc;
#pragma unroll
    for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 16 * m) * CF_COLS + col, src + (t + 16 * m) * Nh);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; img < nS; img += img_stride) {
    float2* base = spec + (size_t)img * 256 * Nh + kxc;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = stage[(t + 16 * m) * CF_COLS + col];   // own copies: no barrier needed
    const int nxt = img + img_stride;
    if (nxt < nS) {                                // the thread refills exactly the slots it has just read
      const float2* src = spec + (size_t)nxt * 256 * Nh + kxc;
#pragma unroll
      for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 16 * m) * CF_COLS + col, src + (t + 16 * m) * Nh);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    fft16<-1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      ex[(k1 * 16 + t) * CF_COLS + col] = cmul(v[k1], tws[k1 * 16 + t]);
    }
    __syncthreads();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = ex[(t * 16 + n2) * CF_COLS + col];
    fft16<-1>(v);                                   // v[k2] = X[t + 16 k2]
    if (FWD) {
      if (live) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) base[(t + 16 * k2) * Nh] = v[k2];
      }
      __syncthreads();                              // `ex` is rewritten by the next image
      continue;
    }
    float scale = 1.0f;
    if (stats) {
      const float2 ms = stats[img];                 // (mean - offset, 1 / std): what is left to subtract from the offset image
      scale = ms.y;
      if (t == 0 && kx == 0) v[0].x -= ms.x * 65536.0f;
    }
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      const float g = gk[k2] * scale;
      v[k2].x *= g;
      v[k2].y *= g;
    }
    fft16<1>(v);                                    // inverse over k2
    __syncthreads();                                // every thread is done reading the forward exchange
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float2 w = tws[n1 * 16 + t];
      ex[(n1 * 16 + t) * CF_COLS + col] = cmul(v[n1], make_float2(w.x, -w.y));
    }
    __syncthreads();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = ex[(t * 16 + k1) * CF_COLS + col];
    fft16<1>(v);                                    // v[m] = x[t + 16 m]
    if (live) {
#pragma unroll
      for (int m = 0; m < 16; ++m) base[(t + 16 * m) * Nh] = v[m];
    }
    __syncthreads();                                // `ex` is rewritten by the next image
  }
}

// ------------------------------------------------------------------------------------------------
// a2 + a3 + the row pass of a5 in one kernel (N = 256): read the raw particle once, write the row-transformed half
// spectrum.  One CTA per image, eight bands of 32 picture rows: the band is staged (transposed for SPIDER stacks,
// upside down for conjugates) in shared memory, then every 16 threads transform a PAIR of real rows as one complex
// FFT-256 (z = row1 + i row2) and separate the two half spectra, X1[k] = (Z[k] + conj Z[N-k]) / 2,
// X2[k] = (Z[k] - conj Z[N-k]) / 2i, with the partner values fetched by warp shuffles.  The moments of a3 are
// accumulated in the same pass over values shifted by the image's first pixel (keeps the DC term small) and handed
// to k_colfilter256, which applies the normalisation in Fourier space.
// ------------------------------------------------------------------------------------------------
constexpr int IR_BP = 257;      // band pitch (floats): odd, so the transposing stores are conflict-free
constexpr int IR_EP = 272;      // (k_colfilter-era constant; the row kernels keep their exchange inside the band, see below)
// Shared memory = the band alone (32.9 KB): the FFT exchange of a row pair — 256 float2 — lives in the two band rows the
// pair has just loaded into registers (2 x 257 floats = 2,056 B), element (a, b) at a * 16 + (b ^ a): unit-stride across the
// 16 lanes of a transform both ways (an XOR swizzle instead of the [16][17] padding, which would not fit).  With 64
// registers that is four CTAs per SM instead of three (the kernel is latency-bound: 24 -> 32 warps).
// PLAIN: the a10 transform of the aligned images (:344): picture-orientation rows, no flip, no offset, no moments.
template <int MB, bool PLAIN = false>
__global__ void __launch_bounds__(256, MB) k_ingest_rowfft256(const float* __restrict__ raw, const uint8_t* __restrict__ flip,
                                                           float2* __restrict__ spec, float2* __restrict__ stats,
                                                           int transposed) {
  constexpr int N = 256, Nh = 129;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [32][IR_BP]
  __shared__ double red[24];
  __shared__ float2 tws[256];                     // tws[k1][t] = W256^(t k1): the constant bank would serialise the 16
  tws[threadIdx.x] = c_tw256[((threadIdx.x >> 4) * (threadIdx.x & 15)) & 255];   // distinct t of a warp on every lookup
  const int i = blockIdx.x;
  const float* src = raw + (size_t)i * N * N;
  float2* out = spec + (size_t)i * N * Nh;
  const bool fl = PLAIN ? false : flip[i] != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = threadIdx.x >> 4, t = threadIdx.x & 15;
  const float half = 0.5f * N, r2lim = half * half;
  const float off = PLAIN ? 0.0f : src[0];
  double s = 0, s2 = 0;
  int cnt = 0;                                     // pixels outside the disc (the only ones that enter the sums)
  float2* e = reinterpret_cast<float2*>(band + (2 * p) * IR_BP);    // exchange of this row pair = its own two band rows
  const int partner = (lane & 16) + ((16 - t) & 15);
  for (int b0 = 0; b0 < N; b0 += 32) {
    float rs = 0.0f, rs2 = 0.0f;
    if (PLAIN) {
#pragma unroll
      for (int k = warp; k < 32; k += 8) {
        const float* rowp = src + (b0 + k) * N;
#pragma unroll
        for (int c = lane; c < N; c += 32) band[k * IR_BP + c] = rowp[c];
      }
    } else if (transposed) {                        // picture[rp][c] = raw[c][rp]; lanes run along rp (contiguous in raw)
      const int rp = b0 + lane;
      const int r = fl ? N - 1 - rp : rp;
      const float x = (float)rp - half + 1.0f, x2 = x * x;     // annularMask.py:24-30, centre (N/2-1, N/2)
#pragma unroll
      for (int c = warp; c < N; c += 8) {             // 32 independent loads in flight per thread
        const float v = src[c * N + r] - off;
        const float y = (float)c - half;
        const bool in = x2 + y * y < r2lim;
        const float m = in ? 0.0f : v;
        cnt += in ? 0 : 1;
        rs += m;
        rs2 = fmaf(m, m, rs2);
        band[lane * IR_BP + c] = v;
      }
    } else {
#pragma unroll
      for (int k = warp; k < 32; k += 8) {
        const int rp = b0 + k;
        const int r = fl ? N - 1 - rp : rp;
        const float x = (float)rp - half + 1.0f, x2 = x * x;
#pragma unroll
        for (int c = lane; c < N; c += 32) {
          const float v = src[r * N + c] - off;
          const float y = (float)c - half;
          const bool in = x2 + y * y < r2lim;
          const float m = in ? 0.0f : v;
          cnt += in ? 0 : 1;
          rs += m;
          rs2 = fmaf(m, m, rs2);
          band[k * IR_BP + c] = v;
        }
      }
    }
    s += (double)rs;
    s2 += (double)rs2;
    __syncthreads();
    float2 v[16];
    const float* b1 = band + (2 * p) * IR_BP + t;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = make_float2(b1[16 * m], b1[IR_BP + 16 * m]);
    fft16<-1>(v);
    __syncwarp();                                   // every lane of the pair has its rows in registers: they become the exchange
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) e[k1 * 16 + (t ^ k1)] = cmul(v[k1], tws[k1 * 16 + t]);
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = e[t * 16 + (n2 ^ t)];
    fft16<-1>(v);                                   // v[k2] = Z[t + 16 k2]
    float2* o1 = out + (b0 + 2 * p) * Nh + t;
    float2* o2 = o1 + Nh;
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {                // k = t + 16 k2 <= 127; partner N - k lives in thread 16 - t, slot 15 - k2
      float2 zn;
      zn.x = __shfl_sync(0xffffffffu, v[15 - k2].x, partner);
      zn.y = __shfl_sync(0xffffffffu, v[15 - k2].y, partner);
      if (t == 0) zn = v[(16 - k2) & 15];           // k = 16 k2: the partner 16 (16 - k2) is the thread's own
      const float2 zk = v[k2];
      o1[16 * k2] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      o2[16 * k2] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
    }
    if (t == 0) {                                   // k = 128 (Nyquist): real for both rows
      o1[128] = make_float2(v[8].x, 0.0f);
      o2[128] = make_float2(v[8].y, 0.0f);
    }
    __syncthreads();                                // the band and the exchange rows are rewritten by the next band
  }
  if (PLAIN) return;
  double dc = (double)cnt;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
  }
  if (lane == 0) { red[warp] = s; red[8 + warp] = s2; red[16 + warp] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0; s2 = 0; dc = 0;
    for (int w = 0; w < 8; ++w) { s += red[w]; s2 += red[8 + w]; dc += red[16 + w]; }
    // the population of a3 is b = x (1 - msk) over ALL N^2 pixels, zeros inside the disc included: undo the offset
    // in the sums (it was applied to the outside pixels only), then the image itself is (x' + off - mean) / std
    const double o = (double)off, n = (double)N * N;
    const double S1 = s + o * dc, S2 = s2 + 2.0 * o * s + o * o * dc;
    const double mean = S1 / n;
    const double var = S2 / n - mean * mean;
    stats[i] = make_float2((float)(mean - o), (float)(1.0 / sqrt(var)));
  }
}

// ------------------------------------------------------------------------------------------------
// The way back (N = 256): inverse row transform of the filtered half spectra fused with what follows it in the
// pipeline, the annular mask (:325) and the row pass of the first periodic spline prefilter (rotatefill.py:24,
// same recursion as k_prefilter_rows_x<8, true> in align.cu).  One CTA per image, bands of 32 rows: 16 threads build
// z = X1 + i X2 from the half spectra of a row pair (Z[k] = X1[k] + i X2[k], Z[N-k] = conj X1[k] + i conj X2[k]),
// run the inverse FFT-256 and leave the two real rows in shared memory; then every warp filters four rows, lane l
// owning samples [8 l, 8 l + 8), and stores them row-coalesced.  The low-passed image itself never reaches HBM.
// ------------------------------------------------------------------------------------------------
constexpr int RP_BP = 264;      // band pitch: 256 samples + one pad float per 32 (conflict-free for both access patterns)
template <int MB>
__global__ void __launch_bounds__(256, MB) k_rowifft_prefilter256(const float2* __restrict__ spec, float* __restrict__ outimg) {
  constexpr int N = 256, Nh = 129, E = 8;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [32][RP_BP]; the exchange of a row pair lives in its rows
  const int i = blockIdx.x;
  const float2* in = spec + (size_t)i * N * Nh;
  float* dst = outimg + (size_t)i * N * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = threadIdx.x >> 4, t = threadIdx.x & 15;
  float2* e = reinterpret_cast<float2*>(band + (2 * p) * RP_BP);    // 2 x 264 floats >= 256 float2, swizzled as in the forward kernel
  constexpr float half = 0.5f * N, r2lim = half * half;
  const float zE = 2.6571717e-05f;                  // z^8, z = sqrt(3) - 2
  __shared__ float2 tws[256];                       // tws[n1][t] = conj W256^(t n1), see k_ingest_rowfft256
  {
    const float2 w = c_tw256[((threadIdx.x >> 4) * (threadIdx.x & 15)) & 255];
    tws[threadIdx.x] = make_float2(w.x, -w.y);
  }
  __syncthreads();
  for (int b0 = 0; b0 < N; b0 += 32) {
    const float2* x1 = in + (b0 + 2 * p) * Nh;
    const float2* x2 = x1 + Nh;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 8; ++m) {                   // k = t + 16 m <= 127
      float2 a = x1[t + 16 * m], b = x2[t + 16 * m];
      if (m == 0 && t == 0) { a.y = 0.0f; b.y = 0.0f; }            // a C2R ignores the imaginary part of the DC term
      v[m] = make_float2(a.x - b.y, a.y + b.x);
    }
#pragma unroll
    for (int m = 8; m < 16; ++m) {                  // k = t + 16 m >= 128: conjugate of entry N - k
      const int kk = N - t - 16 * m;
      float2 a = x1[kk], b = x2[kk];
      if (m == 8 && t == 0) { a.y = 0.0f; b.y = 0.0f; }            // ... and of the Nyquist term
      v[m] = make_float2(a.x + b.y, b.x - a.y);
    }
    fft16<1>(v);
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) e[n1 * 16 + (t ^ n1)] = cmul(v[n1], tws[n1 * 16 + t]);
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = e[t * 16 + (k1 ^ t)];
    fft16<1>(v);                                    // v[m] = z[t + 16 m]: row 2p in .x, row 2p + 1 in .y
    __syncwarp();                                   // the exchange has been read by every lane: its memory becomes the two rows
    float* r1 = band + (2 * p) * RP_BP;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int c = t + 16 * m;
      r1[c + (c >> 5)] = v[m].x;
      r1[RP_BP + c + (c >> 5)] = v[m].y;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {                   // warp-private rows from here on; unrolled: four independent recursions in flight
      const int row = warp + 8 * q;
      float* ln = band + row * RP_BP;
      const float xm = (float)(b0 + row) - half + 1.0f;
      const float lim = r2lim - xm * xm;            // keep the pixel iff y * y < lim
      const int c0 = lane * E;
      float s[E];
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const float y = (float)(c0 + j) - half;
        const float val = ln[c0 + j + ((c0 + j) >> 5)];
        s[j] = (y * y < lim) ? 6.0f * val : 0.0f;
      }
      spline_line_warp<E>(s, zE, lane);
      // lane l holds samples [8 l, 8 l + 8): two 16-byte stores per lane, the warp writes its 1 KB row contiguously
      float4* orow = reinterpret_cast<float4*>(dst + (b0 + row) * N + c0);
      orow[0] = make_float4(s[0], s[1], s[2], s[3]);
      orow[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();                                // the band is rewritten by the next row pairs
  }
}

// ================================================================================================
// N = 128 = 16 x 8 (BASELINE config 2).  The same three passes with EIGHT threads per transform: thread t holds
// z[t + 8 m], m < 16; FFT-16 over m in registers, twiddle W128^(t k1), exchange, and two FFT-8 over t for
// k1 = t and t + 8 leave X[t + 8 j], j = g + 2 k2 — the same "residue + 8 j" set the transform started from, so
// forward, filter and inverse chain in registers exactly as at 256.
// ================================================================================================
// 8-point DFT in registers, natural order in and out: n = 2 n1 + n2, k = k1 + 4 k2
template <int S>
__device__ __forceinline__ void fft8(float2 (&x)[8]) {
  constexpr float R = 0.70710678118654752f;
  fft4<S>(x[0], x[2], x[4], x[6]);                 // x[2 k1]     = A[0][k1]
  fft4<S>(x[1], x[3], x[5], x[7]);                 // x[2 k1 + 1] = A[1][k1]
  x[3] = cmul(x[3], make_float2(R, (float)S * R));
  x[5] = (S < 0) ? make_float2(x[5].y, -x[5].x) : make_float2(-x[5].y, x[5].x);
  x[7] = cmul(x[7], make_float2(-R, (float)S * R));
  float2 y[8];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    y[k1] = cadd(x[2 * k1], x[2 * k1 + 1]);
    y[k1 + 4] = csub(x[2 * k1], x[2 * k1 + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = y[k];
}

// second half of the 128-point transform of a ROW kernel: e = the transform's 128 float2 of exchange space, element
// (k1, t) at k1 * 8 + (t ^ (k1 & 7)) — unit-stride across the 8 lanes both ways; the two transforms of a half-warp sit
// 136 float2 apart, so they use complementary halves of the 16 eight-byte banks.  tw[k1 * 8 + t] = W128^(+-t k1).
template <int S>
__device__ __forceinline__ void fft128_finish(float2 (&v)[16], float2* e, const float2* tw, int t) {
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) e[k1 * 8 + (t ^ (k1 & 7))] = cmul(v[k1], tw[k1 * 8 + t]);
  __syncwarp();
  float2 u0[8], u1[8];
#pragma unroll
  for (int tt = 0; tt < 8; ++tt) {
    u0[tt] = e[t * 8 + (tt ^ t)];
    u1[tt] = e[(t + 8) * 8 + (tt ^ t)];
  }
  fft8<S>(u0);
  fft8<S>(u1);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    v[2 * k2] = u0[k2];                              // X[t + 16 k2]
    v[2 * k2 + 1] = u1[k2];                          // X[t + 8 + 16 k2]
  }
}

constexpr int CF8_COLS = 22;                  // 3 slabs cover Nh = 65 (one idle column)
constexpr int CF8_SLABS = 3;
constexpr int CF8_THREADS = 8 * CF8_COLS;

template <bool FWD = false>
__global__ void __launch_bounds__(CF8_THREADS, 4) k_colfilter128(float2* __restrict__ spec, const float* __restrict__ G,
                                                               const float2* __restrict__ stats, int Nh, int nS,
                                                               int img_stride) {
  extern __shared__ float2 cf_smem[];
  float2* ex = cf_smem;                           // [128][CF8_COLS]
  float2* stage = cf_smem + 128 * CF8_COLS;       // next image's slab
  const int t = threadIdx.x / CF8_COLS, col = threadIdx.x - t * CF8_COLS;
  const int slab = blockIdx.x % CF8_SLABS;
  const int kx = slab * CF8_COLS + col;
  const bool live = kx < Nh;
  const int kxc = live ? kx : 0;
  float gk[16];
  if (!FWD) {
#pragma unroll
    for (int j = 0; j < 16; ++j) gk[j] = live ? G[(t + 8 * j) * Nh + kxc] : 0.0f;
  }
  __shared__ float2 tws[128];                     // tws[k1][t] = W128^(t k1)
  if (threadIdx.x < 128) tws[threadIdx.x] = c_tw256[(2 * (threadIdx.x >> 3) * (threadIdx.x & 7)) & 255];
  __syncthreads();
  int img = blockIdx.x / CF8_SLABS;
  if (img < nS) {
    const float2* src = spec + (size_t)img * 128 * Nh + kxc;
#pragma unroll
    for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 8 * m) * CF8_COLS + col, src + (t + 8 * m) * Nh);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; img < nS; img += img_stride) {
    float2* base = spec + (size_t)img * 128 * Nh + kxc;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = stage[(t + 8 * m) * CF8_COLS + col];
    const int nxt = img + img_stride;
    if (nxt < nS) {
      const float2* src = spec + (size_t)nxt * 128 * Nh + kxc;
#pragma unroll
      for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 8 * m) * CF8_COLS + col, src + (t + 8 * m) * Nh);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    float2 u0[8], u1[8];
    fft16<-1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) ex[(k1 * 8 + t) * CF8_COLS + col] = cmul(v[k1], tws[k1 * 8 + t]);
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      u0[tt] = ex[(t * 8 + tt) * CF8_COLS + col];
      u1[tt] = ex[((t + 8) * 8 + tt) * CF8_COLS + col];
    }
    fft8<-1>(u0);
    fft8<-1>(u1);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      v[2 * k2] = u0[k2];
      v[2 * k2 + 1] = u1[k2];                       // v[j] = X[t + 8 j]
    }
    if (FWD) {
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; ++j) base[(t + 8 * j) * Nh] = v[j];
      }
      __syncthreads();
      continue;
    }
    float scale = 1.0f;
    if (stats) {
      const float2 ms = stats[img];
      scale = ms.y;
      if (t == 0 && kx == 0) v[0].x -= ms.x * 16384.0f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float g = gk[j] * scale;
      v[j].x *= g;
      v[j].y *= g;
    }
    fft16<1>(v);                                    // inverse over j
    __syncthreads();                                // every thread is done reading the forward exchange
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float2 w = tws[n1 * 8 + t];
      ex[(n1 * 8 + t) * CF8_COLS + col] = cmul(v[n1], make_float2(w.x, -w.y));
    }
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      u0[tt] = ex[(t * 8 + tt) * CF8_COLS + col];
      u1[tt] = ex[((t + 8) * 8 + tt) * CF8_COLS + col];
    }
    fft8<1>(u0);
    fft8<1>(u1);
    if (live) {
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        base[(t + 16 * k2) * Nh] = u0[k2];
        base[(t + 8 + 16 * k2) * Nh] = u1[k2];
      }
    }
    __syncthreads();                                // `ex` is rewritten by the next image
  }
}

// a2 + a3 + row pass of a5 at N = 128: one CTA per image, two bands of 64 picture rows; a warp owns 8 band rows = 4 row
// pairs, 8 threads per pair.  The exchange of the warp's four transforms lives in those 8 rows once every lane has its
// values in registers (8 x 133 floats = 532 float2 >= 528).
constexpr int IR8_BP = 133;
template <bool PLAIN = false>
__global__ void __launch_bounds__(256, 4) k_ingest_rowfft128(const float* __restrict__ raw, const uint8_t* __restrict__ flip,
                                                          float2* __restrict__ spec, float2* __restrict__ stats,
                                                          int transposed) {
  constexpr int N = 128, Nh = 65;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [64][IR8_BP]
  __shared__ double red[24];
  __shared__ float2 tws[128];
  if (threadIdx.x < 128) tws[threadIdx.x] = c_tw256[(2 * (threadIdx.x >> 3) * (threadIdx.x & 7)) & 255];
  const int i = blockIdx.x;
  const float* src = raw + (size_t)i * N * N;
  float2* out = spec + (size_t)i * N * Nh;
  const bool fl = PLAIN ? false : flip[i] != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = threadIdx.x >> 3, t = threadIdx.x & 7, q = lane >> 3;
  const float half = 0.5f * N, r2lim = half * half;
  const float off = PLAIN ? 0.0f : src[0];
  double s = 0, s2 = 0;
  int cnt = 0;
  float2* e = reinterpret_cast<float2*>(band + (8 * warp) * IR8_BP) + q * 128 + 8 * ((q + 1) >> 1);
  const int partner = (lane & 24) + ((8 - t) & 7);
  for (int b0 = 0; b0 < N; b0 += 64) {
    float rs = 0.0f, rs2 = 0.0f;
    if (PLAIN) {
#pragma unroll
      for (int k = warp; k < 64; k += 8) {
        const float* rowp = src + (b0 + k) * N;
#pragma unroll
        for (int c = lane; c < N; c += 32) band[k * IR8_BP + c] = rowp[c];
      }
    } else if (transposed) {                        // picture[rp][c] = raw[c][rp]; lanes run along rp
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rp = b0 + 32 * h + lane;
        const int r = fl ? N - 1 - rp : rp;
        const float x = (float)rp - half + 1.0f, x2 = x * x;
#pragma unroll
        for (int c = warp; c < N; c += 8) {
          const float v = src[c * N + r] - off;
          const float y = (float)c - half;
          const bool in = x2 + y * y < r2lim;
          const float m = in ? 0.0f : v;
          cnt += in ? 0 : 1;
          rs += m;
          rs2 = fmaf(m, m, rs2);
          band[(32 * h + lane) * IR8_BP + c] = v;
        }
      }
    } else {
#pragma unroll
      for (int k = warp; k < 64; k += 8) {
        const int rp = b0 + k;
        const int r = fl ? N - 1 - rp : rp;
        const float x = (float)rp - half + 1.0f, x2 = x * x;
#pragma unroll
        for (int c = lane; c < N; c += 32) {
          const float v = src[r * N + c] - off;
          const float y = (float)c - half;
          const bool in = x2 + y * y < r2lim;
          const float m = in ? 0.0f : v;
          cnt += in ? 0 : 1;
          rs += m;
          rs2 = fmaf(m, m, rs2);
          band[k * IR8_BP + c] = v;
        }
      }
    }
    s += (double)rs;
    s2 += (double)rs2;
    __syncthreads();
    float2 v[16];
    const float* b1 = band + (2 * p) * IR8_BP + t;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = make_float2(b1[8 * m], b1[IR8_BP + 8 * m]);
    fft16<-1>(v);
    __syncwarp();                                   // every lane of the warp has its rows in registers: they become the exchange
    fft128_finish<-1>(v, e, tws, t);                // v[j] = Z[t + 8 j]
    float2* o1 = out + (b0 + 2 * p) * Nh + t;
    float2* o2 = o1 + Nh;
#pragma unroll
    for (int j = 0; j < 8; ++j) {                   // k = t + 8 j <= 63; partner N - k lives in thread 8 - t, slot 15 - j
      float2 zn;
      zn.x = __shfl_sync(0xffffffffu, v[15 - j].x, partner);
      zn.y = __shfl_sync(0xffffffffu, v[15 - j].y, partner);
      if (t == 0) zn = v[(16 - j) & 15];            // k = 8 j: the partner 8 (16 - j) is the thread's own
      const float2 zk = v[j];
      o1[8 * j] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      o2[8 * j] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
    }
    if (t == 0) {                                   // k = 64 (Nyquist): real for both rows
      o1[64] = make_float2(v[8].x, 0.0f);
      o2[64] = make_float2(v[8].y, 0.0f);
    }
    __syncthreads();
  }
  if (PLAIN) return;
  double dc = (double)cnt;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
  }
  if (lane == 0) { red[warp] = s; red[8 + warp] = s2; red[16 + warp] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0; s2 = 0; dc = 0;
    for (int w = 0; w < 8; ++w) { s += red[w]; s2 += red[8 + w]; dc += red[16 + w]; }
    const double o = (double)off, n = (double)N * N;
    const double S1 = s + o * dc, S2 = s2 + 2.0 * o * s + o * o * dc;
    const double mean = S1 / n;
    const double var = S2 / n - mean * mean;
    stats[i] = make_float2((float)(mean - o), (float)(1.0 / sqrt(var)));
  }
}

// the way back at N = 128: inverse row transform + annular mask + row pass of the first spline prefilter; bands of 64
// rows, pitch 132 (128 samples + one pad float per 32); lane l of the filtering warp owns samples [4 l, 4 l + 4)
constexpr int RP8_BP = 132;
__global__ void __launch_bounds__(256, 4) k_rowifft_prefilter128(const float2* __restrict__ spec, float* __restrict__ outimg) {
  constexpr int N = 128, Nh = 65, E = 4;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [64][RP8_BP]
  const int i = blockIdx.x;
  const float2* in = spec + (size_t)i * N * Nh;
  float* dst = outimg + (size_t)i * N * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = threadIdx.x >> 3, t = threadIdx.x & 7, q = lane >> 3;
  float2* e = reinterpret_cast<float2*>(band + (8 * warp) * RP8_BP) + q * 128 + 8 * ((q + 1) >> 1);   // 8 x 132 floats = 528 float2
  constexpr float half = 0.5f * N, r2lim = half * half;
  const float zE = 5.1547761e-03f;                  // z^4, z = sqrt(3) - 2
  __shared__ float2 tws[128];                       // conj W128^(t n1)
  if (threadIdx.x < 128) {
    const float2 w = c_tw256[(2 * (threadIdx.x >> 3) * (threadIdx.x & 7)) & 255];
    tws[threadIdx.x] = make_float2(w.x, -w.y);
  }
  __syncthreads();
  for (int b0 = 0; b0 < N; b0 += 64) {
    const float2* x1 = in + (b0 + 2 * p) * Nh;
    const float2* x2 = x1 + Nh;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 8; ++m) {                   // k = t + 8 m <= 63
      float2 a = x1[t + 8 * m], b = x2[t + 8 * m];
      if (m == 0 && t == 0) { a.y = 0.0f; b.y = 0.0f; }
      v[m] = make_float2(a.x - b.y, a.y + b.x);
    }
#pragma unroll
    for (int m = 8; m < 16; ++m) {                  // k = t + 8 m >= 64: conjugate of entry N - k
      const int kk = N - t - 8 * m;
      float2 a = x1[kk], b = x2[kk];
      if (m == 8 && t == 0) { a.y = 0.0f; b.y = 0.0f; }
      v[m] = make_float2(a.x + b.y, b.x - a.y);
    }
    fft16<1>(v);
    fft128_finish<1>(v, e, tws, t);                 // v[j] = z[t + 8 j]: row 2p in .x, row 2p + 1 in .y
    __syncwarp();                                   // the exchange has been read by every lane of the warp: its memory becomes rows
    float* r1 = band + (2 * p) * RP8_BP;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = t + 8 * j;
      r1[c + (c >> 5)] = v[j].x;
      r1[RP8_BP + c + (c >> 5)] = v[j].y;
    }
    __syncthreads();
#pragma unroll
    for (int qq = 0; qq < 8; ++qq) {
      const int row = warp + 8 * qq;
      float* ln = band + row * RP8_BP;
      const float xm = (float)(b0 + row) - half + 1.0f;
      const float lim = r2lim - xm * xm;
      const int c0 = lane * E;
      float sv[E];
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const float y = (float)(c0 + j) - half;
        const float val = ln[c0 + j + ((c0 + j) >> 5)];
        sv[j] = (y * y < lim) ? 6.0f * val : 0.0f;
      }
      spline_line_warp<E>(sv, zE, lane);
      *reinterpret_cast<float4*>(dst + (b0 + row) * N + c0) = make_float4(sv[0], sv[1], sv[2], sv[3]);
    }
    __syncthreads();
  }
}

// ================================================================================================
// N = 320 = 20 x 16 (BASELINE config 5).  A warp-sized thread group cannot hold a 320-point transform with uniform work, so the
// split lives on CTA level.  Column pass: 20 threads per column, thread t holds x[t + 20 m],
// m < 16: FFT-16 in registers, twiddle W320^(t k1), exchange, then threads k1 < 16 run an FFT-20 (4 x 5) over t and hold
// X[k1 + 16 k2], k2 < 20; filter; inverse FFT-20 over k2, conjugate twiddle, exchange, inverse FFT-16 on all 20 threads:
// x[t + 20 n2] again.  The four thread rows t >= 16 idle in the FFT-20 stages as whole warps.
// ================================================================================================
__constant__ float2 c_tw320[320];           // exp(-2 pi i j / 320)

// 5-point DFT in registers, natural order in and out
template <int S>
__device__ __forceinline__ void fft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  constexpr float C1 = 0.30901699437494745f, C2 = -0.80901699437494745f, S1 = 0.95105651629515353f, S2 = 0.58778525229247314f;
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 a1 = make_float2(x0.x + C1 * t1.x + C2 * t2.x, x0.y + C1 * t1.y + C2 * t2.y);
  const float2 a2 = make_float2(x0.x + C2 * t1.x + C1 * t2.x, x0.y + C2 * t1.y + C1 * t2.y);
  const float2 b1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
  const float2 b2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
  x0 = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
  // forward (S < 0): X1 = a1 - i b1, X4 = a1 + i b1, X2 = a2 - i b2, X3 = a2 + i b2;  -i (p, q) = (q, -p)
  const float sg = (S < 0) ? 1.0f : -1.0f;
  x1 = make_float2(a1.x + sg * b1.y, a1.y - sg * b1.x);
  x4 = make_float2(a1.x - sg * b1.y, a1.y + sg * b1.x);
  x2 = make_float2(a2.x + sg * b2.y, a2.y - sg * b2.x);
  x3 = make_float2(a2.x - sg * b2.y, a2.y + sg * b2.x);
}

// 20-point DFT in registers, natural order in and out: n = 5 n1 + n2, k = k1 + 4 k2
template <int S>
__device__ __forceinline__ void fft20(float2 (&x)[20]) {
  // cos / sin of 2 pi j / 20, j = n2 k1 <= 12
  constexpr float C[13] = {1.0f, 0.95105651629515353f, 0.80901699437494745f, 0.58778525229247314f, 0.30901699437494745f, 0.0f,
                           -0.30901699437494745f, -0.58778525229247314f, -0.80901699437494745f, -0.95105651629515353f, -1.0f,
                           -0.95105651629515353f, -0.80901699437494745f};
  constexpr float Sn[13] = {0.0f, 0.30901699437494745f, 0.58778525229247314f, 0.80901699437494745f, 0.95105651629515353f, 1.0f,
                            0.95105651629515353f, 0.80901699437494745f, 0.58778525229247314f, 0.30901699437494745f, 0.0f,
                            -0.30901699437494745f, -0.58778525229247314f};
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) {
    fft4<S>(x[n2], x[5 + n2], x[10 + n2], x[15 + n2]);      // x[n2 + 5 k1] = A[n2][k1]
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
      if (n2 != 0) x[n2 + 5 * k1] = cmul(x[n2 + 5 * k1], make_float2(C[n2 * k1], (float)S * Sn[n2 * k1]));
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft5<S>(x[5 * k1], x[5 * k1 + 1], x[5 * k1 + 2], x[5 * k1 + 3], x[5 * k1 + 4]);   // x[5 k1 + k2] = X[k1 + 4 k2]
  float2 y[20];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) y[k1 + 4 * k2] = x[5 * k1 + k2];
#pragma unroll
  for (int k = 0; k < 20; ++k) x[k] = y[k];
}

constexpr int CF20_COLS = 16;                 // 11 slabs cover Nh = 161 (15 idle columns in the last slab)
constexpr int CF20_SLABS = 11;
constexpr int CF20_THREADS = 20 * CF20_COLS;

template <bool FWD = false>
__global__ void __launch_bounds__(CF20_THREADS, 2) k_colfilter320(float2* __restrict__ spec, const float* __restrict__ G,
                                                               const float2* __restrict__ stats, int nS, int img_stride) {
  constexpr int N = 320, Nh = 161;
  extern __shared__ float2 cf_smem[];
  float2* ex = cf_smem;                           // [320][CF20_COLS]
  float2* stage = cf_smem + N * CF20_COLS;        // next image's slab
  const int t = threadIdx.x / CF20_COLS, col = threadIdx.x - t * CF20_COLS;
  const int slab = blockIdx.x % CF20_SLABS;
  const int kx = slab * CF20_COLS + col;
  const bool live = kx < Nh;
  const int kxc = live ? kx : 0;
  const bool second = t < 16;                     // the threads that run the FFT-20 stages (as k1 = t)
  float gk[20];
  if (!FWD) {
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) gk[k2] = (live && second) ? G[(t + 16 * k2) * Nh + kxc] : 0.0f;
  }
  __shared__ float2 tws[320];                     // tws[k1 * 20 + t] = W320^(t k1), k1 < 16, t < 20
  tws[threadIdx.x] = c_tw320[((threadIdx.x / 20) * (threadIdx.x % 20)) % 320];
  __syncthreads();
  int img = blockIdx.x / CF20_SLABS;
  if (img < nS) {
    const float2* src = spec + (size_t)img * N * Nh + kxc;
#pragma unroll
    for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 20 * m) * CF20_COLS + col, src + (t + 20 * m) * Nh);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; img < nS; img += img_stride) {
    float2* base = spec + (size_t)img * N * Nh + kxc;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = stage[(t + 20 * m) * CF20_COLS + col];
    const int nxt = img + img_stride;
    if (nxt < nS) {
      const float2* src = spec + (size_t)nxt * N * Nh + kxc;
#pragma unroll
      for (int m = 0; m < 16; ++m) cp_async8(stage + (t + 20 * m) * CF20_COLS + col, src + (t + 20 * m) * Nh);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    fft16<-1>(v);                                   // v[k1] = A_t[k1]
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) ex[(k1 * 20 + t) * CF20_COLS + col] = cmul(v[k1], tws[k1 * 20 + t]);
    __syncthreads();
    float2 u[20];
    if (second) {
#pragma unroll
      for (int tt = 0; tt < 20; ++tt) u[tt] = ex[(t * 20 + tt) * CF20_COLS + col];
      fft20<-1>(u);                                 // u[k2] = X[t + 16 k2]
      if (FWD) {
        if (live) {
#pragma unroll
          for (int k2 = 0; k2 < 20; ++k2) base[(t + 16 * k2) * Nh] = u[k2];
        }
      } else {
        float scale = 1.0f;
        if (stats) {                                // rows from k_ingest_rowfft320: normalise here, (x - mean) / std is linear
          const float2 ms = stats[img];
          scale = ms.y;
          if (t == 0 && kx == 0) u[0].x -= ms.x * 102400.0f;
        }
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) {
          const float g = gk[k2] * scale;
          u[k2].x *= g;
          u[k2].y *= g;
        }
        fft20<1>(u);                                // inverse over k2: u[n1] = C_t[n1]
      }
    }
    __syncthreads();                                // every thread is done reading the forward exchange
    if (FWD) continue;
    if (second) {
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) {
        const float2 w = tws[t * 20 + n1];          // W320^(n1 k1), k1 = t
        ex[(n1 * 16 + t) * CF20_COLS + col] = cmul(u[n1], make_float2(w.x, -w.y));
      }
    }
    __syncthreads();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = ex[(t * 16 + k1) * CF20_COLS + col];
    fft16<1>(v);                                    // v[n2] = x[t + 20 n2]
    if (live) {
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) base[(t + 20 * n2) * Nh] = v[n2];
    }
    __syncthreads();                                // `ex` is rewritten by the next image
  }
}

// Row passes at N = 320 with the same 20 x 16 split on CTA level: a band of 32 picture rows = 16 row pairs, thread (t, pair)
// with the pair index fastest, exchanges through the band's own shared memory behind block barriers (the band is in registers
// by then).  The two half spectra of a pair are separated from a natural-order copy of Z in shared memory by the store loop
// (lanes along kx: coalesced).
constexpr int R20_PR = 16;                    // row pairs per band
constexpr int R20_THREADS = 20 * R20_PR;      // 320
constexpr int R20_BP = 321;                   // band pitch (floats) of the forward kernel; also the pitch (float2) of the Z copy
constexpr int R20_RBP = 331;                  // band pitch of the inverse kernel: 320 samples + one pad per 32, odd multiple mod 16
constexpr int R20_SMEM = 32 * R20_RBP * 4;    // 42,368 B >= every view of the buffer

// first half of the forward transform of one row pair: v[m] = z[t + 20 m] -> Z[t + 16 k2] in u (threads t < 16)
__device__ __forceinline__ void fwd320_pair(float2 (&v)[16], float2 (&u)[20], float2* ex, const float2* tws, int t, int pr) {
  fft16<-1>(v);
  __syncthreads();                                  // every thread has its band values in registers: the band becomes the exchange
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) ex[(k1 * 20 + t) * R20_PR + pr] = cmul(v[k1], tws[k1 * 20 + t]);
  __syncthreads();
  if (t < 16) {
#pragma unroll
    for (int tt = 0; tt < 20; ++tt) u[tt] = ex[(t * 20 + tt) * R20_PR + pr];
    fft20<-1>(u);
  }
  __syncthreads();                                  // the exchange has been read: its memory becomes the natural-order copy of Z
}

template <bool PLAIN = false>
__global__ void __launch_bounds__(R20_THREADS, 2) k_ingest_rowfft320(const float* __restrict__ raw, const uint8_t* __restrict__ flip,
                                                                  float2* __restrict__ spec, float2* __restrict__ stats,
                                                                  int transposed) {
  constexpr int N = 320, Nh = 161, NW = R20_THREADS / 32;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [32][R20_BP]
  float2* ex = ir_smem;                                             // [320][R20_PR]
  float2* zb = ir_smem;                                             // [R20_PR][R20_BP] natural-order Z of every pair
  __shared__ double red[3 * NW];
  __shared__ float2 tws[320];
  tws[threadIdx.x] = c_tw320[((threadIdx.x / 20) * (threadIdx.x % 20)) % 320];
  const int i = blockIdx.x;
  const float* src = raw + (size_t)i * N * N;
  float2* out = spec + (size_t)i * N * Nh;
  const bool fl = PLAIN ? false : flip[i] != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = threadIdx.x / R20_PR, pr = threadIdx.x % R20_PR;
  const float half = 0.5f * N, r2lim = half * half;
  const float off = PLAIN ? 0.0f : src[0];
  double s = 0, s2 = 0;
  int cnt = 0;
  __syncthreads();
  for (int b0 = 0; b0 < N; b0 += 32) {
    float rs = 0.0f, rs2 = 0.0f;
    if (PLAIN) {
      for (int k = warp; k < 32; k += NW) {
        const float* rowp = src + (b0 + k) * N;
#pragma unroll
        for (int c = lane; c < N; c += 32) band[k * R20_BP + c] = rowp[c];
      }
    } else if (transposed) {                        // picture[rp][c] = raw[c][rp]; lanes run along rp
      const int rp = b0 + lane;
      const int r = fl ? N - 1 - rp : rp;
      const float x = (float)rp - half + 1.0f, x2 = x * x;
#pragma unroll 8
      for (int c = warp; c < N; c += NW) {
        const float v = src[c * N + r] - off;
        const float y = (float)c - half;
        const bool in = x2 + y * y < r2lim;
        const float m = in ? 0.0f : v;
        cnt += in ? 0 : 1;
        rs += m;
        rs2 = fmaf(m, m, rs2);
        band[lane * R20_BP + c] = v;
      }
    } else {
      for (int k = warp; k < 32; k += NW) {
        const int rp = b0 + k;
        const int r = fl ? N - 1 - rp : rp;
        const float x = (float)rp - half + 1.0f, x2 = x * x;
#pragma unroll
        for (int c = lane; c < N; c += 32) {
          const float v = src[r * N + c] - off;
          const float y = (float)c - half;
          const bool in = x2 + y * y < r2lim;
          const float m = in ? 0.0f : v;
          cnt += in ? 0 : 1;
          rs += m;
          rs2 = fmaf(m, m, rs2);
          band[k * R20_BP + c] = v;
        }
      }
    }
    s += (double)rs;
    s2 += (double)rs2;
    __syncthreads();
    float2 v[16], u[20];
    const float* b1 = band + (2 * pr) * R20_BP + t;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = make_float2(b1[20 * m], b1[R20_BP + 20 * m]);
    fwd320_pair(v, u, ex, tws, t, pr);
    if (t < 16) {
#pragma unroll
      for (int k2 = 0; k2 < 20; ++k2) zb[pr * R20_BP + t + 16 * k2] = u[k2];
    }
    __syncthreads();
    for (int q = warp; q < R20_PR; q += NW) {       // X1[k] = (Z[k] + conj Z[N-k]) / 2, X2[k] = (Z[k] - conj Z[N-k]) / 2i
      const float2* z = zb + q * R20_BP;
      float2* o1 = out + (b0 + 2 * q) * Nh;
      float2* o2 = o1 + Nh;
      for (int k = lane; k < Nh; k += 32) {
        const float2 zk = z[k];
        const float2 zn = z[k == 0 ? 0 : N - k];
        o1[k] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        o2[k] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
      }
    }
    __syncthreads();                                // the buffer is the next band
  }
  if (PLAIN) return;
  double dc = (double)cnt;
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    dc += __shfl_xor_sync(0xffffffffu, dc, o);
  }
  if (lane == 0) { red[warp] = s; red[NW + warp] = s2; red[2 * NW + warp] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0; s2 = 0; dc = 0;
    for (int w = 0; w < NW; ++w) { s += red[w]; s2 += red[NW + w]; dc += red[2 * NW + w]; }
    const double o = (double)off, n = (double)N * N;
    const double S1 = s + o * dc, S2 = s2 + 2.0 * o * s + o * o * dc;
    const double mean = S1 / n;
    const double var = S2 / n - mean * mean;
    stats[i] = make_float2((float)(mean - o), (float)(1.0 / sqrt(var)));
  }
}

// the way back at N = 320: half spectra of a row pair -> Z in natural order -> inverse FFT-320 (FFT-20 on threads t < 16,
// conjugate twiddle, exchange, FFT-16 on all) -> the two real rows in shared memory -> annular mask + row pass of the first
// spline prefilter (lane l owns samples [10 l, 10 l + 10)), row-contiguous stores
__global__ void __launch_bounds__(R20_THREADS, 2) k_rowifft_prefilter320(const float2* __restrict__ spec, float* __restrict__ outimg) {
  constexpr int N = 320, Nh = 161, E = 10, NW = R20_THREADS / 32;
  extern __shared__ float2 ir_smem[];
  float* band = reinterpret_cast<float*>(ir_smem);                  // [32][R20_RBP]
  float2* ex = ir_smem;                                             // [320][R20_PR]
  float2* zb = ir_smem;                                             // [R20_PR][R20_BP]
  const int i = blockIdx.x;
  const float2* in = spec + (size_t)i * N * Nh;
  float* dst = outimg + (size_t)i * N * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = threadIdx.x / R20_PR, pr = threadIdx.x % R20_PR;
  constexpr float half = 0.5f * N, r2lim = half * half;
  const float zE = 1.9077634e-06f;                  // z^10, z = sqrt(3) - 2
  __shared__ float2 tws[320];                       // conj W320^(t k1)
  {
    const float2 w = c_tw320[((threadIdx.x / 20) * (threadIdx.x % 20)) % 320];
    tws[threadIdx.x] = make_float2(w.x, -w.y);
  }
  __syncthreads();
  for (int b0 = 0; b0 < N; b0 += 32) {
    for (int q = warp; q < R20_PR; q += NW) {       // Z[k] = X1[k] + i X2[k], Z[N-k] = conj X1[k] + i conj X2[k]
      const float2* x1 = in + (b0 + 2 * q) * Nh;
      const float2* x2 = x1 + Nh;
      float2* z = zb + q * R20_BP;
      for (int k = lane; k < Nh; k += 32) {
        float2 a = x1[k], b = x2[k];
        if (k == 0 || k == N / 2) { a.y = 0.0f; b.y = 0.0f; }      // a C2R ignores the imaginary parts of the DC and Nyquist terms
        z[k] = make_float2(a.x - b.y, a.y + b.x);
        if (k != 0 && k != N / 2) z[N - k] = make_float2(a.x + b.y, b.x - a.y);
      }
    }
    __syncthreads();
    float2 u[20], v[16];
    if (t < 16) {
#pragma unroll
      for (int k2 = 0; k2 < 20; ++k2) u[k2] = zb[pr * R20_BP + t + 16 * k2];
      fft20<1>(u);                                  // u[n1] = C_t[n1]
    }
    __syncthreads();                                // Z has been read: its memory becomes the exchange
    if (t < 16) {
#pragma unroll
      for (int n1 = 0; n1 < 20; ++n1) ex[(n1 * 16 + t) * R20_PR + pr] = cmul(u[n1], tws[t * 20 + n1]);
    }
    __syncthreads();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = ex[(t * 16 + k1) * R20_PR + pr];
    fft16<1>(v);                                    // v[n2] = z[t + 20 n2]: row 2 pr in .x, row 2 pr + 1 in .y
    __syncthreads();                                // the exchange has been read: its memory becomes the 32 rows
    float* r1 = band + (2 * pr) * R20_RBP;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      const int c = t + 20 * n2;
      r1[c + (c >> 5)] = v[n2].x;
      r1[R20_RBP + c + (c >> 5)] = v[n2].y;
    }
    __syncthreads();
    for (int row = warp; row < 32; row += NW) {
      const float* ln = band + row * R20_RBP;
      const float xm = (float)(b0 + row) - half + 1.0f;
      const float lim = r2lim - xm * xm;            // keep the pixel iff y * y < lim
      const int c0 = lane * E;
      float sv[E];
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const float y = (float)(c0 + j) - half;
        const float val = ln[c0 + j + ((c0 + j) >> 5)];
        sv[j] = (y * y < lim) ? 6.0f * val : 0.0f;
      }
      spline_line_warp<E>(sv, zE, lane);
      float2* orow = reinterpret_cast<float2*>(dst + (b0 + row) * N + c0);
#pragma unroll
      for (int j = 0; j < E / 2; ++j) orow[j] = make_float2(sv[2 * j], sv[2 * j + 1]);
    }
    __syncthreads();                                // the rows are overwritten by the next band's Z
  }
}

bool colfilter_supported(int N) { return N == 256 || N == 128; }
bool colpass_supported(int N) { return N == 320; }

static int ensure_twiddles(mem_ctx* ctx, cudaStream_t st) {
  static thread_local int tw_device = -1;
  if (tw_device != ctx->device) {     // constant memory is per device; every host thread checks its own context
    float2 tw[256];
    for (int j = 0; j < 256; ++j) {
      const double a = -2.0 * M_PI * j / 256.0;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    MEM_CUDA(cudaMemcpyToSymbolAsync(c_tw256, tw, sizeof(tw), 0, cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaStreamSynchronize(st));
    tw_device = ctx->device;
  }
  return 0;
}

static int ensure_twiddles320(mem_ctx* ctx, cudaStream_t st) {
  static thread_local int tw_device = -1;
  if (tw_device != ctx->device) {
    float2 tw[320];
    for (int j = 0; j < 320; ++j) {
      const double a = -2.0 * M_PI * j / 320.0;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    MEM_CUDA(cudaMemcpyToSymbolAsync(c_tw320, tw, sizeof(tw), 0, cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaStreamSynchronize(st));
    tw_device = ctx->device;
  }
  return 0;
}

// column pass alone (N = 320): spec holds row-transformed half spectra; fwd_only = the a10 transform; stats != NULL: the rows
// came from k_ingest_rowfft320 (not yet normalised)
int colpass_run(mem_ctx* ctx, float2* spec, const float* G, const float2* stats, int nS, int N, int fwd_only, cudaStream_t st) {
  if (!colpass_supported(N)) {
    set_error("colpass: no kernel for N = %d", N);
    return 1;
  }
  MEM_CHECK(ensure_twiddles320(ctx, st));
  const int per_slab = std::max(1, std::min(nS, (2 * ctx->sm_count) / CF20_SLABS));
  const size_t smem = 2 * 320 * CF20_COLS * sizeof(float2);
  if (fwd_only) {
    MEM_CUDA(cudaFuncSetAttribute(k_colfilter320<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MEM_LAUNCH(ctx, k_colfilter320<true>, per_slab * CF20_SLABS, CF20_THREADS, smem, st, spec, G, stats, nS, per_slab);
  } else {
    MEM_CUDA(cudaFuncSetAttribute(k_colfilter320<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MEM_LAUNCH(ctx, k_colfilter320<false>, per_slab * CF20_SLABS, CF20_THREADS, smem, st, spec, G, stats, nS, per_slab);
  }
  return 0;
}

// row passes at N = 320 (see k_ingest_rowfft320 / k_rowifft_prefilter320)
int rows320_forward_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float2* spec, float2* stats, int nS, int transposed,
                        int plain, cudaStream_t st) {
  MEM_CHECK(ensure_twiddles320(ctx, st));
  if (plain) MEM_LAUNCH(ctx, k_ingest_rowfft320<true>, nS, R20_THREADS, R20_SMEM, st, raw, flip, spec, stats, transposed);
  else MEM_LAUNCH(ctx, k_ingest_rowfft320<false>, nS, R20_THREADS, R20_SMEM, st, raw, flip, spec, stats, transposed);
  return 0;
}

int rows320_inverse_run(mem_ctx* ctx, const float2* spec, float* out, int nS, cudaStream_t st) {
  MEM_CHECK(ensure_twiddles320(ctx, st));
  MEM_LAUNCH(ctx, k_rowifft_prefilter320, nS, R20_THREADS, R20_SMEM, st, spec, out);
  return 0;
}

int ingest_rowfft_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float2* spec, float2* stats, int nS, int N,
                      int transposed, cudaStream_t st) {
  if (!colfilter_supported(N)) {
    set_error("ingest_rowfft: no kernel for N = %d", N);
    return 1;
  }
  MEM_CHECK(ensure_twiddles(ctx, st));
  if (N == 128) {
    const size_t smem = 64 * IR8_BP * sizeof(float);
    MEM_LAUNCH(ctx, k_ingest_rowfft128<false>, nS, 256, smem, st, raw, flip, spec, stats, transposed);
    return 0;
  }
  const size_t smem = 32 * IR_BP * sizeof(float);
  auto kern = k_ingest_rowfft256<4>;
  if (ctx->rowfft_blocks == 5) kern = k_ingest_rowfft256<5>;
  else if (ctx->rowfft_blocks == 6) kern = k_ingest_rowfft256<6>;
  else if (ctx->rowfft_blocks == 3) kern = k_ingest_rowfft256<3>;
  MEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEM_LAUNCH(ctx, kern, nS, 256, smem, st, raw, flip, spec, stats, transposed);
  return 0;
}

int rowifft_prefilter_run(mem_ctx* ctx, const float2* spec, float* out, int nS, int N, cudaStream_t st) {
  if (!colfilter_supported(N)) {
    set_error("rowifft_prefilter: no kernel for N = %d", N);
    return 1;
  }
  MEM_CHECK(ensure_twiddles(ctx, st));
  if (N == 128) {
    const size_t smem = 64 * RP8_BP * sizeof(float);
    MEM_LAUNCH(ctx, k_rowifft_prefilter128, nS, 256, smem, st, spec, out);
    return 0;
  }
  const size_t smem = 32 * RP_BP * sizeof(float);
  auto kern = k_rowifft_prefilter256<4>;
  if (ctx->rowfft_blocks == 5) kern = k_rowifft_prefilter256<5>;
  else if (ctx->rowfft_blocks == 6) kern = k_rowifft_prefilter256<6>;
  else if (ctx->rowfft_blocks == 3) kern = k_rowifft_prefilter256<3>;
  MEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEM_LAUNCH(ctx, kern, nS, 256, smem, st, spec, out);
  return 0;
}

template <bool FWD>
static int colpass128(mem_ctx* ctx, float2* spec, const float* G, const float2* stats, int nS, cudaStream_t st) {
  static_assert(CF8_SLABS * CF8_COLS >= 65, "slabs must cover the half spectrum");
  const int per_slab = std::max(1, std::min(nS, (4 * ctx->sm_count) / CF8_SLABS));   // four CTAs per SM
  const size_t smem = 2 * 128 * CF8_COLS * sizeof(float2);
  MEM_LAUNCH(ctx, k_colfilter128<FWD>, per_slab * CF8_SLABS, CF8_THREADS, smem, st, spec, G, stats, 65, nS, per_slab);
  return 0;
}

int colfilter_run(mem_ctx* ctx, float2* spec, const float* G, const float2* stats, int nS, int N, cudaStream_t st) {
  if (!colfilter_supported(N)) {
    set_error("colfilter: no kernel for N = %d", N);
    return 1;
  }
  MEM_CHECK(ensure_twiddles(ctx, st));
  if (N == 128) return colpass128<false>(ctx, spec, G, stats, nS, st);
  const int Nh = N / 2 + 1;
  static_assert(CF_SLABS * CF_COLS >= 129, "slabs must cover the half spectrum");
  const int per_slab = std::max(1, std::min(nS, (2 * ctx->sm_count) / CF_SLABS));   // CTAs per slab, two CTAs per SM
  const size_t smem = 2 * 256 * CF_COLS * sizeof(float2);
  MEM_CUDA(cudaFuncSetAttribute(k_colfilter256<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEM_LAUNCH(ctx, k_colfilter256<false>, per_slab * CF_SLABS, CF_THREADS, smem, st, spec, G, stats, Nh, nS, per_slab);
  return 0;
}

// a10 (:344) with our own kernels: real rows -> half spectra (the PLAIN row kernel), then the forward column pass
int fft2_forward_run(mem_ctx* ctx, const float* img, float2* spec, int nS, int N, cudaStream_t st) {
  if (!colfilter_supported(N)) {
    set_error("fft2_forward: no kernel for N = %d", N);
    return 1;
  }
  MEM_CHECK(ensure_twiddles(ctx, st));
  if (N == 128) {
    const size_t smem = 64 * IR8_BP * sizeof(float);
    MEM_LAUNCH(ctx, k_ingest_rowfft128<true>, nS, 256, smem, st, img, (const uint8_t*)nullptr, spec, (float2*)nullptr, 0);
    return colpass128<true>(ctx, spec, nullptr, nullptr, nS, st);
  }
  const size_t smem_r = 32 * IR_BP * sizeof(float);
  auto kr = k_ingest_rowfft256<4, true>;
  MEM_CUDA(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
  MEM_LAUNCH(ctx, kr, nS, 256, smem_r, st, img, (const uint8_t*)nullptr, spec, (float2*)nullptr, 0);
  const int Nh = N / 2 + 1;
  const int per_slab = std::max(1, std::min(nS, (2 * ctx->sm_count) / CF_SLABS));
  const size_t smem_c = 2 * 256 * CF_COLS * sizeof(float2);
  MEM_CUDA(cudaFuncSetAttribute(k_colfilter256<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  MEM_LAUNCH(ctx, k_colfilter256<true>, per_slab * CF_SLABS, CF_THREADS, smem_c, st, spec, (const float*)nullptr,
             (const float2*)nullptr, Nh, nS, per_slab);
  return 0;
}

}  // namespace mem
