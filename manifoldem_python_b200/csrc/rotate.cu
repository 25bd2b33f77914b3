// rotate.cu — the two cubic-B-spline rotations of the alignment (row a7, rotatefill.py:21-41) for boxes that are a
// multiple of 32: out(o) = sum_{4x4} w * coef[(floor(x) - 1 + a) mod N], x = R (o - ctr) + ctr, ctr = (N - 1) / 2.
// Same arithmetic, operation for operation, as the generic k_rotate of align.cu (which stays for the other boxes);
// what changes is how the data moves:
//
//   k_rotate_img   first rotation, one angle per image.  The 50 x 56 source window of a 32 x 32 output tile is staged
//                  with 16-byte loads and stores (columns aligned down to a multiple of 4; the old kernel spent half
//                  of its instructions on 14 scalar loads per thread and their wrap arithmetic), and the tile pitch is
//                  picked per image from {64, 72, 76} floats by its angle (table built on the host by simulating the
//                  bank pattern of the 32 taps of a warp): 1.84 instead of 2.18 wavefronts per load.  Measured (ncu):
//                  the LSU data pipe is 96 % busy — 121 M wavefronts of tap loads, 16 M of staging stores, 31 M on the
//                  global side — so the kernel sits at the floor of this decomposition (a warp's 32 taps lie on a line
//                  of slope (sin, cos) spanning up to 45 diagonals of the bank grid).  An 8 x 4 pixel patch per warp
//                  lowers the tap wavefronts to 1.73 per load but its 4-segment stores cost more than that on the
//                  global side of the same pipe (measured 634 us against 626).
//   k_rotate_common  second rotation: ONE angle (-psi_p, :330) for every image of the PD, so tap positions and weights
//                  are computed once per pixel and shared by four images whose coefficients are interleaved in shared
//                  memory as float4 — one LDS.128 fetches a tap of four images, the index / weight arithmetic is
//                  amortised over them.  An LDS.128 is served per quarter-warp (8 lanes, 8 groups of 4 banks); the
//                  float4 pitch (57..64) is chosen per launch by simulating that pattern for the launch's angle.
#include "common.cuh"

#include <math.h>
#include <algorithm>
#include <stdlib.h>

namespace mem {

constexpr int RT = 32;          // output tile
constexpr int RROWS = 50;       // source rows of a tile: ceil(31 * sqrt(2)) + 4 taps + 1
constexpr int RG = 14;          // 16-byte column groups: 50 + 3 (alignment) -> 56 columns

__device__ __forceinline__ void bspline_w4(float t, float (&w)[4]) {
  const float u = 1.0f - t;
  const float t2 = t * t, u2 = u * u;
  w[0] = u2 * u * (1.0f / 6.0f);
  w[3] = t2 * t * (1.0f / 6.0f);
  w[1] = fmaf(t2, fmaf(0.5f, t, -1.0f), 2.0f / 3.0f);
  w[2] = fmaf(u2, fmaf(0.5f, u, -1.0f), 2.0f / 3.0f);
}

// per-image cos / sin in fp64 (ndimage.rotate: matrix [[c, s], [-s, c]], angle in degrees); entry nS holds the common
// second rotation by -psi_p (:330); pid[i] = index of the tile pitch for the first rotation of image i
__global__ void k_angles2(const double* __restrict__ psi_deg, double psi_p_deg, double2* __restrict__ cs,
                          uint8_t* __restrict__ pid, const uint8_t* __restrict__ pitch_of_deg, int nS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nS) return;
  const double deg = (i < nS ? psi_deg[i] : -psi_p_deg);
  const double a = deg * 0.017453292519943295769;
  double s, c;
  sincos(a, &s, &c);
  cs[i] = make_double2(c, s);
  if (i < nS) {
    double m = fmod(deg, 360.0);
    if (m < 0) m += 360.0;
    int bin = (int)(m + 0.5);
    if (bin >= 360 || bin < 0) bin = 0;       // NaN angles land in bin 0 too
    pid[i] = pitch_of_deg[bin];
  }
}

// batched PDs: the second rotation of image i is by -psi_p of ITS PD (pd_of[i]); cs2 / pid2 per image, so the per-image
// kernel k_rotate_img serves both rotations of a batch
__global__ void k_angles_batch(const int* __restrict__ pd_of, const double* __restrict__ psi_p_deg, double2* __restrict__ cs2,
                               uint8_t* __restrict__ pid2, const uint8_t* __restrict__ pitch_of_deg, int nS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nS) return;
  const double deg = -psi_p_deg[pd_of[i]];
  double s, c;
  sincos(deg * 0.017453292519943295769, &s, &c);
  cs2[i] = make_double2(c, s);
  double m = fmod(deg, 360.0);
  if (m < 0) m += 360.0;
  int bin = (int)(m + 0.5);
  if (bin >= 360 || bin < 0) bin = 0;
  pid2[i] = pitch_of_deg[bin];
}

// bounding-box origin of the source footprint of the tile at (r0, c0): floor(min over the corners) - 1
__device__ __forceinline__ void tile_origin(const double2 a, int N, int r0, int c0, int& o0, int& o1) {
  const double ctr = 0.5 * (N - 1);
  const double tr = r0 - ctr, tc = c0 - ctr, ext = RT - 1;
  const double b0 = a.x * tr + a.y * tc + ctr + fmin(a.x * ext, 0.0) + fmin(a.y * ext, 0.0);
  const double b1 = -a.y * tr + a.x * tc + ctr + fmin(-a.y * ext, 0.0) + fmin(a.x * ext, 0.0);
  o0 = __double2int_rd(b0) - 1;
  o1 = __double2int_rd(b1) - 1;
}

__device__ __forceinline__ int wrap_once(int v, int N) {   // -N <= v < 2N  ->  [0, N)
  v += (v < 0) ? N : 0;
  v -= (v >= N) ? N : 0;
  return v;
}

template <int P>   // tile pitch in floats (multiple of 4)
__device__ __forceinline__ void rotate_img_tile(float* __restrict__ tile, const double2 a, const float* __restrict__ coef,
                                                float* __restrict__ out, int N) {
  const int img = blockIdx.z;
  const int r0 = blockIdx.y * RT, c0 = blockIdx.x * RT;
  int o0, o1;
  tile_origin(a, N, r0, c0, o0, o1);
  const int o1a = o1 & ~3;                              // 16-byte aligned first column (two's complement: floors)
  const float* src = coef + (size_t)img * N * N;
  {
    const int w0 = wrap_once(o0, N), w1 = wrap_once(o1a, N);
#pragma unroll
    for (int it = 0; it < (RROWS * RG + 255) / 256; ++it) {
      const int e = threadIdx.x + it * 256;
      if (e < RROWS * RG) {
        const int i = e / RG, g = e - i * RG;
        int gy = w0 + i;
        gy -= (gy >= N) ? N : 0;
        int gx = w1 + 4 * g;                              // N is a multiple of 4: a group never straddles the wrap
        gx -= (gx >= N) ? N : 0;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + gy * N + gx));
        *reinterpret_cast<float4*>(tile + i * P + 4 * g) = v;
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int dc = lane, dr = wrp;                         // a warp = 32 pixels of a row (one 128-byte store), rows 8 apart
  const double ctr = 0.5 * (N - 1);
  const double rr = (r0 + dr) - ctr, cc = (c0 + dc) - ctr;
  const double xb0 = a.x * rr + a.y * cc + ctr - (double)o0;      // row coordinate inside the staged window
  const double xb1 = -a.y * rr + a.x * cc + ctr - (double)o1a;    // column coordinate
  // 6.26 unsigned fixed point (coordinates are in [1, 52)); steps modulo 2^32, so negative steps need no special case
  unsigned int x0 = (unsigned int)__double2ll_rn(xb0 * 67108864.0), x1 = (unsigned int)__double2ll_rn(xb1 * 67108864.0);
  const unsigned int DX0 = (unsigned int)__double2ll_rn(8.0 * a.x * 67108864.0);
  const unsigned int DX1 = (unsigned int)__double2ll_rn(-8.0 * a.y * 67108864.0);
  float* dst = out + (size_t)img * N * N + (r0 + dr) * N + c0 + dc;
  const int rstep = 8 * N;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i0 = (int)(x0 >> 26), j0 = (int)(x1 >> 26);
    const float t0 = (float)(x0 & 0x3ffffffu) * 1.4901161193847656e-08f;
    const float t1 = (float)(x1 & 0x3ffffffu) * 1.4901161193847656e-08f;
    float wa[4], wb[4];
    bspline_w4(t0, wa);
    bspline_w4(t1, wb);
    const float* p = tile + ((i0 - 1) * P + (j0 - 1));
    float acc = 0.0f;
#pragma unroll
    for (int ai = 0; ai < 4; ++ai) {
      float rs = wb[0] * p[0];
      rs = fmaf(wb[1], p[1], rs);
      rs = fmaf(wb[2], p[2], rs);
      rs = fmaf(wb[3], p[3], rs);
      acc = fmaf(wa[ai], rs, acc);
      p += P;
    }
    dst[q * rstep] = acc;
    x0 += DX0;
    x1 += DX1;
  }
}

constexpr int RP0 = 64, RP1 = 72, RP2 = 76;     // pitch candidates of the first rotation (floats)
__global__ void __launch_bounds__(256) k_rotate_img(const float* __restrict__ coef, float* __restrict__ out,
                                                    const double2* __restrict__ cs, const uint8_t* __restrict__ pid, int N) {
  __shared__ __align__(16) float tile[RROWS * RP2];
  const double2 a = cs[blockIdx.z];
  const int id = pid[blockIdx.z];
  if (id == 0) rotate_img_tile<RP0>(tile, a, coef, out, N);
  else if (id == 1) rotate_img_tile<RP1>(tile, a, coef, out, N);
  else rotate_img_tile<RP2>(tile, a, coef, out, N);
}

// second rotation: the same angle for every image; four images per CTA, coefficients interleaved as float4
template <int P4>   // tile pitch in float4 units
__global__ void __launch_bounds__(256) k_rotate_common(const float* __restrict__ coef, float* __restrict__ out,
                                                       const double2* __restrict__ cs_common, int N, int nS,
                                                       const uint8_t* __restrict__ msk2, float* __restrict__ out_masked,
                                                       int rrows, int rcols) {
  extern __shared__ __align__(16) float4 tile4[];
  const double2 a = cs_common[0];
  const int img0 = blockIdx.z * 4;
  const int r0 = blockIdx.y * RT, c0 = blockIdx.x * RT;
  const size_t NN = (size_t)N * N;
  int o0, o1;
  tile_origin(a, N, r0, c0, o0, o1);
  const int o1a = o1 & ~3;
  {
    // one column per thread: four coalesced 4-byte loads (one per image), ONE 16-byte store; consecutive lanes write
    // consecutive float4 (conflict-free).  Warp w stages rows w, w + 8, ...; lanes cover columns lane and lane + 32.
    const int w0 = wrap_once(o0, N), w1 = wrap_once(o1a, N);
    const float* s0 = coef + (size_t)img0 * NN;
    const size_t d1 = (size_t)(min(img0 + 1, nS - 1) - img0) * NN, d2 = (size_t)(min(img0 + 2, nS - 1) - img0) * NN,
                 d3 = (size_t)(min(img0 + 3, nS - 1) - img0) * NN;
    const int ln = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int gxa = w1 + ln;
    gxa -= (gxa >= N) ? N : 0;
    int gxb = w1 + ln + 32;
    gxb -= (gxb >= N) ? N : 0;
    gxb -= (gxb >= N) ? N : 0;
    const bool second = ln + 32 < rcols;                 // rcols <= 4 RG: the columns this launch's angle can reach
    int gy = w0 + wp;
    gy -= (gy >= N) ? N : 0;
    // rrows <= RROWS: the rows this launch's angle can reach (31 (|cos| + |sin|) + 6) — the shared-memory tile is sized by
    // it, so small angles fit more CTAs per SM (the kernel is bound by shared-memory latency at 4 CTAs)
#pragma unroll
    for (int it = 0; it < (RROWS + 7) / 8; ++it) {
      if (it * 8 + wp < rrows) {                          // warp-uniform
        const float* rowp = s0 + gy * N;
        float4* t = tile4 + (it * 8 + wp) * P4;
        t[ln] = make_float4(__ldg(rowp + gxa), __ldg(rowp + d1 + gxa), __ldg(rowp + d2 + gxa), __ldg(rowp + d3 + gxa));
        if (second)
          t[ln + 32] = make_float4(__ldg(rowp + gxb), __ldg(rowp + d1 + gxb), __ldg(rowp + d2 + gxb), __ldg(rowp + d3 + gxb));
      }
      gy += 8;
      gy -= (gy >= N) ? N : 0;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int dc = lane, dr = wrp;       // a warp = 32 pixels of a row; a quarter-warp (one LDS.128 phase) = 8 of them
  const double ctr = 0.5 * (N - 1);
  const double rr = (r0 + dr) - ctr, cc = (c0 + dc) - ctr;
  const double xb0 = a.x * rr + a.y * cc + ctr - (double)o0;
  const double xb1 = -a.y * rr + a.x * cc + ctr - (double)o1a;
  unsigned int x0 = (unsigned int)__double2ll_rn(xb0 * 67108864.0), x1 = (unsigned int)__double2ll_rn(xb1 * 67108864.0);
  const unsigned int DX0 = (unsigned int)__double2ll_rn(8.0 * a.x * 67108864.0);
  const unsigned int DX1 = (unsigned int)__double2ll_rn(-8.0 * a.y * 67108864.0);
  const int oidx = (r0 + dr) * N + c0 + dc;
  const int rstep = 8 * N;
  const int nimg = min(4, nS - img0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i0 = (int)(x0 >> 26), j0 = (int)(x1 >> 26);
    const float t0 = (float)(x0 & 0x3ffffffu) * 1.4901161193847656e-08f;
    const float t1 = (float)(x1 & 0x3ffffffu) * 1.4901161193847656e-08f;
    float wa[4], wb[4];
    bspline_w4(t0, wa);
    bspline_w4(t1, wb);
    const float4* p = tile4 + ((i0 - 1) * P4 + (j0 - 1));
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ai = 0; ai < 4; ++ai) {
      const float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
      float rs;
      rs = wb[0] * v0.x; rs = fmaf(wb[1], v1.x, rs); rs = fmaf(wb[2], v2.x, rs); rs = fmaf(wb[3], v3.x, rs);
      acc[0] = fmaf(wa[ai], rs, acc[0]);
      rs = wb[0] * v0.y; rs = fmaf(wb[1], v1.y, rs); rs = fmaf(wb[2], v2.y, rs); rs = fmaf(wb[3], v3.y, rs);
      acc[1] = fmaf(wa[ai], rs, acc[1]);
      rs = wb[0] * v0.z; rs = fmaf(wb[1], v1.z, rs); rs = fmaf(wb[2], v2.z, rs); rs = fmaf(wb[3], v3.z, rs);
      acc[2] = fmaf(wa[ai], rs, acc[2]);
      rs = wb[0] * v0.w; rs = fmaf(wb[1], v1.w, rs); rs = fmaf(wb[2], v2.w, rs); rs = fmaf(wb[3], v3.w, rs);
      acc[3] = fmaf(wa[ai], rs, acc[3]);
      p += P4;
    }
    const int o = oidx + q * rstep;
    const bool keep = msk2 ? (msk2[o] != 0) : true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nimg) {
        out[(size_t)(img0 + k) * NN + o] = acc[k];
        if (out_masked) out_masked[(size_t)(img0 + k) * NN + o] = keep ? acc[k] : 0.0f;
      }
    }
    x0 += DX0;
    x1 += DX1;
  }
}

// ------------------------------------------------------------------------------------------------
// host: bank-pattern simulation -> pitch choices
// ------------------------------------------------------------------------------------------------
// worst multiplicity of distinct addresses per bank for `n` lanes; banks = addr mod nb
static int wavefronts(const int* addr, int n, int nb) {
  int worst = 0;
  for (int l = 0; l < n; ++l) {
    int cnt = 0;
    for (int m = 0; m < n; ++m) {
      if (((addr[m] - addr[l]) % nb) != 0) continue;
      bool first = true;                                  // count distinct addresses of this bank once
      for (int k = 0; k < m; ++k)
        if (addr[k] == addr[m]) { first = false; break; }
      if (first) ++cnt;
    }
    worst = std::max(worst, cnt);
  }
  return worst;
}

// mean wavefronts per LDS.32 of a warp that computes 32 pixels of an output row, tile pitch P, rotation (c, s)
static double sim_img(double c, double s, int P) {
  static const double offs[6][2] = {{0.13, 0.71}, {0.52, 0.08}, {0.91, 0.44}, {0.27, 0.95}, {0.66, 0.31}, {0.40, 0.58}};
  double tot = 0;
  for (auto& o : offs) {
    int addr[32];
    for (int l = 0; l < 32; ++l) {
      const int i = (int)floor(30.0 + o[0] + l * s), j = (int)floor(30.0 + o[1] + l * c);
      addr[l] = i * P + j;
    }
    tot += wavefronts(addr, 32, 32);
  }
  return tot / 6;
}

// mean wavefronts per quarter-warp phase of an LDS.128: 8 lanes = 8 consecutive pixels of a row, float4 pitch P4; averaged over
// a 12 x 12 grid of sub-pixel offsets of the first lane (the measured kernel time follows this figure closely: at -40 degrees
// the pitches 55 and 63 take 2.34 ms where 58 takes 1.57)
static double sim_common(double c, double s, int P4) {
  double tot = 0;
  int n = 0;
  for (int a = 0; a < 12; ++a)
    for (int b = 0; b < 12; ++b) {
      const double o0 = (a + 0.37) / 12.0, o1 = (b + 0.61) / 12.0;
      int addr[8];
      for (int l = 0; l < 8; ++l) {
        const int i = (int)floor(30.0 + o0 + l * s), j = (int)floor(30.0 + o1 + l * c);
        addr[l] = i * P4 + j;
      }
      tot += wavefronts(addr, 8, 8);
      ++n;
    }
  return tot / n;
}

static int ensure_pitch_table(mem_ctx* ctx, cudaStream_t st) {
  if (ctx->rot_pitch_tab.p) return 0;
  uint8_t tab[360];
  const int P[3] = {RP0, RP1, RP2};
  for (int d = 0; d < 360; ++d) {
    const double a = d * M_PI / 180.0;
    double best = 1e9;
    int bi = 0;
    for (int k = 0; k < 3; ++k) {
      const double w = sim_img(cos(a), sin(a), P[k]);
      if (w < best - 1e-9) { best = w; bi = k; }
    }
    tab[d] = (uint8_t)bi;
  }
  MEM_CHECK(ctx->rot_pitch_tab.ensure(360));
  MEM_CUDA(cudaMemcpyAsync(ctx->rot_pitch_tab.p, tab, 360, cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

bool rotate_fast_supported(int N) { return N % RT == 0 && N >= 64; }

// angles + pitch ids (cs [nS + 1] double2, pid [nS] bytes)
int rotate_angles_run(mem_ctx* ctx, const double* psi_deg, double psi_p_deg, double2* cs, uint8_t* pid, int nS,
                      cudaStream_t st) {
  MEM_CHECK(ensure_pitch_table(ctx, st));
  MEM_LAUNCH(ctx, k_angles2, (nS + 1 + 127) / 128, 128, 0, st, psi_deg, psi_p_deg, cs, pid,
             ctx->rot_pitch_tab.as<uint8_t>(), nS);
  return 0;
}

int rotate_angles_batch_run(mem_ctx* ctx, const int* pd_of, const double* psi_p_deg, double2* cs2, uint8_t* pid2, int nS,
                            cudaStream_t st) {
  MEM_CHECK(ensure_pitch_table(ctx, st));
  MEM_LAUNCH(ctx, k_angles_batch, (nS + 127) / 128, 128, 0, st, pd_of, psi_p_deg, cs2, pid2, ctx->rot_pitch_tab.as<uint8_t>(), nS);
  return 0;
}

int rotate_img_run(mem_ctx* ctx, const float* coef, float* out, const double2* cs, const uint8_t* pid, int nS, int N,
                   cudaStream_t st) {
  const dim3 grid(N / RT, N / RT, nS);
  MEM_LAUNCH(ctx, k_rotate_img, grid, 256, 0, st, coef, out, cs, pid, N);
  return 0;
}

int rotate_common_run(mem_ctx* ctx, const float* coef, float* out, const double2* cs_common, double angle_deg, int nS,
                      int N, const uint8_t* msk2, float* out_masked, cudaStream_t st) {
  const double a = angle_deg * M_PI / 180.0;
  // the quarter-warp of an LDS.128 walks along an output row: source step (sin, cos) per lane.  The tile holds only the rows and
  // columns this angle can reach (31 (|cos| + |sin|) + taps + alignment); the float4 pitch is the best of the candidates
  // from that width up to 64, by simulating the bank pattern — small tiles mean more CTAs per SM (4 at 45 degrees with the old fixed
  // 50 x 58 tile, 5 now; 8 near the axes)
  const double ext = (RT - 1) * (fabs(cos(a)) + fabs(sin(a)));
  const int rrows = std::min(RROWS, (int)floor(ext + 1e-9) + 6);
  const int rcols = std::min(4 * RG, (int)floor(ext + 1e-9) + 6 + 3);
  double best = 1e9;
  int P4 = 64;
  for (int cand = std::max(40, rcols); cand <= 64; ++cand) {
    const double w = sim_common(cos(a), sin(a), cand);
    if (w < best - 1e-9) { best = w; P4 = cand; }
  }
  if (const char* ov = getenv("MANIFOLDEM_B200_ROT_P4")) {              // experiments: force the pitch
    const int v = atoi(ov);
    if (v >= rcols && v <= 64) P4 = v;
  }
  const dim3 grid(N / RT, N / RT, (nS + 3) / 4);
  const size_t smem = (size_t)rrows * P4 * sizeof(float4);
#define LAUNCH_COMMON(PP)                                                                                   \
  case PP:                                                                                                  \
    MEM_CUDA(cudaFuncSetAttribute(k_rotate_common<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    MEM_LAUNCH(ctx, k_rotate_common<PP>, grid, 256, smem, st, coef, out, cs_common, N, nS, msk2, out_masked, rrows, rcols);  \
    break;
  switch (P4) {
    LAUNCH_COMMON(40) LAUNCH_COMMON(41) LAUNCH_COMMON(42) LAUNCH_COMMON(43) LAUNCH_COMMON(44) LAUNCH_COMMON(45)
    LAUNCH_COMMON(46) LAUNCH_COMMON(47) LAUNCH_COMMON(48) LAUNCH_COMMON(49) LAUNCH_COMMON(50) LAUNCH_COMMON(51)
    LAUNCH_COMMON(52) LAUNCH_COMMON(53) LAUNCH_COMMON(54) LAUNCH_COMMON(55) LAUNCH_COMMON(56) LAUNCH_COMMON(57)
    LAUNCH_COMMON(58) LAUNCH_COMMON(59) LAUNCH_COMMON(60) LAUNCH_COMMON(61) LAUNCH_COMMON(62) LAUNCH_COMMON(63)
    default:
      MEM_CUDA(cudaFuncSetAttribute(k_rotate_common<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MEM_LAUNCH(ctx, k_rotate_common<64>, grid, 256, smem, st, coef, out, cs_common, N, nS, msk2, out_masked, rrows, rcols);
      break;
  }
#undef LAUNCH_COMMON
  return 0;
}

}  // namespace mem
