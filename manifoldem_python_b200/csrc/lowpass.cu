// lowpass.cu — column pass of the a5 low-pass (getDistanceCTF_local_Conj9combinedS2.py:286-293),
//   img <- Re ifft2( fft2(img) * ifftshift(G) ).
// The transform is separable, the filter is not: rows go through cuFFT (1-D R2C / C2R), and everything that
// happens along ky — forward FFT, multiplication by G(ky,kx)/N^2, inverse FFT — is done here in ONE kernel, so
// the half spectrum crosses HBM once (read + write) instead of three times (2-D plan column pass, scale kernel,
// 2-D plan column pass).  HBM-bound: 16 * N * Nh bytes per image.
//
// N = 256 = 16 x 16.  A CTA owns COLS adjacent kx columns of one image, 16 threads per column.
//   forward:  thread t loads rows t + 16 m (m = 0..15) of its column straight from global memory (lanes run along kx:
//             coalesced), FFT-16 over m in registers, twiddle W256^(t k1), exchange through shared memory,
//             FFT-16 over t: the thread acting as k1 now holds X[k1 + 16 k2], k2 = 0..15.
//   filter:   X *= G  (table row-major like the spectrum: coalesced).
//   inverse:  X[k1 + 16 k2] is again a "residue + 16 m" set, so the inverse starts from the registers:
//             inverse FFT-16 over k2, conjugate twiddle, exchange, inverse FFT-16, store rows t + 16 m.
// Shared memory is touched only by the two exchanges; all its accesses are unit-stride across lanes.
#include "common.cuh"

#include <math.h>
#include <algorithm>

namespace mem {

constexpr int CF_COLS = 22;                 // 6 slabs cover Nh = 129 (3 idle columns in the last slab)
constexpr int CF_THREADS = 16 * CF_COLS;

__constant__ float2 c_tw256[256];           // exp(-2 pi i j / 256)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// 4-point DFT, exponent sign S: (a, b, c, d) <- (X0, X1, X2, X3)
template <int S>
__device__ __forceinline__ void fft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = make_float2(a.x + c.x, a.y + c.y), s1 = make_float2(a.x - c.x, a.y - c.y);
  const float2 s2 = make_float2(b.x + d.x, b.y + d.y), s3 = make_float2(b.x - d.x, b.y - d.y);
  a = make_float2(s0.x + s2.x, s0.y + s2.y);
  c = make_float2(s0.x - s2.x, s0.y - s2.y);
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
__global__ void __launch_bounds__(CF_THREADS, 2) k_colfilter256(float2* __restrict__ spec, const float* __restrict__ G,
                                                             int Nh, int nS, int img_stride) {
  extern __shared__ float2 cf_smem[];
  float2* ex = cf_smem;                           // [256][CF_COLS] exchange buffer
  float2* stage = cf_smem + 256 * CF_COLS;        // [256][CF_COLS] next image's slab
  const int t = threadIdx.x / CF_COLS, col = threadIdx.x - t * CF_COLS;
  const int slab = blockIdx.x % CF_SLABS;
  const int kx = slab * CF_COLS + col;
  const bool live = kx < Nh;
  const int kxc = live ? kx : 0;
  float gk[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) gk[k2] = live ? G[(t + 16 * k2) * Nh + kxc] : 0.0f;
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
      const float2 w = c_tw256[(t * k1) & 255];
      ex[(k1 * 16 + t) * CF_COLS + col] = cmul(v[k1], w);
    }
    __syncthreads();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = ex[(t * 16 + n2) * CF_COLS + col];
    fft16<-1>(v);                                   // v[k2] = X[t + 16 k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      v[k2].x *= gk[k2];
      v[k2].y *= gk[k2];
    }
    fft16<1>(v);                                    // inverse over k2
    __syncthreads();                                // every thread is done reading the forward exchange
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float2 w = c_tw256[(t * n1) & 255];
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

bool colfilter_supported(int N) { return N == 256; }

int colfilter_run(mem_ctx* ctx, float2* spec, const float* G, int nS, int N, cudaStream_t st) {
  if (N != 256) {
    set_error("colfilter: no kernel for N = %d", N);
    return 1;
  }
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
  const int Nh = N / 2 + 1;
  static_assert(CF_SLABS * CF_COLS >= 129, "slabs must cover the half spectrum");
  const int per_slab = std::max(1, std::min(nS, (2 * ctx->sm_count) / CF_SLABS));   // CTAs per slab, two CTAs per SM
  const size_t smem = 2 * 256 * CF_COLS * sizeof(float2);
  MEM_CUDA(cudaFuncSetAttribute(k_colfilter256, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEM_LAUNCH(ctx, k_colfilter256, per_slab * CF_SLABS, CF_THREADS, smem, st, spec, G, Nh, nS, per_slab);
  return 0;
}

}  // namespace mem
