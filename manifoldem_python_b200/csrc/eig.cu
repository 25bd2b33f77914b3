// eig.cu — the eigen-solve of the diffusion map (row a19 / SURVEY §8f rank 1: sembeddingonFly.py:27,
// eigsh(L, k = nEigs + 1, maxiter = 300)) on the device: Lanczos with full re-orthogonalisation on the resident
// float64 Laplacian.  The reference's ARPACK call applies L once per iteration from the host; here the Krylov basis,
// the three-term recurrence and both Gram-Schmidt passes stay on the device and a BLOCK of steps is enqueued without
// a host round trip — the host only looks at the 2 x j recurrence coefficients between blocks (Ritz values of the
// tridiagonal matrix, residual estimates |beta_j s_ji|) and asks for the Ritz vectors at the end.
//
//   V      [m_max + 1][nS] float64   Krylov basis, one vector per row (v_0 = normalised deterministic start vector)
//   ab     [2][m_max + 1]  float64   alpha_j = ab[0][j];  beta_j = ab[1][j] = ||w|| after step j - 1 (ab[1][0] unused)
//
// Step j:  w = L v_j;  h = V_{0..j}^T w;  alpha_j = h_j;  w -= V h;   (classical Gram-Schmidt, twice)
//          h' = V^T w; alpha_j += h'_j;   w -= V h';  beta_{j+1} = ||w||;  v_{j+1} = w / beta_{j+1}.
// Every reduction runs in a fixed order (deterministic results).  All HBM-bound: L is read once per step
// (8 nS^2 B; 32 MB at nS = 2,000 stays in the 126 MB L2 between steps), the basis twice per pass.
#include "common.cuh"

namespace mem {

__device__ __forceinline__ double block_sum_256(double v, double* red) {   // fixed-order block reduction, 256 threads
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k];
  return s;
}

// start vector: uniform(-1, 1) from a counter hash (splitmix64) — deterministic, no structure shared with L
__global__ void k_lz_start(double* __restrict__ w, int nS, unsigned long long seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nS) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  w[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// y = L x, one warp per row (L symmetric, row-major), fixed summation order
__global__ void __launch_bounds__(256) k_lz_symv(const double* __restrict__ L, const double* __restrict__ x,
                                                 double* __restrict__ y, int nS) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nS) return;
  const double* r = L + (size_t)row * nS;
  double a0 = 0.0, a1 = 0.0;
  int j = lane;
  for (; j + 32 < nS; j += 64) {
    a0 = fma(r[j], x[j], a0);
    a1 = fma(r[j + 32], x[j + 32], a1);
  }
  if (j < nS) a0 = fma(r[j], x[j], a0);
  double acc = a0 + a1;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

// h[b] = V[b] . w for b < nb: one CTA per basis vector
__global__ void __launch_bounds__(256) k_lz_dots(const double* __restrict__ V, const double* __restrict__ w,
                                                 double* __restrict__ h, int nS) {
  __shared__ double red[8];
  const double* v = V + (size_t)blockIdx.x * nS;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nS; i += 256) acc = fma(v[i], w[i], acc);
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) h[blockIdx.x] = acc;
}

// w -= sum_b h[b] V[b];  alpha[j] (+)= h[j].  CTA = 32 elements x 8 groups of basis vectors (group g sums b = g, g + 8, ...):
// at nS ~ 1,000 a thread per element alone is four CTAs walking 150 dependent loads each (22 us per launch, 62 % of the solver).
__global__ void __launch_bounds__(256) k_lz_update(double* __restrict__ w, const double* __restrict__ V,
                                                   const double* __restrict__ h, int nb, int nS,
                                                   double* __restrict__ alpha_j, int accumulate) {
  __shared__ double part[8][33];
  const int lx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lx;
  if (blockIdx.x == 0 && threadIdx.x == 0) alpha_j[0] = (accumulate ? alpha_j[0] : 0.0) + h[nb - 1];
  double a0 = 0.0, a1 = 0.0;
  if (i < nS) {
    int b = g;
    for (; b + 8 < nb; b += 16) {
      a0 = fma(h[b], V[(size_t)b * nS + i], a0);
      a1 = fma(h[b + 8], V[(size_t)(b + 8) * nS + i], a1);
    }
    if (b < nb) a0 = fma(h[b], V[(size_t)b * nS + i], a0);
  }
  part[g][lx] = a0 + a1;
  __syncthreads();
  if (g == 0 && i < nS) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][lx];
    w[i] -= s;
  }
}

// beta = ||w||, vnext = w / beta (one CTA)
__global__ void __launch_bounds__(1024) k_lz_normalize(const double* __restrict__ w, double* __restrict__ vnext,
                                                       double* __restrict__ beta_out, int nS) {
  __shared__ double red[32];
  __shared__ double tot;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nS; i += 1024) acc = fma(w[i], w[i], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < 32; ++k) s += red[k];
    tot = sqrt(s);
    if (beta_out) beta_out[0] = tot;
  }
  __syncthreads();
  const double inv = 1.0 / tot;
  for (int i = threadIdx.x; i < nS; i += 1024) vnext[i] = w[i] * inv;
}

// Ritz vectors: X[c][i] = sum_b S[b][c] V[b][i], c < k (k <= 32), S [j][k] row-major on the device
template <int KMAX>
__global__ void __launch_bounds__(128) k_lz_ritz(const double* __restrict__ V, const double* __restrict__ S, int j, int k,
                                                 int nS, double* __restrict__ X) {
  extern __shared__ double sS[];                 // [chunk][k]
  const int i = blockIdx.x * 128 + threadIdx.x;
  double acc[KMAX];
#pragma unroll
  for (int c = 0; c < KMAX; ++c) acc[c] = 0.0;
  constexpr int CH = 64;
  for (int b0 = 0; b0 < j; b0 += CH) {
    const int nb = min(CH, j - b0);
    __syncthreads();
    for (int t = threadIdx.x; t < nb * k; t += 128) sS[t] = S[(size_t)b0 * k + t];
    __syncthreads();
    if (i < nS) {
      for (int b = 0; b < nb; ++b) {
        const double v = V[(size_t)(b0 + b) * nS + i];
#pragma unroll
        for (int c = 0; c < KMAX; ++c)
          if (c < k) acc[c] = fma(sS[b * k + c], v, acc[c]);
      }
    }
  }
  if (i < nS) {
#pragma unroll
    for (int c = 0; c < KMAX; ++c)
      if (c < k) X[(size_t)c * nS + i] = acc[c];
  }
}

// Lanczos steps j0 .. j1 - 1 (j0 == 0 also builds v_0).  Everything is enqueued on `st`; no synchronisation.
int lanczos_steps_device(mem_ctx* ctx, const double* L, int nS, double* V, double* ab, int ld_ab, int j0, int j1,
                         cudaStream_t st) {
  if (nS < 2 || j0 < 0 || j1 < j0 || j1 >= ld_ab) {
    set_error("lanczos: need nS >= 2 and 0 <= j0 <= j1 < ld_ab (nS=%d j0=%d j1=%d ld_ab=%d)", nS, j0, j1, ld_ab);
    return 1;
  }
  MEM_CHECK(ctx->scratch.ensure((size_t)(nS + ld_ab + 8) * sizeof(double)));
  double* w = ctx->scratch.as<double>();
  double* h = w + nS;
  double* alpha = ab;
  double* beta = ab + ld_ab;
  const int gN = (nS + 255) / 256;
  if (j0 == 0) {
    MEM_CUDA(cudaMemsetAsync(ab, 0, (size_t)2 * ld_ab * sizeof(double), st));
    MEM_LAUNCH(ctx, k_lz_start, gN, 256, 0, st, w, nS, 0x5DEECE66Dull);
    MEM_LAUNCH(ctx, k_lz_normalize, 1, 1024, 0, st, w, V, (double*)nullptr, nS);
  }
  for (int j = j0; j < j1; ++j) {
    const double* vj = V + (size_t)j * nS;
    MEM_LAUNCH(ctx, k_lz_symv, (nS + 7) / 8, 256, 0, st, L, vj, w, nS);
    for (int pass = 0; pass < 2; ++pass) {
      MEM_LAUNCH(ctx, k_lz_dots, j + 1, 256, 0, st, V, w, h, nS);
      MEM_LAUNCH(ctx, k_lz_update, (nS + 31) / 32, 256, 0, st, w, V, h, j + 1, nS, alpha + j, pass);
    }
    MEM_LAUNCH(ctx, k_lz_normalize, 1, 1024, 0, st, w, V + (size_t)(j + 1) * nS, beta + j + 1, nS);
  }
  return 0;
}

// X [k][nS] (device) = S^T V with S [j][k] float64 on the HOST (the Ritz coefficients of the wanted pairs)
int lanczos_ritz_device(mem_ctx* ctx, const double* V, int nS, int j, const double* S_host, int k, double* X,
                        cudaStream_t st) {
  if (k < 1 || k > 32 || j < 1) {
    set_error("lanczos_ritz: 1 <= k <= 32 and j >= 1 (k=%d j=%d)", k, j);
    return 1;
  }
  MEM_CHECK(ctx->small_out.ensure((size_t)j * k * sizeof(double)));
  double* dS = ctx->small_out.as<double>();
  MEM_CUDA(cudaMemcpyAsync(dS, S_host, (size_t)j * k * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));                       // S_host may be a temporary of the caller
  const size_t smem = (size_t)64 * k * sizeof(double);
  if (k <= 16) MEM_LAUNCH(ctx, k_lz_ritz<16>, (nS + 127) / 128, 128, smem, st, V, dS, j, k, nS, X);
  else MEM_LAUNCH(ctx, k_lz_ritz<32>, (nS + 127) / 128, 128, smem, st, V, dS, j, k, nS, X);
  return 0;
}

}  // namespace mem
