// nlsa.cu — the NLSA / psi-analysis stage (modules/NLSA.py:23-158 with get_wiener.py:10-22,
// svdRF.py:19-32, L2_distance.py:33-41) on the arrays the distance stage leaves behind (D, imgAll, CTF).  All float64:
// the reference computes this stage in float64 and its outputs feed an SVD and a second diffusion map.
//
// What changes against the reference is the amount of work, not the result:
//   * NLSA.py:66-86 Wiener-filters ConOrder x (num - ConOrder) images one by one (an fft2 / ifft2 pair each: 78,400 pairs
//     at num = 2,000, ConOrder = 40) and multiplies the stack by mu_psi.  Everything in that loop is linear, so the
//     weighted sum over the snapshots is taken in Fourier space,
//         G[ii][e](k) = sum_i  F[ind3(ii, i)](k) CTF[ind3](k) / wiener_dom[i](k) * mu_psi[i][e],     ind3 = ConOrder - ii + i - 1,
//     on the Hermitian half plane (images are real, the CTF is even), and only ConOrder x psiTrunc inverse transforms
//     remain.  One forward transform per particle, shared by every psi of the PD (k_nlsa_weight_spectra).
//   * the snapshot order of a psi (posPath[PosPsi1]) is an index list `sel` applied inside the kernels; D, the spectra
//     and the CTF half planes are uploaded / computed once per PD and never permuted in memory.
//   * IMGT (NLSA.py:113-126) is rank 2 by construction (ConImgT = sum_{r<2} U_r s_r V_r^T):
//         IMGT[p][c] = sum_{i < ConOrder} sum_{r < 2} U[i Npix + p][r] * Q[r][i + c],   Q = diag(s) V^T psiC^T (2 x nI, host).
// Layouts: spectra H [n][N][Nh] complex128 (cuFFT D2Z), A / U [ConOrder N^2][E] row-major with the reference's
// transposed pixel order (row = ii N^2 + c N + r for picture pixel (r, c), NLSA.py:79), IMGT [nC][Npix] (frame-major).
#include "common.cuh"
#include <algorithm>

namespace mem {

#define NL_CUFFT(x) MEM_CUFFT(x)

// H[img][k] = rfft2(img)[k] * CTF[img][k_full];  Ch[img][k] = CTF half plane
__global__ void k_nlsa_weight_spectra(double2* __restrict__ H, const double* __restrict__ ctf, double* __restrict__ Ch,
                                      int N, int Nh, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t img = e / ((size_t)N * Nh);
    const int rem = (int)(e - img * (size_t)N * Nh);
    const int r = rem / Nh, c = rem - r * Nh;
    const double w = ctf[(img * N + r) * N + c];
    double2 h = H[e];
    h.x *= w;
    h.y *= w;
    H[e] = h;
    Ch[e] = w;
  }
}

// NLSA.py:30-33: ConD[r][c] = sum_{i < ConOrder} DD[r + i][c + i], DD = D[sel][:, sel]; same summation order
template <class T>
__global__ void __launch_bounds__(256) k_nlsa_cond(const T* __restrict__ D, int nAll, const int* __restrict__ sel, int nI,
                                                   int ConOrder, double* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y;
  if (c >= nI) return;
  double acc = 0.0;
  for (int i = 0; i < ConOrder; ++i) acc += (double)D[(size_t)sel[r + i] * nAll + sel[c + i]];
  out[(size_t)r * nI + c] = acc;
}

// get_wiener.py:16-20: wd[i] = sum_{ii < ConOrder} CTF1[ConOrder - ii + i]^2 + 1/5 (that order); stores 1 / wd
__global__ void __launch_bounds__(256) k_nlsa_wiener(const double* __restrict__ Ch, const int* __restrict__ sel, int nI,
                                                     int ConOrder, int Kh, double* __restrict__ iw) {
  const int k = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  if (k >= Kh) return;
  double acc = 0.0;
  for (int ii = 0; ii < ConOrder; ++ii) {
    const double c = Ch[(size_t)sel[ConOrder - ii + i] * Kh + k];
    acc = acc + c * c;
  }
  iw[(size_t)i * Kh + k] = 1.0 / (acc + 1.0 / 5);
}

// G[ii][e][k] = sum_i H[sel[i + ConOrder - ii - 1]][k] * (mu_psi[i][e] * iw[i][k]);  ST shifts x E columns per thread
constexpr int NL_ST = 4, NL_EMAX = 8;
__global__ void __launch_bounds__(128) k_nlsa_supervector_spectra(const double2* __restrict__ H, const int* __restrict__ sel,
                                                                  const double* __restrict__ iw,
                                                                  const double* __restrict__ mu_psi, int nI, int ConOrder,
                                                                  int E, int e0, int Kh, double2* __restrict__ G) {
  extern __shared__ double s_mu[];                       // [chunk][NL_EMAX]
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int ii0 = blockIdx.y * NL_ST;
  const int ne = min(NL_EMAX, E - e0);
  double2 acc[NL_ST][NL_EMAX];
#pragma unroll
  for (int s = 0; s < NL_ST; ++s)
#pragma unroll
    for (int e = 0; e < NL_EMAX; ++e) acc[s][e] = make_double2(0.0, 0.0);
  constexpr int CH = 128;
  for (int i0 = 0; i0 < nI; i0 += CH) {
    const int ni = min(CH, nI - i0);
    __syncthreads();
    for (int t = threadIdx.x; t < ni * NL_EMAX; t += 128) {
      const int i = t / NL_EMAX, e = t - i * NL_EMAX;
      s_mu[t] = e < ne ? mu_psi[(size_t)(i0 + i) * E + e0 + e] : 0.0;
    }
    __syncthreads();
    if (k < Kh) {
      // shift s of the tile reads snapshot row (i + ConOrder - ii0 - 1) - s: from one i to the next the window of NL_ST rows
      // slides by one, so only the leading row is loaded (hw[0]); the others move down a register
      const int lead = ConOrder - ii0 - 1;
      double2 hw[NL_ST];
#pragma unroll
      for (int s = 1; s < NL_ST; ++s) {
        const int row = i0 + lead - s;                    // rows of iteration i0 behind the leading one (row >= 0 iff ii0 + s < ConOrder)
        hw[s] = row >= 0 ? H[(size_t)sel[row] * Kh + k] : make_double2(0.0, 0.0);
      }
      for (int i = 0; i < ni; ++i) {
        hw[0] = H[(size_t)sel[i0 + i + lead] * Kh + k];
        const double w = iw[(size_t)(i0 + i) * Kh + k];
        double wm[NL_EMAX];
#pragma unroll
        for (int e = 0; e < NL_EMAX; ++e) wm[e] = s_mu[i * NL_EMAX + e] * w;
#pragma unroll
        for (int s = 0; s < NL_ST; ++s) {
          if (ii0 + s < ConOrder) {
#pragma unroll
            for (int e = 0; e < NL_EMAX; ++e) {
              acc[s][e].x = fma(hw[s].x, wm[e], acc[s][e].x);
              acc[s][e].y = fma(hw[s].y, wm[e], acc[s][e].y);
            }
          }
        }
#pragma unroll
        for (int s = NL_ST - 1; s > 0; --s) hw[s] = hw[s - 1];
      }
    }
  }
  if (k < Kh) {
#pragma unroll
    for (int s = 0; s < NL_ST; ++s) {
      const int ii = ii0 + s;
      if (ii < ConOrder) {
#pragma unroll
        for (int e = 0; e < NL_EMAX; ++e)
          if (e < ne) G[((size_t)ii * E + e0 + e) * Kh + k] = acc[s][e];
      }
    }
  }
}

// A[ii N^2 + c N + r][e] = g[ii][e][r][c] / N^2 * msk2[r][c]   (NLSA.py:75-79: ifft2(...).real * msk2, img.T.reshape(-1))
__global__ void __launch_bounds__(256) k_nlsa_pack(const double* __restrict__ g, const double* __restrict__ msk2, int N, int E,
                                                   double* __restrict__ A) {
  __shared__ double tile[32][33];
  const int ii = blockIdx.z / E, e = blockIdx.z - ii * E;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const double* src = g + ((size_t)ii * E + e) * N * N;
  const double inv = 1.0 / ((double)N * N);
  for (int y = ty; y < 32; y += 8) {
    const int r = r0 + y, c = c0 + tx;
    if (r < N && c < N) tile[y][tx] = src[(size_t)r * N + c] * inv * (msk2 ? msk2[(size_t)r * N + c] : 1.0);
  }
  __syncthreads();
  for (int y = ty; y < 32; y += 8) {
    const int c = c0 + y, r = r0 + tx;                   // consecutive threads -> consecutive r (rows of A)
    if (r < N && c < N) A[((size_t)ii * N * N + (size_t)c * N + r) * E + e] = tile[tx][y];
  }
}

// partial A^T A: CTA b sums rows [b * rows_per, ...) into part[b][E][E] (fixed order), reduced on the host side of the call
__global__ void __launch_bounds__(256) k_nlsa_gram_small(const double* __restrict__ A, size_t rows, int E, size_t rows_per,
                                                         double* __restrict__ part) {
  __shared__ double red[8];
  const size_t a = (size_t)blockIdx.x * rows_per, b = min(rows, a + rows_per);
  for (int pq = 0; pq < E * E; ++pq) {
    const int p = pq / E, q = pq - p * E;
    if (q < p) continue;
    double acc = 0.0;
    for (size_t r = a + threadIdx.x; r < b; r += 256) acc = fma(A[r * E + p], A[r * E + q], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      part[(size_t)blockIdx.x * E * E + pq] = s;
    }
  }
}

// the same for E <= 8 (psiTrunc's default): every thread walks its rows once, the E (E + 1) / 2 products of a row accumulate in
// registers (A is read ONCE, 64 contiguous bytes per row, instead of once per (p, q) pair with a stride of E doubles)
__global__ void __launch_bounds__(256) k_nlsa_gram_small8(const double* __restrict__ A, size_t rows, int E, size_t rows_per,
                                                          double* __restrict__ part) {
  __shared__ double red[8];
  const size_t a = (size_t)blockIdx.x * rows_per, b = min(rows, a + rows_per);
  double acc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
  for (size_t r = a + threadIdx.x; r < b; r += 256) {
    double x[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) x[p] = p < E ? A[r * E + p] : 0.0;
    int k = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int q = p; q < 8; ++q) {
        acc[k] = fma(x[p], x[q], acc[k]);
        ++k;
      }
  }
  int k = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int q = p; q < 8; ++q, ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0 && p < E && q < E) {
        double s2 = 0.0;
        for (int w = 0; w < 8; ++w) s2 += red[w];
        part[(size_t)blockIdx.x * E * E + p * E + q] = s2;
      }
    }
}

// U = A M (svdRF.py:25: U = A (V S^-1)), M [E][E] row-major in constant-sized shared memory
__global__ void __launch_bounds__(256) k_nlsa_project(const double* __restrict__ A, const double* __restrict__ M, size_t rows,
                                                      int E, double* __restrict__ U) {
  __shared__ double sM[32 * 32];
  for (int t = threadIdx.x; t < E * E; t += 256) sM[t] = M[t];
  __syncthreads();
  const size_t r = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  double a[32];
  for (int p = 0; p < E; ++p) a[p] = A[r * E + p];
  for (int q = 0; q < E; ++q) {
    double acc = 0.0;
    for (int p = 0; p < E; ++p) acc = fma(a[p], sM[p * E + q], acc);
    U[r * E + q] = acc;
  }
}

// NLSA.py:95-103: Topo_mean[p][e] = mean_k U[k Npix + p][e]
__global__ void __launch_bounds__(256) k_nlsa_topo(const double* __restrict__ U, int Npix, int ConOrder, int E,
                                                   double* __restrict__ topo) {
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (size_t)Npix * E) return;
  const size_t p = t / E;
  const int e = (int)(t - p * E);
  double acc = 0.0;
  for (int k = 0; k < ConOrder; ++k) acc += U[((size_t)k * Npix + p) * E + e];
  topo[t] = acc / ConOrder;
}

// IMGT[c][p] = sum_i sum_{r<2} U[i Npix + p][r] Q[r][i + c]   (NLSA.py:106-126); CT frames per thread
constexpr int NL_CT = 8;
__global__ void __launch_bounds__(128) k_nlsa_reconstruct(const double* __restrict__ U, const double* __restrict__ Q, int Npix,
                                                          int ConOrder, int E, int nI, int nC, double* __restrict__ IMGT) {
  extern __shared__ double sQ[];                         // [2][ConOrder + NL_CT]
  const int c0 = blockIdx.y * NL_CT;
  const int span = ConOrder + NL_CT;
  for (int t = threadIdx.x; t < 2 * span; t += 128) {
    const int r = t / span, m = t - r * span;
    sQ[t] = (c0 + m < nI) ? Q[(size_t)r * nI + c0 + m] : 0.0;
  }
  __syncthreads();
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= Npix) return;
  double acc[NL_CT];
#pragma unroll
  for (int c = 0; c < NL_CT; ++c) acc[c] = 0.0;
  for (int i = 0; i < ConOrder; ++i) {
    const double u0 = U[((size_t)i * Npix + p) * E], u1 = U[((size_t)i * Npix + p) * E + 1];
#pragma unroll
    for (int c = 0; c < NL_CT; ++c) acc[c] = fma(u1, sQ[span + i + c], fma(u0, sQ[i + c], acc[c]));
  }
#pragma unroll
  for (int c = 0; c < NL_CT; ++c)
    if (c0 + c < nC) IMGT[(size_t)(c0 + c) * Npix + p] = acc[c];
}

// NLSA.py:129-136: every frame to mean 0, std 1 (population std, two passes); one CTA per frame.  Also aa = sum x^2.
__global__ void __launch_bounds__(256) k_nlsa_normalize(double* __restrict__ IMGT, int Npix, double* __restrict__ aa) {
  __shared__ double red[8];
  __shared__ double bc;
  double* x = IMGT + (size_t)blockIdx.x * Npix;
  auto bsum = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  double s = 0.0;
  for (int p = threadIdx.x; p < Npix; p += 256) s += x[p];
  const double mean = bsum(s) / Npix;
  double v = 0.0;
  for (int p = threadIdx.x; p < Npix; p += 256) {
    const double d = x[p] - mean;
    v = fma(d, d, v);
  }
  const double sd = sqrt(bsum(v) / Npix);
  double q = 0.0;
  for (int p = threadIdx.x; p < Npix; p += 256) {
    const double y = (x[p] - mean) / sd;
    x[p] = y;
    q = fma(y, y, q);
  }
  q = bsum(q);
  if (threadIdx.x == 0) aa[blockIdx.x] = q;
}

// L2_distance.py:36-41 then **2 (NLSA.py:144): D2[a][b] = sqrt(t)^2, t = aa[a] + aa[b] - 2 <x_a, x_b>, t < 1e-8 -> 0.
// 128 x 128 tile per CTA (256 threads, 8 x 8 per thread with rows ty + 16 i / columns tx + 16 j: the operand reads from shared
// memory are broadcasts or unit-stride), K chunks of 8; upper-triangle tiles only, mirrored on the way out.  64 FMA per 16
// shared-memory loads: the first version (4 x 4 per thread) ran at 10 % of the float64 pipe.
__global__ void __launch_bounds__(256) k_nlsa_l2(const double* __restrict__ X, const double* __restrict__ aa, int nC, int Npix,
                                                 double* __restrict__ D2) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ double sa[8][128], sb[8][128];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  const int lr = threadIdx.x >> 1, lk = (threadIdx.x & 1) * 4;     // loader: row lr of the tile, 4 consecutive k
  const int ra = bi * 128 + lr, rb = bj * 128 + lr;
  const double* pa = X + (size_t)min(ra, nC - 1) * Npix + lk;
  const double* pb = X + (size_t)min(rb, nC - 1) * Npix + lk;
  for (int k0 = 0; k0 < Npix; k0 += 8) {
    double va[4], vb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool ok = k0 + lk + q < Npix;
      va[q] = (ok && ra < nC) ? pa[k0 + q] : 0.0;
      vb[q] = (ok && rb < nC) ? pb[k0 + q] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[lk + q][lr] = va[q];
      sb[lk + q][lr] = vb[q];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sa[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = sb[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = bi * 128 + ty + 16 * i, c = bj * 128 + tx + 16 * j;
      if (r < nC && c < nC) {
        double t = aa[r] + aa[c] - 2 * acc[i][j];
        if (t < 1e-8) t = 0.0;
        const double sq = sqrt(t);
        t = sq * sq;
        D2[(size_t)r * nC + c] = t;
        if (bj > bi) D2[(size_t)c * nC + r] = t;
      }
    }
}

// ------------------------------------------------------------------------------------------------
// fit_1D_open_manifold_3D.op (:60-146) on the device: x_ij = a_j cos(j pi tau_i) + b_j, j = 1..3.  The reference solves
// one quintic per point and iteration with np.roots (nS x <= 101 companion-matrix eigenproblems in a Python loop); here a thread
// owns a point: the real roots of
// d R_p / d beta (beta = cos(pi tau)) inside [-1, 1] are isolated through the derivative chain (between two consecutive
// critical points a polynomial is monotone: one sign test + safeguarded Newton per interval, no root can be missed), the
// candidate with the smallest residual wins (tau = 0 and 1 always compete, solve_d_R_d_tau_p_3D.py:48-50), and the
// 2 x 2 normal equations for (a_j, b_j) are block reductions in a fixed order (k_manifold_tau / k_manifold_update below).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double horner(const double* c, int n, double x) {     // c[0] x^n + ... + c[n]
  double v = c[0];
  for (int k = 1; k <= n; ++k) v = fma(v, x, c[k]);
  return v;
}

// root of the degree-n polynomial c in [lo, hi] given f(lo) f(hi) <= 0 and monotone f: Newton with bisection safeguard
__device__ double bracketed_root(const double* c, const double* dc, int n, double lo, double hi, double flo, double fhi) {
  if (flo == 0.0) return lo;
  if (fhi == 0.0) return hi;
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    const double f = horner(c, n, x);
    if (f == 0.0) return x;
    if ((f < 0.0) == (flo < 0.0)) { lo = x; flo = f; } else { hi = x; fhi = f; }
    if (hi - lo <= 4.5e-16 * fmax(1.0, fabs(x))) break;
    const double d = horner(dc, n - 1, x);
    double xn = (d != 0.0) ? x - f / d : lo - 1.0;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (xn == x) break;
    x = xn;
  }
  return x;
}

// real roots of c (degree n <= 5, leading coefficients may vanish) inside [-1, 1], ascending; crit = the roots of its
// derivative inside (-1, 1) (ascending, ncrit of them)
__device__ int roots_between(const double* c, const double* dc, int n, const double* crit, int ncrit, double* out) {
  int cnt = 0;
  double lo = -1.0, flo = horner(c, n, lo);
  for (int k = 0; k <= ncrit; ++k) {
    const double hi = (k < ncrit) ? crit[k] : 1.0;
    const double fhi = horner(c, n, hi);
    if ((flo <= 0.0 && fhi >= 0.0) || (flo >= 0.0 && fhi <= 0.0)) {
      if (!(flo == 0.0 && fhi == 0.0 && lo != hi && horner(c, n, 0.5 * (lo + hi)) == 0.0)) {   // not identically zero
        const double r = bracketed_root(c, dc, n, lo, hi, flo, fhi);
        if (cnt == 0 || r > out[cnt - 1] + 1e-15) out[cnt++] = r;
      }
    }
    lo = hi;
    flo = fhi;
  }
  return cnt;
}

__device__ double manifold_tau(const double x0, const double x1, const double x2, const double* a, const double* b) {
  // d R / d beta, solve_d_R_d_tau_p_3D.py:38-43
  double c5[6] = {48 * a[2] * a[2], 0.0, 8 * a[1] * a[1] - 48 * a[2] * a[2], -12 * a[2] * (x2 - b[2]),
                  a[0] * a[0] - 4 * a[1] * a[1] + 9 * a[2] * a[2] - 4 * a[1] * (x1 - b[1]),
                  -a[0] * (x0 - b[0]) + 3 * a[2] * (x2 - b[2])};
  double c4[5], c3[4], c2[3], c1[2];
  for (int k = 0; k < 5; ++k) c4[k] = (5 - k) * c5[k];
  for (int k = 0; k < 4; ++k) c3[k] = (4 - k) * c4[k];
  for (int k = 0; k < 3; ++k) c2[k] = (3 - k) * c3[k];
  for (int k = 0; k < 2; ++k) c1[k] = (2 - k) * c2[k];
  double r1[1], r2[2], r3[3], r4[4], r5[5];
  int n1 = 0;
  if (c1[0] != 0.0) {                                     // root of the linear c1 = derivative of the quadratic c2
    const double r = -c1[1] / c1[0];
    if (r > -1.0 && r < 1.0) r1[n1++] = r;
  }
  const int n2 = roots_between(c2, c1, 2, r1, n1, r2);
  int m2 = 0;
  double q2[2];
  for (int k = 0; k < n2; ++k) if (r2[k] > -1.0 && r2[k] < 1.0) q2[m2++] = r2[k];
  const int n3 = roots_between(c3, c2, 3, q2, m2, r3);
  int m3 = 0;
  double q3[3];
  for (int k = 0; k < n3; ++k) if (r3[k] > -1.0 && r3[k] < 1.0) q3[m3++] = r3[k];
  const int n4 = roots_between(c4, c3, 4, q3, m3, r4);
  int m4 = 0;
  double q4[4];
  for (int k = 0; k < n4; ++k) if (r4[k] > -1.0 && r4[k] < 1.0) q4[m4++] = r4[k];
  const int n5 = roots_between(c5, c4, 5, q4, m4, r5);
  // candidates: arccos(beta) / pi for every root, then 0 and 1; the smallest R_p wins (first one on a tie)
  double best_tau = 0.0, best = INFINITY;
  for (int k = 0; k < n5 + 2; ++k) {
    const double tau = (k < n5) ? acos(r5[k]) / M_PI : (k == n5 ? 0.0 : 1.0);
    const double e0 = x0 - b[0] - a[0] * cos(tau * 1 * M_PI);
    const double e1 = x1 - b[1] - a[1] * cos(tau * 2 * M_PI);
    const double e2 = x2 - b[2] - a[2] * cos(tau * 3 * M_PI);
    const double R = e0 * e0 + e1 * e1 + e2 * e2;
    if (R < best) { best = R; best_tau = tau; }
  }
  return best_tau;
}

// The alternating iteration as two kernels per step — every point's tau over the whole grid, then one CTA for the 2 x 2 normal
// equations — chained through a small device state {stop, converged, iterations}: the host enqueues ten steps at a time and
// reads `stop` in between (a single CTA doing everything left 147 SMs idle: 4.4 ms for 960 points, now ~1 ms).
//   k_manifold_tau     skipped once stop is set; otherwise tau_p for every point from the current (a, b)
//   k_manifold_update  stop set -> return; converged set -> the taus just computed belong to the final (a, b): set stop;
//                      else new (a, b) from the taus, iteration count + 1, converged if both relative changes are small
__global__ void __launch_bounds__(128) k_manifold_tau(const double* __restrict__ x, int nS, const double* __restrict__ ab,
                                                      double* __restrict__ tau, const int* __restrict__ state) {
  if (state[0]) return;
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= nS) return;
  double a[3] = {ab[0], ab[1], ab[2]}, b[3] = {ab[3], ab[4], ab[5]};
  tau[p] = manifold_tau(x[3 * p], x[3 * p + 1], x[3 * p + 2], a, b);
}

__global__ void __launch_bounds__(1024) k_manifold_update(const double* __restrict__ x, int nS, double* __restrict__ ab,
                                                          const double* __restrict__ tau, double da_max, double db_max,
                                                          int* __restrict__ state) {
  __shared__ double red[32][12];
  __shared__ double tot[12];
  if (state[0]) return;
  if (state[1]) {
    if (threadIdx.x == 0) state[0] = 1;
    return;
  }
  double acc[12];
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
  for (int p = threadIdx.x; p < nS; p += 1024) {
    const double t = tau[p];
    for (int j = 0; j < 3; ++j) {
      const double cj = cos(t * (M_PI * (j + 1)));
      const double xv = x[3 * p + j];
      acc[j] = fma(cj, cj, acc[j]);
      acc[3 + j] += cj;
      acc[6 + j] = fma(xv, cj, acc[6 + j]);
      acc[9 + j] += xv;
    }
  }
  for (int k = 0; k < 12; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double s2 = 0.0;
    for (int w = 0; w < 32; ++w) s2 += red[w][threadIdx.x];
    tot[threadIdx.x] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double da = 0.0, db = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double A11 = tot[j], A12 = tot[3 + j], b1 = tot[6 + j], b2 = tot[9 + j], A22 = (double)nS;
      const double det = A11 * A22 - A12 * A12;
      const double an = (b1 * A22 - A12 * b2) / det, bn = (A11 * b2 - A12 * b1) / det;
      da = fmax(da, fabs(an - ab[j]) / (fabs(an) + 1e-4));
      db = fmax(db, fabs(bn - ab[3 + j]) / (fabs(bn) + 1e-4));
      ab[j] = an;
      ab[3 + j] = bn;
    }
    state[2] += 1;
    if (da * 100 < da_max && db * 100 < db_max) state[1] = 1;
  }
}

// x [nS][3] HOST, ab [6] HOST in / out, tau [nS] HOST out.  Synchronises.
int manifold_fit_host(mem_ctx* ctx, const double* x_host, int nS, double* ab_host, double* tau_host, int max_iter,
                      double da_max, double db_max, int* iters_host, cudaStream_t st) {
  if (nS < 1) {
    set_error("manifold_fit: nS >= 1");
    return 1;
  }
  MEM_CHECK(ctx->small_out.ensure((size_t)(4 * nS + 8) * sizeof(double)));
  double* dx = ctx->small_out.as<double>();
  double* dab = dx + 3 * nS;
  double* dtau = dab + 6;
  int* dstate = reinterpret_cast<int*>(dtau + nS);          // {stop, converged, iterations}
  MEM_CUDA(cudaMemcpyAsync(dx, x_host, (size_t)3 * nS * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(dab, ab_host, 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemsetAsync(dstate, 0, 4 * sizeof(int), st));
  const int grid = (nS + 127) / 128;
  MEM_LAUNCH(ctx, k_manifold_tau, grid, 128, 0, st, dx, nS, dab, dtau, dstate);
  int state[4] = {0, 0, 0, 0};
  for (int done = 0; done < max_iter && !state[0];) {
    const int n = std::min(10, max_iter - done);
    for (int k = 0; k < n; ++k) {
      MEM_LAUNCH(ctx, k_manifold_update, 1, 1024, 0, st, dx, nS, dab, dtau, da_max, db_max, dstate);
      MEM_LAUNCH(ctx, k_manifold_tau, grid, 128, 0, st, dx, nS, dab, dtau, dstate);
    }
    done += n;
    MEM_CUDA(cudaMemcpyAsync(state, dstate, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    MEM_CUDA(cudaStreamSynchronize(st));
    if (state[1]) break;                                   // converged: the tau kernel behind that update has run
  }
  MEM_CUDA(cudaMemcpyAsync(ab_host, dab, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaMemcpyAsync(tau_host, dtau, (size_t)nS * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  iters_host[0] = state[2];
  return 0;
}

// ------------------------------------------------------------------------------------------------ host side
// float64 2-D plans, cached per context (creating one costs 1 - 15 ms; a PD asks for the same two shapes once per psi)
static int plan_many_d(mem_ctx* ctx, cufftHandle* h, int N, int batch, cufftType type, cudaStream_t st) {
  const long long key = ((long long)type << 48) + ((long long)N << 28) + batch;
  auto it = ctx->plans_d.find(key);
  if (it == ctx->plans_d.end()) {
    if (ctx->plans_d.size() >= 16) {                      // a few shapes per PD size: keep the cache small
      for (auto& kv : ctx->plans_d) cufftDestroy(kv.second);
      ctx->plans_d.clear();
    }
    int n[2] = {N, N};
    cufftHandle pl;
    NL_CUFFT(cufftPlanMany(&pl, 2, n, nullptr, 1, 0, nullptr, 1, 0, type, batch));
    it = ctx->plans_d.emplace(key, pl).first;
  }
  *h = it->second;
  NL_CUFFT(cufftSetStream(*h, st));
  return 0;
}

int nlsa_spectra_device(mem_ctx* ctx, const double* img, const double* ctf, int n, int N, double2* H, double* Ch,
                        cudaStream_t st) {
  if (n < 1 || N < 2) {
    set_error("nlsa_spectra: need n >= 1 and N >= 2 (n=%d N=%d)", n, N);
    return 1;
  }
  const int Nh = N / 2 + 1;
  // batches of at most 512 images bound the cuFFT work area
  for (int a = 0; a < n; a += 512) {
    const int nb = std::min(512, n - a);
    cufftHandle pl;
    MEM_CHECK(plan_many_d(ctx, &pl, N, nb, CUFFT_D2Z, st));
    cufftResult r = cufftExecD2Z(pl, const_cast<double*>(img) + (size_t)a * N * N,
                                 reinterpret_cast<cufftDoubleComplex*>(H + (size_t)a * N * Nh));
    if (r != CUFFT_SUCCESS) {
      set_error("cufftExecD2Z failed (%d)", (int)r);
      return 1;
    }
  }
  const size_t total = (size_t)n * N * Nh;
  MEM_LAUNCH(ctx, k_nlsa_weight_spectra, (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, st, H, ctf, Ch, N, Nh,
             total);
  return 0;
}

int nlsa_cond_device(mem_ctx* ctx, const void* D, int elem_bytes, int nAll, const int* sel, int num, int ConOrder,
                     double* out, cudaStream_t st) {
  const int nI = num - ConOrder;
  if (nI < 1 || ConOrder < 1 || (elem_bytes != 4 && elem_bytes != 8)) {
    set_error("nlsa_cond: need 1 <= ConOrder < num and float32 / float64 D (num=%d ConOrder=%d)", num, ConOrder);
    return 1;
  }
  const dim3 grid((nI + 255) / 256, nI);
  if (elem_bytes == 8) MEM_LAUNCH(ctx, k_nlsa_cond<double>, grid, 256, 0, st, (const double*)D, nAll, sel, nI, ConOrder, out);
  else MEM_LAUNCH(ctx, k_nlsa_cond<float>, grid, 256, 0, st, (const float*)D, nAll, sel, nI, ConOrder, out);
  return 0;
}

// A [ConOrder N^2][E] from the weighted spectra; mu_psi [nI][E] float64 on the HOST; msk2 [N][N] float64 device or null
int nlsa_supervectors_device(mem_ctx* ctx, const double2* H, const double* Ch, const int* sel, const double* mu_psi_host,
                             int num, int ConOrder, int E, int N, const double* msk2, double* A, cudaStream_t st) {
  const int nI = num - ConOrder, Nh = N / 2 + 1, Kh = N * Nh;
  if (nI < 1 || ConOrder < 1 || E < 1 || E > 32) {
    set_error("nlsa_supervectors: need 1 <= ConOrder < num and 1 <= E <= 32 (num=%d ConOrder=%d E=%d)", num, ConOrder, E);
    return 1;
  }
  const size_t b_iw = ((size_t)nI * Kh * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_mu = ((size_t)nI * E * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_G = ((size_t)ConOrder * E * Kh * sizeof(double2) + 255) & ~(size_t)255;
  const size_t b_g = (size_t)ConOrder * E * N * N * sizeof(double);
  MEM_CHECK(ctx->scratch.ensure(b_iw + b_mu + b_G + b_g));
  uint8_t* base = ctx->scratch.as<uint8_t>();
  double* iw = reinterpret_cast<double*>(base);
  double* mu = reinterpret_cast<double*>(base + b_iw);
  double2* G = reinterpret_cast<double2*>(base + b_iw + b_mu);
  double* g = reinterpret_cast<double*>(base + b_iw + b_mu + b_G);
  MEM_CUDA(cudaMemcpyAsync(mu, mu_psi_host, (size_t)nI * E * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));                      // mu_psi_host may be a temporary of the caller
  MEM_LAUNCH(ctx, k_nlsa_wiener, dim3((Kh + 255) / 256, nI), 256, 0, st, Ch, sel, nI, ConOrder, Kh, iw);
  for (int e0 = 0; e0 < E; e0 += NL_EMAX)
    MEM_LAUNCH(ctx, k_nlsa_supervector_spectra, dim3((Kh + 127) / 128, (ConOrder + NL_ST - 1) / NL_ST), 128,
               128 * NL_EMAX * sizeof(double), st, H, sel, iw, mu, nI, ConOrder, E, e0, Kh, G);
  {
    cufftHandle pl;
    MEM_CHECK(plan_many_d(ctx, &pl, N, ConOrder * E, CUFFT_Z2D, st));
    cufftResult r = cufftExecZ2D(pl, reinterpret_cast<cufftDoubleComplex*>(G), g);
    if (r != CUFFT_SUCCESS) {
      set_error("cufftExecZ2D failed (%d)", (int)r);
      return 1;
    }
  }
  MEM_LAUNCH(ctx, k_nlsa_pack, dim3((N + 31) / 32, (N + 31) / 32, ConOrder * E), 256, 0, st, g, msk2, N, E, A);
  return 0;
}

// AtA [E][E] (HOST) = A^T A
int nlsa_gram_small_device(mem_ctx* ctx, const double* A, long long rows, int E, double* AtA_host, cudaStream_t st) {
  if (E < 1 || E > 32 || rows < 1) {
    set_error("nlsa_gram_small: 1 <= E <= 32, rows >= 1");
    return 1;
  }
  const int nb = (int)std::min<long long>(296, (rows + 1023) / 1024);
  const size_t rows_per = ((size_t)rows + nb - 1) / nb;
  MEM_CHECK(ctx->small_out.ensure((size_t)nb * E * E * sizeof(double)));
  double* part = ctx->small_out.as<double>();
  MEM_CUDA(cudaMemsetAsync(part, 0, (size_t)nb * E * E * sizeof(double), st));
  if (E <= 8) MEM_LAUNCH(ctx, k_nlsa_gram_small8, nb, 256, 0, st, A, (size_t)rows, E, rows_per, part);
  else MEM_LAUNCH(ctx, k_nlsa_gram_small, nb, 256, 0, st, A, (size_t)rows, E, rows_per, part);
  std::vector<double> h((size_t)nb * E * E);
  MEM_CUDA(cudaMemcpyAsync(h.data(), part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  for (int p = 0; p < E; ++p)
    for (int q = p; q < E; ++q) {
      double s = 0.0;
      for (int b = 0; b < nb; ++b) s += h[(size_t)b * E * E + p * E + q];
      AtA_host[p * E + q] = AtA_host[q * E + p] = s;
    }
  return 0;
}

// U = A M (M [E][E] HOST), Topo_mean [Npix][E] -> HOST
int nlsa_project_device(mem_ctx* ctx, const double* A, long long rows, int E, const double* M_host, double* U, int Npix,
                        int ConOrder, double* topo_host, cudaStream_t st) {
  if (E < 2 || E > 32 || rows != (long long)Npix * ConOrder) {
    set_error("nlsa_project: 2 <= E <= 32 and rows == Npix * ConOrder");
    return 1;
  }
  MEM_CHECK(ctx->small_out.ensure((size_t)E * E * sizeof(double) + (size_t)Npix * E * sizeof(double)));
  double* dM = ctx->small_out.as<double>();
  double* dT = dM + E * E;
  MEM_CUDA(cudaMemcpyAsync(dM, M_host, (size_t)E * E * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  MEM_LAUNCH(ctx, k_nlsa_project, (unsigned)((rows + 255) / 256), 256, 0, st, A, dM, (size_t)rows, E, U);
  MEM_LAUNCH(ctx, k_nlsa_topo, (unsigned)(((size_t)Npix * E + 255) / 256), 256, 0, st, U, Npix, ConOrder, E, dT);
  MEM_CUDA(cudaMemcpyAsync(topo_host, dT, (size_t)Npix * E * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// IMGT [nC][Npix] (device, normalised frames) and D2 [nC][nC] (device) = L2_distance(IMGT, IMGT)^2; Q [2][nI] HOST
int nlsa_reconstruct_device(mem_ctx* ctx, const double* U, int Npix, int ConOrder, int E, const double* Q_host, int nI, int nC,
                            double* IMGT, double* D2, cudaStream_t st) {
  if (nC < 1 || nC + ConOrder > nI + 0 || E < 2) {
    set_error("nlsa_reconstruct: need 1 <= nC <= nI - ConOrder and E >= 2 (nC=%d nI=%d ConOrder=%d)", nC, nI, ConOrder);
    return 1;
  }
  MEM_CHECK(ctx->small_out.ensure((size_t)(2 * nI + nC) * sizeof(double)));
  double* dQ = ctx->small_out.as<double>();
  double* aa = dQ + 2 * nI;
  MEM_CUDA(cudaMemcpyAsync(dQ, Q_host, (size_t)2 * nI * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  MEM_LAUNCH(ctx, k_nlsa_reconstruct, dim3((Npix + 127) / 128, (nC + NL_CT - 1) / NL_CT), 128,
             2 * (ConOrder + NL_CT) * sizeof(double), st, U, dQ, Npix, ConOrder, E, nI, nC, IMGT);
  MEM_LAUNCH(ctx, k_nlsa_normalize, nC, 256, 0, st, IMGT, Npix, aa);
  if (D2) MEM_LAUNCH(ctx, k_nlsa_l2, dim3((nC + 127) / 128, (nC + 127) / 128), 256, 0, st, IMGT, aa, nC, Npix, D2);
  return 0;
}

}  // namespace mem
