// preprocess.cu — everything of getDistanceCTF_local_Conj9combinedS2.op except the contraction:
// ingest/normalise (a2,a3), low-pass (a5), in-plane alignment (a7), CTF (a8), FFT + phase flip (a10),
// Wiener/flip averages (a11), intensity (a13) and the operand writer that feeds the contraction (a12).
// All kernels are HBM-bound; reference lines are cited per kernel.
#include "common.cuh"

#include <math.h>
#include <algorithm>
#include <stdarg.h>

namespace mem {

// own FFT kernels (lowpass.cu) exist for this box and are not switched off
static inline bool own_fft(const mem_ctx* ctx, int N) { return colfilter_supported(N) && !ctx->cufft_lowpass; }


int ingest_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float* out, int nS, int N, int transposed,
               cudaStream_t st);
int shift_run(mem_ctx* ctx, const float* raw, const double* shift, float* tmp, float* out, int nS, int N,
              cudaStream_t st);
int align_run(mem_ctx* ctx, float* A, float* B, float* imgAll, const double* psi_deg, double psi_p_deg, double2* cs,
              const uint8_t* msk2, int nS, int N, cudaStream_t st, int rows_done);
int align_batch_run(mem_ctx* ctx, float* A, float* B, float* imgAll, const double* psi_deg, double2* cs, const double2* cs2,
                    const uint8_t* pid2, int nS, int N, cudaStream_t st, int rows_done);

// ------------------------------------------------------------------------------------------------
// error string + arena
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return 0;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + (bytes >> 3) + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    e = cudaMalloc(&p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    p = nullptr;
    return 1;
  }
  cap = want;
  return 0;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

// ------------------------------------------------------------------------------------------------
// geometry tables (host build, once per N / filter)
// ------------------------------------------------------------------------------------------------
static inline int freq_of(int k, int N) { return k < (N + 1) / 2 ? k : k - N; }

int geometry_prepare(mem_ctx* ctx, int N, int filter_type, int filter_order, double Qc) {
  Geometry& g = ctx->geom;
  if (g.N == N && g.filter_type == filter_type && g.filter_order == filter_order && g.filter_Qc == Qc) return 0;
  const int Nh = N / 2 + 1, Kh = N * Nh;
  std::vector<int> r2(Kh);
  int r2max = 0;
  for (int ky = 0; ky < N; ++ky)
    for (int kx = 0; kx < Nh; ++kx) {
      int fy = freq_of(ky, N), fx = kx;
      int v = fy * fy + fx * fx;
      r2[ky * Nh + kx] = v;
      r2max = std::max(r2max, v);
    }
  // bins numbered by the address of their first pixel (not by radius): consecutive threads of the radial
  // operand kernel then start their gathers at neighbouring addresses; the column order of S1/S2 is arbitrary
  // as long as both use the same one
  std::vector<int> bin_id(r2max + 1, -1);
  std::vector<int> r2_of_bin;
  for (int p = 0; p < Kh; ++p)
    if (bin_id[r2[p]] < 0) {
      bin_id[r2[p]] = (int)r2_of_bin.size();
      r2_of_bin.push_back(r2[p]);
    }
  const int Kr = (int)r2_of_bin.size();
  std::vector<int> bin_of_pix(Kh), bin_start(Kr + 1, 0), bin_pix(Kh);
  for (int p = 0; p < Kh; ++p) {
    bin_of_pix[p] = bin_id[r2[p]];
    bin_start[bin_of_pix[p] + 1]++;
  }
  for (int b = 0; b < Kr; ++b) bin_start[b + 1] += bin_start[b];
  {
    std::vector<int> fill(bin_start.begin(), bin_start.end() - 1);
    for (int p = 0; p < Kh; ++p) bin_pix[fill[bin_of_pix[p]]++] = p;
  }
  // low-pass table, getDistanceCTF...py:139-167 + :288 (ifftshift) with the 1/N^2 of the unnormalised inverse FFT
  std::vector<float> G(Kh);
  for (int p = 0; p < Kh; ++p) {
    double Q = sqrt((double)r2[p]) / (N / 2.0);
    double gv;
    if (filter_type == 1)
      gv = exp(-(log(2.0) / 2.0) * (Q / Qc) * (Q / Qc));
    else
      gv = sqrt(1.0 / (1.0 + pow(Q / Qc, 2.0 * filter_order)));
    G[p] = (float)(gv / ((double)N * N));
  }
  // S3 membership: one representative of every conjugate pair, weight-2 entries only
  const bool even = (N % 2 == 0);
  std::vector<int> special;
  std::vector<int> s3_col(Kh, -1);
  int j = 0;
  for (int ky = 0; ky < N; ++ky)
    for (int kx = 0; kx < Nh; ++kx) {
      const int p = ky * Nh + kx;
      const bool selfcol = (kx == 0) || (even && kx == N / 2);
      if (!selfcol) {
        s3_col[p] = j++;
      } else {
        const bool selfrow = (ky == 0) || (even && ky == N / 2);
        if (selfrow)
          special.push_back(p);
        else if (ky < (N + 1) / 2)
          s3_col[p] = j++;
      }
    }
  const int K3 = j;
  const int n_special = (int)special.size();
  const int n1_blocks = (Kr + n_special + 31) / 32;
  const int n3_blocks = (2 * K3 + 31) / 32;
  for (int p = 0; p < Kh; ++p)
    if (s3_col[p] >= 0) s3_col[p] = 64 * n1_blocks + 2 * s3_col[p];
  while (special.size() < 4) special.push_back(-1);

  // folded quadrant (rows a and N-a share |k|^2): CSR of its entries by bin
  const int Na = N / 2 + 1, Kq = Na * Nh;
  std::vector<int> fold_bin(Kq), fold_start(Kr + 1, 0), fold_ent(Kq);
  for (int e = 0; e < Kq; ++e) {
    fold_bin[e] = bin_of_pix[e];          // row a <= N/2 of the half spectrum has |fy| = a
    fold_start[fold_bin[e] + 1]++;
  }
  for (int b = 0; b < Kr; ++b) fold_start[b + 1] += fold_start[b];
  {
    std::vector<int> fill(fold_start.begin(), fold_start.end() - 1);
    for (int e = 0; e < Kq; ++e) fold_ent[fill[fold_bin[e]]++] = e;
  }
  MEM_CHECK(g.fold_bin.ensure(Kq * sizeof(int)));
  MEM_CHECK(g.fold_start.ensure((Kr + 1) * sizeof(int)));
  MEM_CHECK(g.fold_ent.ensure(Kq * sizeof(int)));
  MEM_CUDA(cudaMemcpy(g.fold_bin.p, fold_bin.data(), Kq * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.fold_start.p, fold_start.data(), (Kr + 1) * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.fold_ent.p, fold_ent.data(), Kq * sizeof(int), cudaMemcpyHostToDevice));
  g.Na = Na;
  MEM_CHECK(g.Gtab.ensure(Kh * sizeof(float)));
  MEM_CHECK(g.bin_of_pix.ensure(Kh * sizeof(int)));
  MEM_CHECK(g.r2_of_bin.ensure(Kr * sizeof(int)));
  MEM_CHECK(g.bin_start.ensure((Kr + 1) * sizeof(int)));
  MEM_CHECK(g.bin_pix.ensure(Kh * sizeof(int)));
  MEM_CHECK(g.s3_col.ensure(Kh * sizeof(int)));
  MEM_CHECK(g.special_pix.ensure(4 * sizeof(int)));
  MEM_CUDA(cudaMemcpy(g.Gtab.p, G.data(), Kh * sizeof(float), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.bin_of_pix.p, bin_of_pix.data(), Kh * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.r2_of_bin.p, r2_of_bin.data(), Kr * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.bin_start.p, bin_start.data(), (Kr + 1) * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.bin_pix.p, bin_pix.data(), Kh * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.s3_col.p, s3_col.data(), Kh * sizeof(int), cudaMemcpyHostToDevice));
  MEM_CUDA(cudaMemcpy(g.special_pix.p, special.data(), 4 * sizeof(int), cudaMemcpyHostToDevice));
  g.N = N; g.Nh = Nh; g.Kh = Kh; g.Kr = Kr; g.n_special = n_special; g.n1_blocks = n1_blocks;
  g.K3 = K3; g.n3_blocks = n3_blocks; g.ldz = 32LL * (2 * n1_blocks + n3_blocks);
  g.filter_type = filter_type; g.filter_order = filter_order; g.filter_Qc = Qc;
  return 0;
}

// rows_only: batch x N one-dimensional transforms of length N (the row pass of the fused low-pass, lowpass.cu)
int fft_get(mem_ctx* ctx, int N, int batch, FftPlan* out, bool rows_only) {
  long long key = ((long long)N << 24) + batch + (rows_only ? (1LL << 60) : 0);
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end()) {
    *out = it->second;
    return 0;
  }
  FftPlan pl;
  int n[2] = {N, N};
  const int rank = rows_only ? 1 : 2, count = rows_only ? batch * N : batch;
  size_t ws1 = 0, ws2 = 0;
  MEM_CUFFT(cufftCreate(&pl.r2c));
  MEM_CUFFT(cufftSetAutoAllocation(pl.r2c, 0));
  MEM_CUFFT(cufftMakePlanMany(pl.r2c, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, count, &ws1));
  MEM_CUFFT(cufftCreate(&pl.c2r));
  MEM_CUFFT(cufftSetAutoAllocation(pl.c2r, 0));
  MEM_CUFFT(cufftMakePlanMany(pl.c2r, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, count, &ws2));
  MEM_CHECK(ctx->fft_work.ensure(std::max(ws1, ws2)));
  // the work area may have moved: re-attach it to every plan
  ctx->plans[key] = pl;
  for (auto& kv : ctx->plans) {
    MEM_CUFFT(cufftSetWorkArea(kv.second.r2c, ctx->fft_work.p));
    MEM_CUFFT(cufftSetWorkArea(kv.second.c2r, ctx->fft_work.p));
  }
  *out = pl;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

// a5 (:286-293): spectrum *= ifftshift(G)/N^2
__global__ void k_specmul(float2* __restrict__ spec, const float* __restrict__ G, int Kh, size_t total) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const float g = G[e % Kh];
    float2 v = spec[e];
    v.x *= g;
    v.y *= g;
    spec[e] = v;
  }
}

// a8 (ctemh_cryoFrank.py:24-44) at the distinct radii: C[i][b].
// gamma in fp64, reduced mod 2 pi in fp64, sin/cos in fp32 on the reduced argument (|err| ~ 1e-7).
struct CtfConst { double w1, w2_per_df, k2_scale, env_scale, ampc; };
__device__ __forceinline__ float ctf_eval(const CtfConst& cc, double df, int r2) {
  const double k2 = cc.k2_scale * (double)r2;
  const double g = (0.5 * cc.w1 * k2 - cc.w2_per_df * df) * k2;
  const double gr = g - 6.283185307179586476925 * rint(g * 0.15915494309189533576888);
  float s, c;
  sincosf((float)gr, &s, &c);
  float v = s - (float)cc.ampc * c;
  if (cc.env_scale != 0.0) v *= expf((float)(-k2 * cc.env_scale));
  return v;
}
__global__ void k_ctf_bins(const double* __restrict__ df, const int* __restrict__ r2_of_bin, float* __restrict__ cbin,
                           int Kr, CtfConst cc) {
  const int i = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Kr) return;
  cbin[(size_t)i * Kr + b] = ctf_eval(cc, df[i], r2_of_bin[b]);
}

// full-precision CTF field for the output record (:339, stored ifftshift-ed, flattened by :393)
__global__ void k_ctf_full(const double* __restrict__ df, double* __restrict__ out, int N, CtfConst cc) {
  const int i = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * N) return;
  const int ky = e / N, kx = e - ky * N;
  const int h = (N + 1) / 2;
  const int fy = ky < h ? ky : ky - N, fx = kx < h ? kx : kx - N;
  const double k2 = cc.k2_scale * (double)(fy * fy + fx * fx);
  const double g = (0.5 * cc.w1 * k2 - cc.w2_per_df * df[i]) * k2;
  double v = sin(g) - cc.ampc * cos(g);
  if (cc.env_scale != 0.0) v *= exp(-k2 * cc.env_scale);
  out[(size_t)i * N * N + e] = v;
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - hi);
}

// a12 operands, radial part: S1[i][b] = C_i(b)^2 / 4, S2[i][b] = P_i(b) = sum_{|k|^2 = r2(b)} |F_i(k)|^2
// (full spectrum = half spectrum with weight 2 on the non-self-conjugate columns), summed in a fixed order.
// Columns [Kr, Kr+n_special) carry the purely real self-conjugate pixels: S1 = -x/4, S2 = x, x = C*Re F.
__global__ void k_operands_radial(const float2* __restrict__ spec, const float2* __restrict__ Mspec,
                                  const float* __restrict__ cbin,
                                  const int* __restrict__ bin_start, const int* __restrict__ bin_pix,
                                  const int* __restrict__ bin_of_pix, const int* __restrict__ special_pix,
                                  float* __restrict__ zhi, float* __restrict__ zlo, int N, int Nh, int Kh, int Kr,
                                  int n_special, int n1_blocks, int64_t ldz) {
  const int i = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int w1 = 32 * n1_blocks;
  if (b >= w1) return;
  const float2* F = spec + (size_t)i * Kh;
  float s1 = 0.0f, s2 = 0.0f;
  if (b < Kr) {
    const float c = cbin[(size_t)i * Kr + b];
    s1 = 0.25f * c * c;
    float acc = 0.0f;
    const bool even = (N & 1) == 0;
    for (int q = bin_start[b]; q < bin_start[b + 1]; ++q) {
      const int p = bin_pix[q];
      const int kx = p % Nh;
      const float2 f = F[p], mm = Mspec[p];
      const float gx = fmaf(-c, mm.x, f.x), gy = fmaf(-c, mm.y, f.y);   // residual after removing C_i * M
      const float m = gx * gx + gy * gy;
      acc += ((kx == 0) || (even && kx == N / 2)) ? m : 2.0f * m;
    }
    s2 = acc;
  } else if (b < Kr + n_special) {
    const int p = special_pix[b - Kr];
    const float cs = cbin[(size_t)i * Kr + bin_of_pix[p]];
    const float x = cs * fmaf(-cs, Mspec[p].x, F[p].x);
    s1 = -0.25f * x;
    s2 = x;
  }
  float h, l;
  split_tf32(s1, h, l);
  zhi[(size_t)i * ldz + b] = h;
  zlo[(size_t)i * ldz + b] = l;
  split_tf32(s2, h, l);
  zhi[(size_t)i * ldz + w1 + b] = h;
  zlo[(size_t)i * ldz + w1 + b] = l;
}

// Same operands, one block per image.  Pass 1 reads the spectrum coalesced, two rows (ky = a and N - a, which share
// |k|^2 and therefore the CTF value) per folded row, and leaves w (|G(a,kx)|^2 + |G(N-a,kx)|^2), G = F - C M, in
// shared memory: (N/2+1) x Nh floats, 66 KB at N = 256, so three blocks share an SM.  Pass 2 gathers every bin's
// entries from shared memory in a fixed order.  The global gather of the kernel above uses a quarter of every
// 32-byte sector and one CTF lookup per pixel; this one is bound by the spectrum read.
constexpr int RADIAL_THREADS = 768;      // measured: 512 -> 768 threads, fft_ctf_operands 1.164 -> 1.127 ms at C4 (2 CTAs / SM either way)
// The same pass also writes the S3 operands A = C (F - C M) of every pixel it visits (hi/lo split, the job of
// k_operands_s3 below) and, if FLIP, leaves sign(C) F in place for the C2R of :346-347 — the spectrum is read once
// for all three operand groups.
template <bool FLIP, int RT = RADIAL_THREADS, int UU = 2>
__global__ void __launch_bounds__(RT, (RT <= 768 ? 2 : 1))
k_operands_radial_sm(float2* __restrict__ spec, const float2* __restrict__ Mspec, const float* __restrict__ cbin,
                     const int* __restrict__ fold_bin, const int* __restrict__ fold_start,
                     const int* __restrict__ fold_ent, const int* __restrict__ bin_of_pix,
                     const int* __restrict__ special_pix, const int* __restrict__ s3_col, float* __restrict__ zhi,
                     float* __restrict__ zlo, int N, int Nh, int Na, int Kh, int Kr, int n_special, int n1_blocks,
                     int64_t ldz, const int* __restrict__ pd_of) {
  extern __shared__ float pw[];
  const int i = blockIdx.x;
  if (pd_of) Mspec += (size_t)pd_of[i] * Kh;       // batched PDs: one common component per PD
  const int w1 = 32 * n1_blocks;
  float2* F = spec + (size_t)i * Kh;
  float* zh = zhi + (size_t)i * ldz;
  float* zl = zlo + (size_t)i * ldz;
  const int Kq = Na * Nh;
  float* cb = pw + Kq;                       // this image's CTF row, staged so the per-entry lookups stay on chip
  for (int b = threadIdx.x; b < Kr; b += RT) cb[b] = cbin[(size_t)i * Kr + b];
  __syncthreads();
  const int nyq = (N & 1) ? -1 : N / 2;
  const int step_x = RT % Nh, step_a = RT / Nh;
  int a = threadIdx.x / Nh, kx = threadIdx.x % Nh;
  // four entries per thread and trip, every load issued before the first use (the kernel is latency-bound otherwise)
  constexpr int U = UU;
  for (int e0 = threadIdx.x; e0 < Kq; e0 += U * RT) {
    float2 f[U], f2[U], mm[U], m2[U];
    int bn[U], sc[U], sc2[U], pe2[U];
    float wself[U], wpart[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * RT;
      const bool valid = e < Kq;
      const int ec = valid ? e : 0;
      const bool part = valid && a != 0 && 2 * a != N;
      const int e2 = part ? (N - a) * Nh + kx : ec;     // no partner row: re-read the entry itself with weight 0
      const float w = (kx == 0 || kx == nyq) ? 1.0f : 2.0f;
      wself[u] = w;
      wpart[u] = part ? w : 0.0f;
      bn[u] = fold_bin[ec];
      sc[u] = valid ? s3_col[ec] : -1;
      sc2[u] = part ? s3_col[e2] : -1;
      pe2[u] = part ? e2 : -1;
      f[u] = F[ec];
      mm[u] = Mspec[ec];
      f2[u] = F[e2];
      m2[u] = Mspec[e2];
      kx += step_x;
      a += step_a;
      if (kx >= Nh) { kx -= Nh; ++a; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * RT;
      if (e < Kq) {
        const float c = cb[bn[u]];
        const float gx = fmaf(-c, mm[u].x, f[u].x), gy = fmaf(-c, mm[u].y, f[u].y);
        const float hx = fmaf(-c, m2[u].x, f2[u].x), hy = fmaf(-c, m2[u].y, f2[u].y);
        pw[e] = wself[u] * (gx * gx + gy * gy) + wpart[u] * (hx * hx + hy * hy);
        if (sc[u] >= 0) {
          float2 h, l;
          split_tf32(c * gx, h.x, l.x);
          split_tf32(c * gy, h.y, l.y);
          *reinterpret_cast<float2*>(zh + sc[u]) = h;
          *reinterpret_cast<float2*>(zl + sc[u]) = l;
        }
        if (sc2[u] >= 0) {
          float2 h, l;
          split_tf32(c * hx, h.x, l.x);
          split_tf32(c * hy, h.y, l.y);
          *reinterpret_cast<float2*>(zh + sc2[u]) = h;
          *reinterpret_cast<float2*>(zl + sc2[u]) = l;
        }
        if (FLIP) {
          const float sg = (c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : 0.0f);
          F[e] = make_float2(sg * f[u].x, sg * f[u].y);
          if (pe2[u] >= 0) F[pe2[u]] = make_float2(sg * f2[u].x, sg * f2[u].y);
        }
      }
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < w1; b += RT) {
    float s1 = 0.0f, s2 = 0.0f;
    if (b < Kr) {
      const float c = cb[b];
      s1 = 0.25f * c * c;
      float acc = 0.0f;
      const int q1 = fold_start[b + 1];
      for (int q = fold_start[b]; q < q1; ++q) acc += pw[fold_ent[q]];
      s2 = acc;
    } else if (b < Kr + n_special) {
      const int p = special_pix[b - Kr];
      const float cs = cb[bin_of_pix[p]];
      float fx = F[p].x;
      if (FLIP) fx *= (cs > 0.0f) ? 1.0f : ((cs < 0.0f) ? -1.0f : 0.0f);   // pass 1 left sign(C) F here; sign^2 = 1
      const float x = cs * fmaf(-cs, Mspec[p].x, fx);
      s1 = -0.25f * x;
      s2 = x;
    }
    float h, l;
    split_tf32(s1, h, l);
    zh[b] = h;
    zl[b] = l;
    split_tf32(s2, h, l);
    zh[w1 + b] = h;
    zl[w1 + b] = l;
  }
}

// a11 partial sums (pass 1 over the spectra).  Thread = one half-spectrum pixel, loops over the images of
// its group (grid.y groups): sum C*F (for the common component M), sum C*Fw (Wiener numerator; Fw = spectrum
// of the unmasked image when msk2 is used), sum C^2, sum sign(C)*F — all in fp64.
__global__ void __launch_bounds__(256) k_spec_sums(const float2* __restrict__ spec, const float2* __restrict__ specw,
                                                   const float* __restrict__ cbin, const int* __restrict__ bin_of_pix,
                                                   double2* __restrict__ part_cf, double2* __restrict__ part_cfw,
                                                   double* __restrict__ part_c2, double2* __restrict__ part_fl,
                                                   int nS, int Kh, int Kr, int per_group, int img_stride) {
  // img_stride > 1: the sums run over every img_stride-th image only (nS = number of images visited).  Used when the sums
  // serve nothing but the common component M of the contraction operands: D is invariant under F_i -> F_i - C_i M for ANY
  // common M, so M only has to remove the bulk of the shared signal, which a subset estimates as well.
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Kh) return;
  const int g = blockIdx.y;
  const int i0 = g * per_group, i1 = min(nS, i0 + per_group);
  const int b = bin_of_pix[p];
  double2 scf = make_double2(0, 0), scw = make_double2(0, 0), sfl = make_double2(0, 0);
  double sc2 = 0;
  constexpr int U = 4;                       // loads of four images in flight per thread (latency, not HBM, binds otherwise)
  for (int ib = i0; ib < i1; ib += U) {
    float cu[U];
    float2 fu[U], fwu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = (size_t)min(ib + u, i1 - 1) * img_stride;
      cu[u] = cbin[i * Kr + b];
      fu[u] = spec[i * Kh + p];
      if (specw) fwu[u] = specw[i * Kh + p];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ib + u < i1) {
        const float c = cu[u];
        const float2 f = fu[u];
        const float sg = (c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : 0.0f);
        scf.x += (double)(c * f.x);
        scf.y += (double)(c * f.y);
        if (specw) {
          scw.x += (double)(c * fwu[u].x);
          scw.y += (double)(c * fwu[u].y);
        }
        sc2 += (double)(c * c);
        sfl.x += (double)(sg * f.x);
        sfl.y += (double)(sg * f.y);
      }
    }
  }
  part_cf[(size_t)g * Kh + p] = scf;
  if (specw) part_cfw[(size_t)g * Kh + p] = scw;
  part_c2[(size_t)g * Kh + p] = sc2;
  part_fl[(size_t)g * Kh + p] = sfl;
}

// The common component of EVERY PD of a batch in one launch: M_p = sum_i C_i F_i / sum_i C_i^2 over the images of PD p
// (grid.y = PD; the per-PD k_spec_sums + k_avg_spectra pair of the single-PD path costs two launches per PD, which is what a
// batch of reference-sized PDs spends its time on).  fp64 sums, four loads in flight per thread.
__global__ void __launch_bounds__(256) k_mspec_batch(const float2* __restrict__ spec, const float* __restrict__ cbin,
                                                     const int* __restrict__ bin_of_pix, const int* __restrict__ pd_start,
                                                     float2* __restrict__ Mspec, int Kh, int Kr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Kh) return;
  const int i0 = pd_start[blockIdx.y], i1 = pd_start[blockIdx.y + 1];
  const int b = bin_of_pix[p];
  double2 scf = make_double2(0, 0);
  double sc2 = 0;
  constexpr int U = 4;
  for (int ib = i0; ib < i1; ib += U) {
    float cu[U];
    float2 fu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = (size_t)min(ib + u, i1 - 1);
      cu[u] = cbin[i * Kr + b];
      fu[u] = spec[i * Kh + p];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ib + u < i1) {
        scf.x += (double)(cu[u] * fu[u].x);
        scf.y += (double)(cu[u] * fu[u].y);
        sc2 += (double)(cu[u] * cu[u]);
      }
    }
  }
  Mspec[(size_t)blockIdx.y * Kh + p] = (sc2 > 1e-30) ? make_float2((float)(scf.x / sc2), (float)(scf.y / sc2)) : make_float2(0.0f, 0.0f);
}

// a12 operands, S3 part (pass 2) + a10 phase flip.  The distance is invariant under F_i -> F_i - C_i M for
// ANY common M (C_j (C_i M) - C_i (C_j M) = 0); with M = sum C F / sum C^2 the residuals carry only noise and
// conformational signal, which removes the catastrophic cancellation in a + b - 2c for similar images
// (DESIGN.md §3).  A = C * (F - C M) -> Z (hi/lo);  spec <- sign(C) * F in place for the C2R of :346-347.
__global__ void __launch_bounds__(256) k_operands_s3(float2* __restrict__ spec, const float2* __restrict__ Mspec,
                                                     const float* __restrict__ cbin, const int* __restrict__ bin_of_pix,
                                                     const int* __restrict__ s3_col, float* __restrict__ zhi,
                                                     float* __restrict__ zlo, int nS, int Kh, int Kr, int64_t ldz,
                                                     int per_group, int write_z, int flip) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Kh) return;
  const int g = blockIdx.y;
  const int i0 = g * per_group, i1 = min(nS, i0 + per_group);
  const int b = bin_of_pix[p];
  const int col = s3_col[p];
  const float2 m = Mspec[p];
  // four images per trip, all loads ahead of the (possibly aliasing, in-place) stores: with one image per trip the
  // kernel is bound by load latency, not by HBM
  constexpr int U = 4;
  for (int ib = i0; ib < i1; ib += U) {
    float c[U];
    float2 f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = min(ib + u, i1 - 1);
      c[u] = cbin[(size_t)i * Kr + b];
      f[u] = spec[(size_t)i * Kh + p];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = ib + u;
      if (i < i1) {
        if (flip) {
          const float sg = (c[u] > 0.0f) ? 1.0f : ((c[u] < 0.0f) ? -1.0f : 0.0f);
          spec[(size_t)i * Kh + p] = make_float2(sg * f[u].x, sg * f[u].y);
        }
        if (write_z && col >= 0) {
          const float gx = fmaf(-c[u], m.x, f[u].x), gy = fmaf(-c[u], m.y, f[u].y);
          float2 h, l;
          split_tf32(c[u] * gx, h.x, l.x);
          split_tf32(c[u] * gy, h.y, l.y);
          *reinterpret_cast<float2*>(zhi + (size_t)i * ldz + col) = h;
          *reinterpret_cast<float2*>(zlo + (size_t)i * ldz + col) = l;
        }
      }
    }
  }
}

// zero the K padding at the end of S3 (columns [64*n1 + 2*K3, ldz))
__global__ void k_zero_tail(float* __restrict__ zhi, float* __restrict__ zlo, int nS, int64_t ldz, int from) {
  const int i = blockIdx.x;
  for (int c = from + threadIdx.x; c < ldz; c += blockDim.x) {
    zhi[(size_t)i * ldz + c] = 0.0f;
    zlo[(size_t)i * ldz + c] = 0.0f;
  }
}

// a11 (:353-367, :422-430): avgspec[0] = sum_i C_i Fw_i / wd,  wd = -(sum_i C_i^2 + 1/5);  avgspec[1] = sum_i sign(C_i) F_i;
// Mspec = sum_i C_i F_i / sum_i C_i^2  (the common component removed from the contraction operands)
__global__ void k_avg_spectra(const double2* __restrict__ part_cf, const double2* __restrict__ part_cfw,
                              const double* __restrict__ part_c2, const double2* __restrict__ part_fl,
                              float2* __restrict__ avgspec, float2* __restrict__ Mspec, int Kh, int G) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Kh) return;
  double2 cf = make_double2(0, 0), cw = make_double2(0, 0), fl = make_double2(0, 0);
  double c2 = 0;
  for (int g = 0; g < G; ++g) {
    const double2 a = part_cf[(size_t)g * Kh + p], b = part_fl[(size_t)g * Kh + p];
    cf.x += a.x; cf.y += a.y; fl.x += b.x; fl.y += b.y;
    if (part_cfw) {
      const double2 w = part_cfw[(size_t)g * Kh + p];
      cw.x += w.x; cw.y += w.y;
    }
    c2 += part_c2[(size_t)g * Kh + p];
  }
  if (!part_cfw) cw = cf;
  const double wd = -(c2 + 1.0 / 5.0);
  avgspec[p] = make_float2((float)(cw.x / wd), (float)(cw.y / wd));
  avgspec[Kh + p] = make_float2((float)fl.x, (float)fl.y);
  Mspec[p] = (c2 > 1e-30) ? make_float2((float)(cf.x / c2), (float)(cf.y / c2)) : make_float2(0.0f, 0.0f);
}

// a10/a13: imgAllFlip = irfft(...)/N^2 (scale in place) and per-group partial sums of squares (:400)
__global__ void __launch_bounds__(256) k_flip_scale_intensity(float* __restrict__ flipimg, double* __restrict__ part_int,
                                                              int nS, int NN, int per_group, float scale) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NN) return;
  const int g = blockIdx.y;
  const int i0 = g * per_group, i1 = min(nS, i0 + per_group);
  double acc = 0;
  for (int i = i0; i < i1; ++i) {
    const float v = flipimg[(size_t)i * NN + e] * scale;
    flipimg[(size_t)i * NN + e] = v;
    acc += (double)v * v;
  }
  part_int[(size_t)g * NN + e] = acc;
}

// final small outputs: imgAvg, imgAvgFlip (x msk2 / (nS N^2)), imgAllIntensity
__global__ void k_small_outputs(const float* __restrict__ avgimg, const double* __restrict__ part_int,
                                const uint8_t* __restrict__ msk2, float* __restrict__ imgAvg,
                                float* __restrict__ imgAvgFlip, float* __restrict__ intensity, int NN, int G, int nS) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NN) return;
  const float m = (msk2 ? (msk2[e] ? 1.0f : 0.0f) : 1.0f) / ((float)nS * (float)NN);
  if (imgAvg) imgAvg[e] = avgimg[e] * m;
  if (imgAvgFlip) imgAvgFlip[e] = avgimg[NN + e] * m;
  if (intensity) {
    double a = 0;
    for (int g = 0; g < G; ++g) a += part_int[(size_t)g * NN + e];
    intensity[e] = (float)(a / nS);
  }
}

// ------------------------------------------------------------------------------------------------
// host pipeline
// ------------------------------------------------------------------------------------------------
static int run_fft(mem_ctx* ctx, int N, int nS, bool forward, float* real, float2* cplx, cudaStream_t st,
                   bool rows_only = false) {
  const int Kh = N * (N / 2 + 1);
  const int BMAX = 1024;
  for (int i0 = 0; i0 < nS; i0 += BMAX) {
    const int b = std::min(BMAX, nS - i0);
    FftPlan pl;
    MEM_CHECK(fft_get(ctx, N, b, &pl, rows_only));
    if (forward) {
      MEM_CUFFT(cufftSetStream(pl.r2c, st));
      MEM_CUFFT(cufftExecR2C(pl.r2c, real + (size_t)i0 * N * N, reinterpret_cast<cufftComplex*>(cplx + (size_t)i0 * Kh)));
    } else {
      MEM_CUFFT(cufftSetStream(pl.c2r, st));
      MEM_CUFFT(cufftExecC2R(pl.c2r, reinterpret_cast<cufftComplex*>(cplx + (size_t)i0 * Kh), real + (size_t)i0 * N * N));
    }
    ctx->launches += 1;   // library call, counted once
  }
  return 0;
}

static CtfConst make_ctf_const(const mem_pd_params* prm) {
  CtfConst cc;
  const double wav = 12.3986 / sqrt((2 * 511.0 + prm->EkV) * prm->EkV);
  cc.w1 = M_PI * (prm->Cs * 1.0e7) * wav * wav * wav;
  cc.w2_per_df = M_PI * wav;
  const double half = prm->N / 2.0;
  cc.k2_scale = 1.0 / (half * half) / (4.0 * prm->pix_size * prm->pix_size);   // k = Q / (2 pix), Q = r / (N/2)
  if (isinf(prm->gaussEnv)) {
    cc.env_scale = 0.0;
  } else {
    const double sig = prm->gaussEnv / sqrt(2 * log(2.0));
    cc.env_scale = 1.0 / (2 * sig * sig);
  }
  cc.ampc = prm->AmpContrast;
  return cc;
}

int pd_distance_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, cudaStream_t st) {
  const int nS = prm->nS, N = prm->N;
  if (nS <= 0 || N <= 3) {
    set_error("bad shape nS=%d N=%d", nS, N);
    return 1;
  }
  if (prm->relion_shift && (!io->shift || prm->transposed)) {
    set_error("relion_shift needs io->shift and picture-orientation images (transposed = 0)");
    return 1;
  }
  MEM_CHECK(geometry_prepare(ctx, N, prm->filter_type, prm->filter_order, prm->filter_Qc));
  const Geometry& g = ctx->geom;
  const size_t NN = (size_t)N * N, Kh = g.Kh;
  const size_t img_bytes = (size_t)nS * NN * sizeof(float);
  const size_t spec_bytes = (size_t)nS * Kh * sizeof(float2);
  MEM_CHECK(ctx->imgA.ensure(img_bytes));
  MEM_CHECK(ctx->imgB.ensure(img_bytes));
  MEM_CHECK(ctx->spec.ensure(spec_bytes));
  MEM_CHECK(ctx->rot_cs.ensure((size_t)(nS + 1) * sizeof(double2)));
  MEM_CHECK(ctx->cbin.ensure((size_t)nS * g.Kr * sizeof(float)));
  float* imgAll = io->imgAll;
  if (!imgAll) {
    MEM_CHECK(ctx->imgAll.ensure(img_bytes));
    imgAll = ctx->imgAll.as<float>();
  }
  float* imgFlip = io->imgAllFlip;
  const bool need_flip = io->imgAllFlip || io->imgAllIntensity;
  if (need_flip && !imgFlip) {
    MEM_CHECK(ctx->imgFlip.ensure(img_bytes));
    imgFlip = ctx->imgFlip.as<float>();
  }
  const bool want_knn = prm->knn_k > 0 && io->knn_idx && io->knn_val && !prm->avg_only;
  if (prm->knn_k > 0 && !prm->avg_only && (!io->knn_idx || !io->knn_val)) {
    set_error("knn_k = %d but io.knn_idx / io.knn_val are not set", prm->knn_k);
    return 1;
  }
  const bool want_D = (io->D != nullptr || want_knn) && !prm->avg_only;
  if (want_D) {
    MEM_CHECK(ctx->zhi.ensure((size_t)nS * g.ldz * sizeof(float)));
    MEM_CHECK(ctx->zlo.ensure((size_t)nS * g.ldz * sizeof(float)));
  }
  const int per_group = 64;
  const int G = (nS + per_group - 1) / per_group;
  MEM_CHECK(ctx->part_cf.ensure((size_t)G * Kh * sizeof(double2)));
  MEM_CHECK(ctx->part_c2.ensure((size_t)G * Kh * sizeof(double)));
  MEM_CHECK(ctx->part_fl.ensure((size_t)G * Kh * sizeof(double2)));
  MEM_CHECK(ctx->part_int.ensure((size_t)G * NN * sizeof(double)));
  MEM_CHECK(ctx->avgspec.ensure(3 * Kh * sizeof(float2)));
  MEM_CHECK(ctx->avgimg.ensure(2 * NN * sizeof(float)));
  if (io->msk2) {
    MEM_CHECK(ctx->spec2.ensure(spec_bytes));
    MEM_CHECK(ctx->part_cfw.ensure((size_t)G * Kh * sizeof(double2)));
  }

  float* A = ctx->imgA.as<float>();
  float* B = ctx->imgB.as<float>();
  float2* spec = ctx->spec.as<float2>();
  double2* cs = ctx->rot_cs.as<double2>();

  MEM_CUDA(cudaEventRecord(ctx->ev[0], st));
  // ---- a2/a3 ingest + normalise, a5 low-pass -> B
  int rows_done = 0;
  const float* picture = io->raw;                 // what the ingest reads: the raw stack, or its RELION-shifted copy
  int transposed = prm->transposed;
  if (prm->relion_shift) {   // :263-264 shift(order=3, mode='wrap') before the flip / normalisation
    MEM_CHECK(shift_run(ctx, io->raw, io->shift, A, B, nS, N, st));
    picture = B;
    transposed = 0;
  }
  if (own_fft(ctx, N)) {
    // One pass of ours turns the raw particles into row-transformed half spectra (ingest, moments and the R2C row pass
    // fused), one more does the whole column pass — FFT, normalisation, * G, inverse FFT — and a third brings the rows
    // back (C2R, annular mask, row pass of the first spline prefilter): three crossings of HBM instead of nine.
    MEM_CHECK(ctx->stats.ensure((size_t)nS * sizeof(float2)));
    float2* stats = ctx->stats.as<float2>();
    MEM_CHECK(ingest_rowfft_run(ctx, picture, io->flip, spec, stats, nS, N, transposed, st));
    MEM_CHECK(colfilter_run(ctx, spec, g.Gtab.as<float>(), stats, nS, N, st));
    // inverse row transform fused with the annular mask and the row pass of the first spline prefilter -> A
    MEM_CHECK(rowifft_prefilter_run(ctx, spec, A, nS, N, st));
    rows_done = 1;
  } else if (colpass_supported(N) && !ctx->cufft_lowpass) {
    // N = 320: rows through cuFFT's batched 1-D plans, the whole column pass (FFT, * G, inverse FFT) in one kernel of ours
    if (!ctx->cufft_rows320) {                       // own row passes as well: three crossings, as at 256
      MEM_CHECK(ctx->stats.ensure((size_t)nS * sizeof(float2)));
      float2* stats = ctx->stats.as<float2>();
      MEM_CHECK(rows320_forward_run(ctx, picture, io->flip, spec, stats, nS, transposed, 0, st));
      MEM_CHECK(colpass_run(ctx, spec, g.Gtab.as<float>(), stats, nS, N, 0, st));
      MEM_CHECK(rows320_inverse_run(ctx, spec, A, nS, st));
      rows_done = 1;
    } else {
      MEM_CHECK(ingest_run(ctx, picture, io->flip, A, nS, N, transposed, st));
      MEM_CHECK(run_fft(ctx, N, nS, true, A, spec, st, true));
      MEM_CHECK(colpass_run(ctx, spec, g.Gtab.as<float>(), nullptr, nS, N, 0, st));
      MEM_CHECK(run_fft(ctx, N, nS, false, B, spec, st, true));
    }
  } else {
    MEM_CHECK(ingest_run(ctx, picture, io->flip, A, nS, N, transposed, st));
    MEM_CHECK(run_fft(ctx, N, nS, true, A, spec, st));
    const size_t total = (size_t)nS * Kh;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    MEM_LAUNCH(ctx, k_specmul, grid, 256, 0, st, spec, g.Gtab.as<float>(), (int)Kh, total);
    MEM_CHECK(run_fft(ctx, N, nS, false, B, spec, st));
  }
  MEM_CUDA(cudaEventRecord(ctx->ev[1], st));
  // ---- a7 alignment: two periodic cubic-spline rotations
  MEM_CHECK(align_run(ctx, A, B, imgAll, io->psi_deg, prm->psi_p_deg, cs, io->msk2, nS, N, st, rows_done));
  MEM_CUDA(cudaEventRecord(ctx->ev[2], st));
  // ---- a10 FFT of img*msk2 (and of img for the Wiener average when they differ)
  float2* specw = nullptr;
  if (io->msk2) {
    MEM_CHECK(run_fft(ctx, N, nS, true, B, spec, st));
    specw = ctx->spec2.as<float2>();
    MEM_CHECK(run_fft(ctx, N, nS, true, imgAll, specw, st));
  } else if (own_fft(ctx, N) && !ctx->cufft_a10) {
    MEM_CHECK(fft2_forward_run(ctx, imgAll, spec, nS, N, st));
  } else if (colpass_supported(N) && !ctx->cufft_lowpass && !ctx->cufft_a10) {
    if (!ctx->cufft_rows320) MEM_CHECK(rows320_forward_run(ctx, imgAll, nullptr, spec, nullptr, nS, 0, 1, st));
    else MEM_CHECK(run_fft(ctx, N, nS, true, imgAll, spec, st, true));
    MEM_CHECK(colpass_run(ctx, spec, nullptr, nullptr, nS, N, 1, st));
  } else {
    MEM_CHECK(run_fft(ctx, N, nS, true, imgAll, spec, st));
  }
  // ---- a8 CTF at the distinct radii, a12 operands, a10 flip, a11 partial sums
  const CtfConst cc = make_ctf_const(prm);
  MEM_LAUNCH(ctx, k_ctf_bins, dim3((g.Kr + 255) / 256, nS), 256, 0, st, io->df, g.r2_of_bin.as<int>(),
             ctx->cbin.as<float>(), g.Kr, cc);
  float* zhi = ctx->zhi.as<float>();
  float* zlo = ctx->zlo.as<float>();
  double2* part_cfw = specw ? ctx->part_cfw.as<double2>() : (double2*)nullptr;
  float2* Mspec = ctx->avgspec.as<float2>() + 2 * Kh;
  // the Wiener / flip averages need the sums over every image; when only D (or the neighbour lists) is asked for, the sums
  // feed nothing but M, and every 8th image of a large PD estimates the shared signal as well (>= 128 images)
  const bool sums_for_M_only = !(io->imgAvg || io->imgAvgFlip || io->imgAllIntensity) && !ctx->full_sums;
  const int sum_stride = (sums_for_M_only && nS >= 1024) ? 8 : 1;
  const int nSum = (nS + sum_stride - 1) / sum_stride;
  const int Gs = (nSum + per_group - 1) / per_group;
  MEM_LAUNCH(ctx, k_spec_sums, dim3((g.Kh + 255) / 256, Gs), 256, 0, st, spec, specw, ctx->cbin.as<float>(),
             g.bin_of_pix.as<int>(), ctx->part_cf.as<double2>(), part_cfw, ctx->part_c2.as<double>(),
             ctx->part_fl.as<double2>(), nSum, g.Kh, g.Kr, per_group, sum_stride);
  MEM_LAUNCH(ctx, k_avg_spectra, (g.Kh + 255) / 256, 256, 0, st, ctx->part_cf.as<double2>(), part_cfw,
             ctx->part_c2.as<double>(), ctx->part_fl.as<double2>(), ctx->avgspec.as<float2>(), Mspec, g.Kh, Gs);
  bool s3_done = false;      // the shared-memory radial kernel also writes S3 and the flipped spectra
  if (want_D) {
    const size_t pw_bytes = ((size_t)g.Na * g.Nh + g.Kr) * sizeof(float);
    if (pw_bytes <= 200 * 1024) {
      auto kern = need_flip ? k_operands_radial_sm<true> : k_operands_radial_sm<false>;
      int rthreads = RADIAL_THREADS;
      if (!need_flip) {                                  // experiment switch (mem_ctx_set_option "radial_variant")
        switch (ctx->radial_variant) {
          case 1: kern = k_operands_radial_sm<false, 640, 2>; rthreads = 640; break;
          case 2: kern = k_operands_radial_sm<false, 512, 2>; rthreads = 512; break;
          case 3: kern = k_operands_radial_sm<false, 512, 4>; rthreads = 512; break;
          case 4: kern = k_operands_radial_sm<false, 1024, 2>; rthreads = 1024; break;
          case 5: kern = k_operands_radial_sm<false, 512, 1>; rthreads = 512; break;
          default: break;
        }
      }
      MEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_bytes));
      MEM_LAUNCH(ctx, kern, nS, rthreads, pw_bytes, st, spec, Mspec, ctx->cbin.as<float>(),
                 g.fold_bin.as<int>(), g.fold_start.as<int>(), g.fold_ent.as<int>(), g.bin_of_pix.as<int>(),
                 g.special_pix.as<int>(), g.s3_col.as<int>(), zhi, zlo, N, g.Nh, g.Na, g.Kh, g.Kr, g.n_special,
                 g.n1_blocks, g.ldz, (const int*)nullptr);
      s3_done = true;
    } else {   // boxes above ~ 450 px: the folded power map does not fit shared memory
      MEM_LAUNCH(ctx, k_operands_radial, dim3((32 * g.n1_blocks + 255) / 256, nS), 256, 0, st, spec, Mspec,
                 ctx->cbin.as<float>(), g.bin_start.as<int>(), g.bin_pix.as<int>(), g.bin_of_pix.as<int>(),
                 g.special_pix.as<int>(), zhi, zlo, N, g.Nh, g.Kh, g.Kr, g.n_special, g.n1_blocks, g.ldz);
    }
    const int from = 64 * g.n1_blocks + 2 * g.K3;
    if (from < g.ldz) MEM_LAUNCH(ctx, k_zero_tail, nS, 64, 0, st, zhi, zlo, nS, g.ldz, from);
  }
  if ((want_D || need_flip) && !s3_done)
    MEM_LAUNCH(ctx, k_operands_s3, dim3((g.Kh + 255) / 256, G), 256, 0, st, spec, Mspec, ctx->cbin.as<float>(),
               g.bin_of_pix.as<int>(), g.s3_col.as<int>(), zhi, zlo, nS, g.Kh, g.Kr, g.ldz, per_group,
               want_D ? 1 : 0, need_flip ? 1 : 0);
  MEM_CUDA(cudaEventRecord(ctx->ev[3], st));
  // ---- a10/a11/a13 phase-flipped images, averages, intensity
  if (need_flip) {
    MEM_CHECK(run_fft(ctx, N, nS, false, imgFlip, spec, st));
    MEM_LAUNCH(ctx, k_flip_scale_intensity, dim3(((int)NN + 255) / 256, G), 256, 0, st, imgFlip,
               ctx->part_int.as<double>(), nS, (int)NN, per_group, 1.0f / (float)NN);
  }
  if (io->imgAvg || io->imgAvgFlip || io->imgAllIntensity) {
    MEM_CHECK(run_fft(ctx, N, 2, false, ctx->avgimg.as<float>(), ctx->avgspec.as<float2>(), st));
    MEM_LAUNCH(ctx, k_small_outputs, ((int)NN + 255) / 256, 256, 0, st, ctx->avgimg.as<float>(),
               ctx->part_int.as<double>(), io->msk2, io->imgAvg, io->imgAvgFlip,
               need_flip ? io->imgAllIntensity : (float*)nullptr, (int)NN, G, nS);
  }
  if (io->CTF)
    MEM_LAUNCH(ctx, k_ctf_full, dim3(((int)NN + 255) / 256, nS), 256, 0, st, io->df, io->CTF, N, cc);
  MEM_CUDA(cudaEventRecord(ctx->ev[4], st));
  // ---- a12 contraction
  if (want_D) {
    mem_contract_shape shp;
    shp.nS = nS; shp.n1_blocks = g.n1_blocks; shp.n3_blocks = g.n3_blocks; shp.ldz = g.ldz;
    KnnOut knn;
    knn.k = prm->knn_k; knn.idx = io->knn_idx; knn.val = io->knn_val;
    MEM_CHECK(contract_run(ctx, &shp, zhi, zlo, io->D, prm->contraction, prm->k_chunk_blocks, prm->split_k, st,
                           want_knn ? &knn : nullptr));
  }
  MEM_CUDA(cudaEventRecord(ctx->ev[5], st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// A GROUP of PDs in one call (GetDistancesS2.py:94-120 hands the worker one PD at a time; the tessellation's PDs are
// small — 117..450 particles in the demo, at most 2,000 — so a PD alone cannot fill 148 SMs).  The per-image stages
// (ingest, low-pass, alignment, FFT, CTF) run ONCE over the concatenated stack of all PDs: an image's second rotation
// angle is looked up from its PD.  The per-PD reductions (Wiener sums, common component) and the operand writer run per
// PD on their slice, and ONE grouped tcgen05 launch contracts all PDs (contract_tc_grouped).  Results are those of the
// one-PD-at-a-time path (same kernels on the same values; the second rotation goes through the per-image kernel).
//   prm->nS = images of all PDs; io->raw / flip / psi_deg / df concatenated; io->D = the n_pd matrices back to back
//   (PD g at float offset sum_{h<g} nS_h^2); io->imgAll optional.  No msk2, no RELION shift, no flip / average outputs.
int pd_distance_batch_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, int n_pd, const int* pd_start,
                             const double* psi_p_deg, cudaStream_t st) {
  const int nS = prm->nS, N = prm->N;
  if (n_pd < 1 || pd_start[0] != 0 || pd_start[n_pd] != nS) {
    set_error("batch: pd_start must run from 0 to nS over n_pd >= 1 PDs");
    return 1;
  }
  if (io->msk2 || prm->relion_shift || io->imgAllFlip || io->CTF || io->imgAvg || io->imgAvgFlip || io->imgAllIntensity ||
      io->knn_idx || prm->avg_only || !io->D) {
    set_error("batch: only D (and imgAll) are produced; no msk2 / RELION shift / flip / average / kNN outputs");
    return 1;
  }
  if (!rotate_fast_supported(N)) {
    set_error("batch: box size %d is not a multiple of 32", N);
    return 1;
  }
  MEM_CHECK(geometry_prepare(ctx, N, prm->filter_type, prm->filter_order, prm->filter_Qc));
  const Geometry& g = ctx->geom;
  const size_t NN = (size_t)N * N, Kh = g.Kh;
  const size_t img_bytes = (size_t)nS * NN * sizeof(float), spec_bytes = (size_t)nS * Kh * sizeof(float2);
  MEM_CHECK(ctx->imgA.ensure(img_bytes));
  MEM_CHECK(ctx->imgB.ensure(img_bytes));
  MEM_CHECK(ctx->spec.ensure(spec_bytes));
  MEM_CHECK(ctx->rot_cs.ensure((size_t)(nS + 1) * sizeof(double2)));
  MEM_CHECK(ctx->cbin.ensure((size_t)nS * g.Kr * sizeof(float)));
  MEM_CHECK(ctx->zhi.ensure((size_t)nS * g.ldz * sizeof(float)));
  MEM_CHECK(ctx->zlo.ensure((size_t)nS * g.ldz * sizeof(float)));
  float* imgAll = io->imgAll;
  if (!imgAll) {
    MEM_CHECK(ctx->imgAll.ensure(img_bytes));
    imgAll = ctx->imgAll.as<float>();
  }
  int max_n = 0;
  for (int p = 0; p < n_pd; ++p) max_n = std::max(max_n, pd_start[p + 1] - pd_start[p]);
  const int per_group = 64;
  const int Gmax = (max_n + per_group - 1) / per_group;
  MEM_CHECK(ctx->part_cf.ensure((size_t)Gmax * Kh * sizeof(double2)));
  MEM_CHECK(ctx->part_c2.ensure((size_t)Gmax * Kh * sizeof(double)));
  MEM_CHECK(ctx->part_fl.ensure((size_t)Gmax * Kh * sizeof(double2)));
  MEM_CHECK(ctx->avgspec.ensure(3 * Kh * sizeof(float2)));
  // per-image PD index, per-PD psi_p, per-image second-rotation angle and pitch id
  const size_t b_pd = ((size_t)nS * sizeof(int) + 255) & ~(size_t)255, b_pp = ((size_t)n_pd * sizeof(double) + 255) & ~(size_t)255;
  const size_t b_cs = ((size_t)nS * sizeof(double2) + 255) & ~(size_t)255;
  const size_t b_id = ((size_t)nS + 255) & ~(size_t)255;
  MEM_CHECK(ctx->batch_aux.ensure(b_pd + b_pp + b_cs + b_id + (size_t)(n_pd + 1) * sizeof(int)));
  uint8_t* aux = ctx->batch_aux.as<uint8_t>();
  int* pd_of = reinterpret_cast<int*>(aux);
  double* d_pp = reinterpret_cast<double*>(aux + b_pd);
  double2* cs2 = reinterpret_cast<double2*>(aux + b_pd + b_pp);
  uint8_t* pid2 = aux + b_pd + b_pp + b_cs;
  int* d_start = reinterpret_cast<int*>(aux + b_pd + b_pp + b_cs + b_id);
  {
    std::vector<int> h(nS);
    for (int p = 0; p < n_pd; ++p)
      for (int i = pd_start[p]; i < pd_start[p + 1]; ++i) h[i] = p;
    MEM_CUDA(cudaMemcpyAsync(pd_of, h.data(), (size_t)nS * sizeof(int), cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaMemcpyAsync(d_pp, psi_p_deg, (size_t)n_pd * sizeof(double), cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaMemcpyAsync(d_start, pd_start, (size_t)(n_pd + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaStreamSynchronize(st));                  // host temporaries
  }
  float* A = ctx->imgA.as<float>();
  float* B = ctx->imgB.as<float>();
  float2* spec = ctx->spec.as<float2>();
  double2* cs = ctx->rot_cs.as<double2>();
  MEM_CUDA(cudaEventRecord(ctx->ev[0], st));
  int rows_done = 0;
  if (own_fft(ctx, N)) {
    MEM_CHECK(ctx->stats.ensure((size_t)nS * sizeof(float2)));
    float2* stats = ctx->stats.as<float2>();
    MEM_CHECK(ingest_rowfft_run(ctx, io->raw, io->flip, spec, stats, nS, N, prm->transposed, st));
    MEM_CHECK(colfilter_run(ctx, spec, g.Gtab.as<float>(), stats, nS, N, st));
    MEM_CHECK(rowifft_prefilter_run(ctx, spec, A, nS, N, st));
    rows_done = 1;
  } else if (colpass_supported(N) && !ctx->cufft_lowpass) {
    if (!ctx->cufft_rows320) {
      MEM_CHECK(ctx->stats.ensure((size_t)nS * sizeof(float2)));
      float2* stats = ctx->stats.as<float2>();
      MEM_CHECK(rows320_forward_run(ctx, io->raw, io->flip, spec, stats, nS, prm->transposed, 0, st));
      MEM_CHECK(colpass_run(ctx, spec, g.Gtab.as<float>(), stats, nS, N, 0, st));
      MEM_CHECK(rows320_inverse_run(ctx, spec, A, nS, st));
      rows_done = 1;
    } else {
      MEM_CHECK(ingest_run(ctx, io->raw, io->flip, A, nS, N, prm->transposed, st));
      MEM_CHECK(run_fft(ctx, N, nS, true, A, spec, st, true));
      MEM_CHECK(colpass_run(ctx, spec, g.Gtab.as<float>(), nullptr, nS, N, 0, st));
      MEM_CHECK(run_fft(ctx, N, nS, false, B, spec, st, true));
    }
  } else {
    MEM_CHECK(ingest_run(ctx, io->raw, io->flip, A, nS, N, prm->transposed, st));
    MEM_CHECK(run_fft(ctx, N, nS, true, A, spec, st));
    const size_t total = (size_t)nS * Kh;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    MEM_LAUNCH(ctx, k_specmul, grid, 256, 0, st, spec, g.Gtab.as<float>(), (int)Kh, total);
    MEM_CHECK(run_fft(ctx, N, nS, false, B, spec, st));
  }
  MEM_CUDA(cudaEventRecord(ctx->ev[1], st));
  MEM_CHECK(rotate_angles_batch_run(ctx, pd_of, d_pp, cs2, pid2, nS, st));
  MEM_CHECK(align_batch_run(ctx, A, B, imgAll, io->psi_deg, cs, cs2, pid2, nS, N, st, rows_done));
  MEM_CUDA(cudaEventRecord(ctx->ev[2], st));
  if (own_fft(ctx, N) && !ctx->cufft_a10) {
    MEM_CHECK(fft2_forward_run(ctx, imgAll, spec, nS, N, st));
  } else if (colpass_supported(N) && !ctx->cufft_lowpass && !ctx->cufft_a10) {
    if (!ctx->cufft_rows320) MEM_CHECK(rows320_forward_run(ctx, imgAll, nullptr, spec, nullptr, nS, 0, 1, st));
    else MEM_CHECK(run_fft(ctx, N, nS, true, imgAll, spec, st, true));
    MEM_CHECK(colpass_run(ctx, spec, nullptr, nullptr, nS, N, 1, st));
  } else {
    MEM_CHECK(run_fft(ctx, N, nS, true, imgAll, spec, st));
  }
  const CtfConst cc = make_ctf_const(prm);
  MEM_LAUNCH(ctx, k_ctf_bins, dim3((g.Kr + 255) / 256, nS), 256, 0, st, io->df, g.r2_of_bin.as<int>(),
             ctx->cbin.as<float>(), g.Kr, cc);
  float* zhi = ctx->zhi.as<float>();
  float* zlo = ctx->zlo.as<float>();
  float2* Mspec = ctx->avgspec.as<float2>() + 2 * Kh;
  const size_t pw_bytes = ((size_t)g.Na * g.Nh + g.Kr) * sizeof(float);
  const bool radial_sm = pw_bytes <= 200 * 1024;
  if (radial_sm) MEM_CUDA(cudaFuncSetAttribute(k_operands_radial_sm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_bytes));
  std::vector<float*> Dp(n_pd);
  size_t doff = 0;
  for (int p = 0; p < n_pd; ++p) {
    Dp[p] = io->D + doff;
    doff += (size_t)(pd_start[p + 1] - pd_start[p]) * (pd_start[p + 1] - pd_start[p]);
  }
  if (radial_sm) {
    // one launch for the common components of all PDs, one for the operands of all images (M looked up by the image's PD)
    MEM_CHECK(ctx->mspec_all.ensure((size_t)n_pd * Kh * sizeof(float2)));
    float2* M_all = ctx->mspec_all.as<float2>();
    MEM_LAUNCH(ctx, k_mspec_batch, dim3((g.Kh + 255) / 256, n_pd), 256, 0, st, spec, ctx->cbin.as<float>(), g.bin_of_pix.as<int>(),
               d_start, M_all, g.Kh, g.Kr);
    MEM_LAUNCH(ctx, k_operands_radial_sm<false>, nS, RADIAL_THREADS, pw_bytes, st, spec, M_all, ctx->cbin.as<float>(),
               g.fold_bin.as<int>(), g.fold_start.as<int>(), g.fold_ent.as<int>(), g.bin_of_pix.as<int>(),
               g.special_pix.as<int>(), g.s3_col.as<int>(), zhi, zlo, N, g.Nh, g.Na, g.Kh, g.Kr, g.n_special, g.n1_blocks, g.ldz,
               (const int*)pd_of);
  } else {
    for (int p = 0; p < n_pd; ++p) {
      const int s0 = pd_start[p], n = pd_start[p + 1] - s0;
      const int G = (n + per_group - 1) / per_group;
      float2* spec_p = spec + (size_t)s0 * Kh;
      float* cbin_p = ctx->cbin.as<float>() + (size_t)s0 * g.Kr;
      float* zhi_p = zhi + (size_t)s0 * g.ldz;
      float* zlo_p = zlo + (size_t)s0 * g.ldz;
      MEM_LAUNCH(ctx, k_spec_sums, dim3((g.Kh + 255) / 256, G), 256, 0, st, spec_p, (const float2*)nullptr, cbin_p,
                 g.bin_of_pix.as<int>(), ctx->part_cf.as<double2>(), (double2*)nullptr, ctx->part_c2.as<double>(),
                 ctx->part_fl.as<double2>(), n, g.Kh, g.Kr, per_group, 1);
      MEM_LAUNCH(ctx, k_avg_spectra, (g.Kh + 255) / 256, 256, 0, st, ctx->part_cf.as<double2>(), (double2*)nullptr,
                 ctx->part_c2.as<double>(), ctx->part_fl.as<double2>(), ctx->avgspec.as<float2>(), Mspec, g.Kh, G);
      MEM_LAUNCH(ctx, k_operands_radial, dim3((32 * g.n1_blocks + 255) / 256, n), 256, 0, st, spec_p, Mspec, cbin_p,
                 g.bin_start.as<int>(), g.bin_pix.as<int>(), g.bin_of_pix.as<int>(), g.special_pix.as<int>(), zhi_p, zlo_p,
                 N, g.Nh, g.Kh, g.Kr, g.n_special, g.n1_blocks, g.ldz);
      MEM_LAUNCH(ctx, k_operands_s3, dim3((g.Kh + 255) / 256, G), 256, 0, st, spec_p, Mspec, cbin_p, g.bin_of_pix.as<int>(),
                 g.s3_col.as<int>(), zhi_p, zlo_p, n, g.Kh, g.Kr, g.ldz, per_group, 1, 0);
    }
  }
  const int from = 64 * g.n1_blocks + 2 * g.K3;
  if (from < g.ldz) MEM_LAUNCH(ctx, k_zero_tail, nS, 64, 0, st, zhi, zlo, nS, g.ldz, from);
  MEM_CUDA(cudaEventRecord(ctx->ev[3], st));
  MEM_CUDA(cudaEventRecord(ctx->ev[4], st));
  mem_contract_shape shp;
  shp.nS = nS; shp.n1_blocks = g.n1_blocks; shp.n3_blocks = g.n3_blocks; shp.ldz = g.ldz;
  MEM_CHECK(contract_tc_grouped(ctx, &shp, n_pd, pd_start, zhi, zlo, Dp.data(), st));
  MEM_CUDA(cudaEventRecord(ctx->ev[5], st));
  return 0;
}

// a8 alone (ctemh_cryoFrank.op, ctemh_cryoFrank.py:24-44, as stored in the record: ifftshift-ed, float64, flattened):
// df [nS] and out [nS][N*N] are HOST pointers; the same kernel as inside the PD pipeline, so a record that stores df
// instead of the 8 N^2 bytes per particle of the CTF field gets bit-identical values back.  Synchronises.
int ctf_host(mem_ctx* ctx, const mem_pd_params* prm, const double* df, double* out) {
  const int nS = prm->nS, N = prm->N;
  if (nS < 1 || N < 2) {
    set_error("ctf: bad shape (nS=%d N=%d)", nS, N);
    return 1;
  }
  cudaStream_t st = ctx->stream;
  const size_t NN = (size_t)N * N;
  MEM_CHECK(ctx->df.ensure((size_t)nS * sizeof(double)));
  MEM_CHECK(ctx->ctf64.ensure((size_t)nS * NN * sizeof(double)));
  MEM_CUDA(cudaMemcpyAsync(ctx->df.p, df, (size_t)nS * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_LAUNCH(ctx, k_ctf_full, dim3(((int)NN + 255) / 256, nS), 256, 0, st, ctx->df.as<double>(), ctx->ctf64.as<double>(), N,
             make_ctf_const(prm));
  MEM_CUDA(cudaMemcpyAsync(out, ctx->ctf64.p, (size_t)nS * NN * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace mem
