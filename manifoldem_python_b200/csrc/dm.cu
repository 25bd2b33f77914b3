// dm.cu — device side of the diffusion-map front end (DMembeddingII.op, rows a15-a18):
// kNN lists, OR-symmetrised graph, Ferguson log-sum sweep, Gaussian-kernel Laplacian.
// All HBM/L2-bound integer + fp64 work; no tensor cores.
#include "common.cuh"

#include <math.h>
#include <algorithm>
#include <vector>

namespace mem {

// ---------------------------------------------------------------------------------------------
// a15  DMembeddingII.initialize :43-57 — per point the k smallest distances, self first.
// One CTA per point; (key, index) pairs bitonic-sorted in shared memory, ties broken by index.
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(1024) k_knn_sort(const T* __restrict__ D, int nS, int P, int k,
                                                   int* __restrict__ idx, double* __restrict__ val) {
  extern __shared__ uint8_t sm_raw[];
  double* key = reinterpret_cast<double*>(sm_raw);
  int* id = reinterpret_cast<int*>(sm_raw + (size_t)P * sizeof(double));
  const int i = blockIdx.x;
  const T* row = D + (size_t)i * nS;        // D symmetric: column i == row i
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    double v = INFINITY;
    if (j < nS) v = (j == i) ? -INFINITY : (double)row[j];   // D[iS,iS] = -inf, :48
    key[j] = v;
    id[j] = j;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double a = key[lo], b = key[hi];
        const int ia = id[lo], ib = id[hi];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == up) {
          key[lo] = b; key[hi] = a;
          id[lo] = ib; id[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int a = threadIdx.x; a < k; a += blockDim.x) {
    idx[(size_t)i * k + a] = id[a];
    val[(size_t)i * k + a] = (a == 0) ? 0.0 : key[a];   // yVal1[0,iS] = 0, :54
  }
}

// Rows longer than one shared-memory sort (nS > 16,384): the same bitonic network over P = 2^m padded entries, cut
// into chunks of KNN_CH.  Every compare-exchange with stride < KNN_CH stays inside a chunk and runs in shared
// memory; the log2(P / KNN_CH) levels above it take one pass per stride >= KNN_CH through the (key, index) rows in
// global memory.  Same (value, index) order as k_knn_sort, so ties resolve identically.
constexpr int KNN_CH = 16384;

template <class T>
__device__ __forceinline__ void knn_cmpx(T* key, int* id, int lo, int hi, bool up) {
  const T a = key[lo], b = key[hi];
  const int ia = id[lo], ib = id[hi];
  const bool gt = (a > b) || (a == b && ia > ib);
  if (gt == up) {
    key[lo] = b; key[hi] = a;
    id[lo] = ib; id[hi] = ia;
  }
}

// grid (P / KNN_CH, nS).  first != 0: load the chunk from D and run every level up to KNN_CH; otherwise load it from
// the workspace and finish level `size` (strides KNN_CH/2 .. 1).  last != 0: the chunk holding ranks < k writes idx/val.
template <class T>
__global__ void __launch_bounds__(1024) k_knn_chunk(const T* __restrict__ D, int nS, int P, int k, int size, int first,
                                                    int last, T* __restrict__ wkey, int* __restrict__ wid,
                                                    int* __restrict__ idx, double* __restrict__ val) {
  extern __shared__ uint8_t sm_raw[];
  T* key = reinterpret_cast<T*>(sm_raw);
  int* id = reinterpret_cast<int*>(sm_raw + (size_t)KNN_CH * sizeof(T));
  const int i = blockIdx.y, base = blockIdx.x * KNN_CH;
  T* gk = wkey + (size_t)i * P + base;
  int* gi = wid + (size_t)i * P + base;
  if (first) {
    const T* row = D + (size_t)i * nS;
    for (int t = threadIdx.x; t < KNN_CH; t += blockDim.x) {
      const int j = base + t;
      T v = (T)INFINITY;
      if (j < nS) v = (j == i) ? (T)(-INFINITY) : row[j];
      key[t] = v;
      id[t] = j;
    }
  } else {
    for (int t = threadIdx.x; t < KNN_CH; t += blockDim.x) {
      key[t] = gk[t];
      id[t] = gi[t];
    }
  }
  __syncthreads();
  for (int sz = first ? 2 : size; sz <= (first ? KNN_CH : size); sz <<= 1) {
    for (int stride = min(sz, KNN_CH) >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (KNN_CH >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        knn_cmpx(key, id, lo, lo + stride, (((base + lo) & sz) == 0));
      }
      __syncthreads();
    }
  }
  if (last) {
    for (int t = threadIdx.x; t < KNN_CH; t += blockDim.x) {
      const int a = base + t;
      if (a < k) {
        idx[(size_t)i * k + a] = id[t];
        val[(size_t)i * k + a] = (a == 0) ? 0.0 : (double)key[t];
      }
    }
  } else {
    for (int t = threadIdx.x; t < KNN_CH; t += blockDim.x) {
      gk[t] = key[t];
      gi[t] = id[t];
    }
  }
}

// one compare-exchange pass of level `size` at a stride >= KNN_CH; grid (P / 2 / 256, nS)
template <class T>
__global__ void __launch_bounds__(256) k_knn_global_step(T* __restrict__ wkey, int* __restrict__ wid, int P, int size,
                                                         int stride) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (P >> 1)) return;
  const int lo = 2 * t - (t & (stride - 1));
  knn_cmpx(wkey + (size_t)blockIdx.y * P, wid + (size_t)blockIdx.y * P, lo, lo + stride, ((lo & size) == 0));
}

template <class T>
static int knn_chunked(mem_ctx* ctx, const T* D, int nS, int P, int k, int* idx, double* val, cudaStream_t st) {
  const size_t rows = (size_t)nS * P;
  MEM_CHECK(ctx->knn_ws.ensure(rows * (sizeof(T) + sizeof(int))));
  T* wkey = ctx->knn_ws.as<T>();
  int* wid = reinterpret_cast<int*>(wkey + rows);
  const size_t smem = (size_t)KNN_CH * (sizeof(T) + sizeof(int));
  MEM_CUDA(cudaFuncSetAttribute(k_knn_chunk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 gc(P / KNN_CH, nS), gs((P / 2 + 255) / 256, nS);
  MEM_LAUNCH(ctx, k_knn_chunk<T>, gc, 1024, smem, st, D, nS, P, k, 0, 1, 0, wkey, wid, idx, val);
  for (int size = 2 * KNN_CH; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride >= KNN_CH; stride >>= 1)
      MEM_LAUNCH(ctx, k_knn_global_step<T>, gs, 256, 0, st, wkey, wid, P, size, stride);
    MEM_LAUNCH(ctx, k_knn_chunk<T>, gc, 1024, smem, st, D, nS, P, k, size, 0, size == P ? 1 : 0, wkey, wid, idx, val);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a15 for k << nS: selection instead of a full sort.  One CTA per point; the row becomes order-preserving unsigned
// keys in shared memory, an 8-bit-digit radix select (most significant digit first, 256-bin shared-memory
// histograms) finds the k-th smallest key, the k winners are gathered — exact ties at the threshold in index
// order — and only those are sorted, by (value, index) like k_knn_sort, so both kernels return identical lists.
// WS: the row is assembled on the fly from the split-K partial tiles of the contraction
//     D[i][j] = 4 * sum_s ws[s][min(i,j)][max(i,j)]   (same operation order as k_contract_finalize: bit-identical),
// so the nS x nS matrix is never written (BASELINE config 3: "kNN epilogue only").
// ---------------------------------------------------------------------------------------------
constexpr int KSEL_T = 256;
constexpr int KSEL_MAXK = 2048;

__device__ __forceinline__ uint32_t ord_key(float f) {
  if (f == 0.0f) f = 0.0f;                               // -0 and +0 compare equal in the sort: one key
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ uint64_t ord_key(double f) {
  if (f == 0.0) f = 0.0;
  const uint64_t u = (uint64_t)__double_as_longlong(f);
  return u ^ ((u >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_val(uint32_t u) {
  u ^= (u >> 31) ? 0x80000000u : 0xffffffffu;
  return (double)__uint_as_float(u);
}
__device__ __forceinline__ double ord_val(uint64_t u) {
  u ^= (u >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull;
  return __longlong_as_double((long long)u);
}
template <class T> struct OrdOf { typedef uint32_t U; };
template <> struct OrdOf<double> { typedef uint64_t U; };

template <class T, bool WS>
__global__ void __launch_bounds__(KSEL_T) k_knn_select(const T* __restrict__ D, const float* __restrict__ ws, int ldw,
                                                       int nslices, int nS, int nS_pad, int k, int P2,
                                                       int* __restrict__ idx, double* __restrict__ val) {
  typedef typename OrdOf<T>::U U;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  U* ukey = reinterpret_cast<U*>(sm_raw);               // [nS_pad] the whole row
  U* selk = ukey + nS_pad;                              // [P2] winners
  int* seli = reinterpret_cast<int*>(selk + P2);        // [P2]
  __shared__ int hist[256];
  __shared__ int wtot[KSEL_T / 32];
  __shared__ int s_bin, s_rem, s_cnt;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  // ---- the row as ordered keys; D[i][i] = -inf (DMembeddingII.py:48)
  for (int j = tid; j < nS; j += KSEL_T) {
    T v;
    if (WS) {
      const size_t a = (size_t)min(i, j), b = (size_t)max(i, j);
      float acc = 0.0f;
      for (int s = 0; s < nslices; ++s) acc += ws[((size_t)s * ldw + a) * ldw + b];
      v = (T)(4.0f * acc);
    } else {
      v = D[(size_t)i * nS + j];
    }
    ukey[j] = ord_key(j == i ? (T)(-INFINITY) : v);
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  // ---- radix select: after the last digit `prefix` is the k-th smallest key and `rem` how many of the
  // entries equal to it belong to the list
  U prefix = 0, mask = 0;
  int rem = k;
  for (int shift = 8 * (int)sizeof(U) - 8; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    for (int j = tid; j < nS; j += KSEL_T) {
      const U u = ukey[j];
      if ((u & mask) == prefix) atomicAdd(&hist[(int)((u >> shift) & 0xff)], 1);
    }
    __syncthreads();
    const int h = hist[tid];
    int inc = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wtot[wrp] = inc;
    __syncthreads();
    for (int w = 0; w < wrp; ++w) inc += wtot[w];
    const int exc = inc - h;
    if (h > 0 && exc < rem && rem <= inc) {              // exactly one digit holds the k-th smallest
      s_bin = tid;
      s_rem = rem - exc;
    }
    __syncthreads();
    prefix |= (U)s_bin << shift;
    mask |= (U)0xff << shift;
    rem = s_rem;
  }
  const U thr = prefix;
  const int n_less = k - rem;
  // ---- winners below the threshold (any order: they are sorted afterwards)
  for (int j = tid; j < nS; j += KSEL_T) {
    const U u = ukey[j];
    if (u < thr) {
      const int slot = atomicAdd(&s_cnt, 1);
      selk[slot] = u;
      seli[slot] = j;
    }
  }
  // ---- entries equal to the threshold: the `rem` smallest indices (ordered compaction, chunk by chunk)
  int running = 0;
  for (int c0 = 0; c0 < nS && running < rem; c0 += KSEL_T) {
    const int j = c0 + tid;
    const bool f = j < nS && ukey[j] == thr;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wtot[wrp] = __popc(b);
    __syncthreads();
    int off = running, tot = 0;
#pragma unroll
    for (int w = 0; w < KSEL_T / 32; ++w) {
      const int c = wtot[w];
      if (w < wrp) off += c;
      tot += c;
    }
    if (f) {
      const int r = off + __popc(b & ((1u << lane) - 1u));
      if (r < rem) {
        selk[n_less + r] = thr;
        seli[n_less + r] = j;
      }
    }
    running += tot;
    __syncthreads();
  }
  for (int a = k + tid; a < P2; a += KSEL_T) {          // padding sorts last
    selk[a] = ~(U)0;
    seli[a] = 0x7fffffff;
  }
  __syncthreads();
  // ---- sort the winners by (value, index)
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P2 >> 1); t += KSEL_T) {
        const int lo = 2 * t - (t & (stride - 1));
        knn_cmpx(selk, seli, lo, lo + stride, ((lo & size) == 0));
      }
      __syncthreads();
    }
  }
  for (int a = tid; a < k; a += KSEL_T) {
    idx[(size_t)i * k + a] = seli[a];
    val[(size_t)i * k + a] = (a == 0) ? 0.0 : ord_val(selk[a]);   // yVal1[0,iS] = 0, :54
  }
}

static int g_knn_mode = 0;   // 0 = auto, 1 = always the full sort, 2 = selection whenever it fits
void knn_set_mode(int mode) { g_knn_mode = mode; }

// shared memory of a selection launch, 0 when the row / list does not fit (the caller sorts instead)
template <class T>
static size_t knn_select_smem(int nS, int k, int* nS_pad, int* P2) {
  typedef typename OrdOf<T>::U U;
  if (k > KSEL_MAXK) return 0;
  int p = 32;
  while (p < k) p <<= 1;
  *P2 = p;
  *nS_pad = (nS + 3) & ~3;
  const size_t bytes = (size_t)*nS_pad * sizeof(U) + (size_t)p * (sizeof(U) + sizeof(int));
  return bytes <= 200 * 1024 ? bytes : 0;
}
template <class T>
static bool knn_use_select(int nS, int k) {
  int a, b;
  if (g_knn_mode == 1 || knn_select_smem<T>(nS, k, &a, &b) == 0) return false;
  return g_knn_mode == 2 || 4 * k <= nS;               // a list that is most of the row: sort the row
}
template <class T, bool WS>
static int knn_select_launch(mem_ctx* ctx, const T* D, const float* ws, int ldw, int nslices, int nS, int k, int* idx,
                             double* val, cudaStream_t st) {
  int nS_pad = 0, P2 = 0;
  const size_t smem = knn_select_smem<T>(nS, k, &nS_pad, &P2);
  if (!smem || k < 1 || k > nS) {
    set_error("knn selection: unsupported shape (k=%d nS=%d)", k, nS);
    return 1;
  }
  MEM_CUDA(cudaFuncSetAttribute((k_knn_select<T, WS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MEM_LAUNCH(ctx, (k_knn_select<T, WS>), nS, KSEL_T, smem, st, D, ws, ldw, nslices, nS, nS_pad, k, P2, idx, val);
  return 0;
}
// kNN lists straight from the split-K workspace of the contraction; returns 2 when the shape needs the assembled D
int knn_from_workspace(mem_ctx* ctx, const float* ws, int ldw, int nslices, int nS, int k, int* idx, double* val,
                       cudaStream_t st) {
  if (k < 1 || k > nS) {
    set_error("knn: need 1 <= k <= nS (k=%d nS=%d)", k, nS);
    return 1;
  }
  if (!knn_use_select<float>(nS, k)) return 2;
  return knn_select_launch<float, true>(ctx, (const float*)nullptr, ws, ldw, nslices, nS, k, idx, val, st);
}

template <class T>
static int knn_device_t(mem_ctx* ctx, const T* D, int nS, int k, int* idx, double* val, cudaStream_t st) {
  if (k < 1 || k > nS) {
    set_error("knn: need 1 <= k <= nS (k=%d nS=%d)", k, nS);
    return 1;
  }
  if (knn_use_select<T>(nS, k))
    return knn_select_launch<T, false>(ctx, D, (const float*)nullptr, 0, 0, nS, k, idx, val, st);
  int P = 1;
  while (P < nS) P <<= 1;
  if (P > KNN_CH) return knn_chunked<T>(ctx, D, nS, P, k, idx, val, st);
  const size_t smem = (size_t)P * (sizeof(double) + sizeof(int));
  MEM_CUDA(cudaFuncSetAttribute(k_knn_sort<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = std::max(32, std::min(1024, P / 2));
  MEM_LAUNCH(ctx, k_knn_sort<T>, nS, threads, smem, st, D, nS, P, k, idx, val);
  return 0;
}
int knn_device(mem_ctx* ctx, const double* D, int nS, int k, int* idx, double* val, cudaStream_t st) {
  return knn_device_t<double>(ctx, D, nS, k, idx, val, st);
}
// same on the float32 D the distance stage leaves on the device (no host round trip of the N x N matrix)
int knn_device_f32(mem_ctx* ctx, const float* D, int nS, int k, int* idx, double* val, cudaStream_t st) {
  return knn_device_t<float>(ctx, D, nS, k, idx, val, st);
}

// ---------------------------------------------------------------------------------------------
// a16  DMembeddingII.op :113-140 — union kNN graph with value d^2, dense form, absent = -1.
//   pass 1: Y[i][j] = d^2 for directed list entries (zeros: flag matrix Zf[i][j] = 1)
//   pass 2: M[i][j] = a^2 + b^2 - a b, a = sqrt(Y[i][j]) or 0, b = sqrt(Y[j][i]) or 0;
//           entries whose result is 0 are absent unless the zero flag is set (value 0).
// ---------------------------------------------------------------------------------------------
__global__ void k_graph_scatter(const int* __restrict__ idx, const double* __restrict__ val, int nS, int k,
                                double* __restrict__ Y, uint8_t* __restrict__ Zf) {
  const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (e >= (size_t)nS * k) return;
  const int i = (int)(e / k);
  const int j = idx[e];
  const double v = val[e];
  if (v < 1e-6) Zf[(size_t)i * nS + j] = 1;
  else Y[(size_t)i * nS + j] = v;
}
__global__ void k_graph_combine(const double* __restrict__ Y, const uint8_t* __restrict__ Zf, int nS,
                                double* __restrict__ M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= nS) return;
  const double yij = Y[(size_t)i * nS + j], yji = Y[(size_t)j * nS + i];
  const double a = yij > 0 ? sqrt(yij) : 0.0, b = yji > 0 ? sqrt(yji) : 0.0;
  // separately rounded like NumPy's y**2 + (y**2).T - y*y.T (no FMA contraction: the graph is bit-exact)
  const double m = __dsub_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(a, b));
  double out = -1.0;
  if (m != 0.0) out = m;
  else if (Zf[(size_t)i * nS + j]) out = 0.0;
  M[(size_t)i * nS + j] = out;
}

int graph_dense_device(mem_ctx* ctx, const int* idx, const double* val, int nS, int k, double* M, cudaStream_t st) {
  const size_t nn = (size_t)nS * nS;
  MEM_CHECK(ctx->scratch.ensure(nn * sizeof(double) + nn));
  double* Y = ctx->scratch.as<double>();
  uint8_t* Zf = reinterpret_cast<uint8_t*>(Y + nn);
  MEM_CUDA(cudaMemsetAsync(Y, 0, nn * sizeof(double) + nn, st));
  const size_t tot = (size_t)nS * k;
  MEM_LAUNCH(ctx, k_graph_scatter, (unsigned)((tot + 255) / 256), 256, 0, st, idx, val, nS, k, Y, Zf);
  MEM_LAUNCH(ctx, k_graph_combine, dim3((nS + 255) / 256, nS), 256, 0, st, Y, Zf, nS, M);
  return 0;
}

// Row-major compaction of the graph entries (M >= 0) into a dense vector: for k < nS the Ferguson sweep then
// touches nnz <= 2 nS k values instead of nS^2 slots.  Deterministic order (row, then column).
__global__ void __launch_bounds__(256) k_row_counts(const double* __restrict__ M, int nS, int* __restrict__ counts) {
  __shared__ int red[8];
  const int i = blockIdx.x;
  int c = 0;
  for (int j = threadIdx.x; j < nS; j += 256) c += (M[(size_t)i * nS + j] >= 0.0);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    counts[i] = t;
  }
}
__global__ void __launch_bounds__(256) k_row_compact(const double* __restrict__ M, int nS, const long long* __restrict__ offs,
                                                     double* __restrict__ out) {
  __shared__ int wsum[8];
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long base = offs[i];
  for (int j0 = 0; j0 < nS; j0 += 256) {
    const int j = j0 + threadIdx.x;
    const double v = (j < nS) ? M[(size_t)i * nS + j] : -1.0;
    const bool keep = v >= 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += wsum[w];
      total += wsum[w];
    }
    if (keep) out[base + before + __popc(m & ((1u << lane) - 1u))] = v;
    base += total;
    __syncthreads();
  }
}
int graph_compact_device(mem_ctx* ctx, const double* M, int nS, double* out, long long* count) {
  cudaStream_t st = ctx->stream;
  const size_t off_bytes = ((size_t)nS * sizeof(int) + 15) & ~(size_t)15;
  MEM_CHECK(ctx->small_out.ensure(off_bytes + (size_t)nS * sizeof(long long)));
  int* d_counts = ctx->small_out.as<int>();
  MEM_LAUNCH(ctx, k_row_counts, nS, 256, 0, st, M, nS, d_counts);
  std::vector<int> counts(nS);
  MEM_CUDA(cudaMemcpyAsync(counts.data(), d_counts, nS * sizeof(int), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  std::vector<long long> offs(nS);
  long long tot = 0;
  for (int i = 0; i < nS; ++i) { offs[i] = tot; tot += counts[i]; }
  long long* d_offs = reinterpret_cast<long long*>(ctx->small_out.as<uint8_t>() + off_bytes);
  MEM_CUDA(cudaMemcpyAsync(d_offs, offs.data(), nS * sizeof(long long), cudaMemcpyHostToDevice, st));
  MEM_LAUNCH(ctx, k_row_compact, nS, 256, 0, st, M, nS, d_offs, out);
  MEM_CUDA(cudaStreamSynchronize(st));
  *count = tot;
  return 0;
}

// column sums: CTA = 32 columns x 32 row lanes; lane r adds rows r, r+32, ... in four independent chains, then the
// 32 partial sums of a column are added in a fixed order (deterministic; coalesced 256-byte row segments)
__global__ void __launch_bounds__(1024) k_colsum(const double* __restrict__ W, int rows, int cols,
                                                 double* __restrict__ d, int take_sqrt) {
  __shared__ double part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (j < cols) {
    int i = ty;
    for (; i + 96 < rows; i += 128) {
      a0 += W[(size_t)i * cols + j];
      a1 += W[(size_t)(i + 32) * cols + j];
      a2 += W[(size_t)(i + 64) * cols + j];
      a3 += W[(size_t)(i + 96) * cols + j];
    }
    for (; i < rows; i += 32) a0 += W[(size_t)i * cols + j];
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && j < cols) {
    double t = 0;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += part[r][tx];
    d[j] = take_sqrt ? sqrt(t) : t;
  }
}

// ---------------------------------------------------------------------------------------------
// a17  fergusonE.op :36-43 — for every eps: sum over present entries with d2/(2 eps) < thr of
// exp(-d2/(2 eps)).  Each thread keeps 8 distances in registers and sweeps all eps; per-eps block
// sums go to partial[block][eps], reduced over the blocks in a fixed order by k_colsum (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int FG_V = 8, FG_THREADS = 256, FG_ET = 32;
__global__ void __launch_bounds__(FG_THREADS) k_ferguson(const double* __restrict__ d2, size_t n,
                                                         const double* __restrict__ inv2eps, int nEps, double thr,
                                                         double* __restrict__ partial) {
  __shared__ double wsum[FG_THREADS / 32][FG_ET];
  double x[FG_V];
  const size_t base = (size_t)blockIdx.x * FG_THREADS * FG_V;
#pragma unroll
  for (int v = 0; v < FG_V; ++v) {
    const size_t j = base + (size_t)v * FG_THREADS + threadIdx.x;
    x[v] = (j < n) ? d2[j] : -1.0;   // negative = absent
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e0 = 0; e0 < nEps; e0 += FG_ET) {
    for (int ee = 0; ee < FG_ET && e0 + ee < nEps; ++ee) {
      const double s = inv2eps[e0 + ee];
      double acc = 0;
#pragma unroll
      for (int v = 0; v < FG_V; ++v) {
        const double d = x[v] * s;
        if (x[v] >= 0.0 && d < thr) acc += exp(-d);
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) wsum[warp][ee] = acc;
    }
    __syncthreads();
    if (threadIdx.x < FG_ET && e0 + threadIdx.x < nEps) {
      double t = 0;
      for (int w = 0; w < FG_THREADS / 32; ++w) t += wsum[w][threadIdx.x];
      partial[(size_t)blockIdx.x * nEps + e0 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// The sweep proper (nEps = 1,501).  For one graph entry only ~207 of the 1,501 eps give anything but exactly 0
// (past the cut d2/(2 eps) >= thr) or exactly 1 (d2/(2 eps) < 2^-54), and only ~46 of those need a real
// exponential: below d2/(2 eps) = 1e-3 a degree-4 Taylor polynomial is exact to 1e-17.  Which of the four classes an
// (entry, eps) pair falls in is monotone in d2, so the CTA first sorts its 2,048 entries in shared memory; for a tile
// of 32 consecutive eps (one per lane) three binary searches with the tile's largest / smallest 1/(2 eps) then cut the
// sorted chunk into [saturated | polynomial | exponential | cut].  The saturated entries are a count, the cut ones are
// skipped, the 8 warps share the two middle ranges (entries broadcast from shared memory, no cross-lane reduction, no
// per-entry tests) and their sums are added in a fixed order.  ~77 exp + ~184 polynomials per entry instead of 1,501
// exp.  Same partial[block][eps] layout as k_ferguson; both are reduced over the blocks in a fixed order by k_colsum.
constexpr int FT_V = 2048, FT_THREADS = 256, FT_WARPS = FT_THREADS / 32;
// exp(-x) for 0 <= x < 1e-3
__device__ __forceinline__ double exp_small(double x) {
  return fma(x, fma(x, fma(x, fma(x, 1.0 / 24.0, -1.0 / 6.0), 0.5), -1.0), 1.0);
}
// entries of the sorted chunk for which d * s < bound (monotone in d >= 0)
__device__ __forceinline__ int count_below(const double* sv, double s, double bound) {
  int lo = 0, hi = FT_V;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sv[mid] * s < bound) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(FT_THREADS) k_ferguson_sorted(const double* __restrict__ d2, size_t n,
                                                                const double* __restrict__ inv2eps, int nEps, double thr,
                                                                double* __restrict__ partial) {
  __shared__ double sv[FT_V];
  __shared__ double red[FT_WARPS][32];
  __shared__ int bounds[64][3];                     // per tile: end of saturated, of polynomial, of exponential range
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t base = (size_t)blockIdx.x * FT_V;
  for (int t = tid; t < FT_V; t += FT_THREADS) {
    const size_t j = base + t;
    const double v = (j < n) ? d2[j] : -1.0;
    sv[t] = v >= 0.0 ? v : INFINITY;                // absent (negative) entries sort to the end: always cut
  }
  __syncthreads();
  for (int size = 2; size <= FT_V; size <<= 1)      // bitonic sort, ascending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < FT_V / 2; t += FT_THREADS) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const double a = sv[lo], b = sv[hi];
        if ((a > b) == ((lo & size) == 0)) { sv[lo] = b; sv[hi] = a; }
      }
      __syncthreads();
    }
  const double tiny = 5.5511151231257827e-17;       // 2^-54: exp(-x) rounds to 1
  const bool shortcuts = thr >= 1e-3;               // both shortcuts assume the term is on the near side of the cut
  const int nTiles = (nEps + 31) / 32;
  for (int t0 = 0; t0 < nTiles; t0 += 64) {         // 64 tiles (2,048 eps) per round: one round for the 1,501-point grid
    const int tiles = min(64, nTiles - t0);
    if (tid < 3 * tiles) {
      const int tile = t0 + tid / 3, what = tid % 3;
      double smax = 0.0, smin = INFINITY;
      for (int e = tile * 32; e < min(nEps, tile * 32 + 32); ++e) {
        smax = fmax(smax, inv2eps[e]);
        smin = fmin(smin, inv2eps[e]);
      }
      int c;
      if (what == 0) c = shortcuts ? count_below(sv, smax, tiny) : 0;
      else if (what == 1) c = shortcuts ? count_below(sv, smax, 1e-3) : 0;
      else c = count_below(sv, smin, thr);
      bounds[tid / 3][what] = c;
    }
    __syncthreads();
    for (int tt = 0; tt < tiles; ++tt) {
      const int e = (t0 + tt) * 32 + lane;
      const double s = inv2eps[min(e, nEps - 1)];
      const int n_sat = bounds[tt][0], n_poly = max(bounds[tt][1], n_sat), n_exp = max(bounds[tt][2], n_poly);
      double acc = 0.0;
      for (int v = n_sat + warp; v < n_poly; v += FT_WARPS) acc += exp_small(sv[v] * s);
      for (int v = n_poly + warp; v < n_exp; v += FT_WARPS) {
        const double x = sv[v] * s;
        if (x < thr) acc += exp(-x);
      }
      red[warp][lane] = acc;
      __syncthreads();
      if (warp == 0 && e < nEps) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < FT_WARPS; ++w) t += red[w][lane];
        partial[(size_t)blockIdx.x * nEps + e] = t + (double)n_sat;
      }
      __syncthreads();
    }
  }
}

int ferguson_device(mem_ctx* ctx, const double* d2, int64_t n, const double* logEps, int nEps, double thr,
                    double* out_host) {
  cudaStream_t st = ctx->stream;
  std::vector<double> inv(nEps);
  for (int e = 0; e < nEps; ++e) inv[e] = 1.0 / (2.0 * exp(logEps[e]));
  const int nBlocks = (int)((n + (size_t)FG_THREADS * FG_V - 1) / ((size_t)FG_THREADS * FG_V));
  MEM_CHECK(ctx->small_out.ensure((size_t)(2 * nEps) * sizeof(double) + (size_t)nBlocks * nEps * sizeof(double)));
  double* d_inv = ctx->small_out.as<double>();
  double* d_out = d_inv + nEps;
  double* d_part = d_out + nEps;
  MEM_CUDA(cudaMemcpyAsync(d_inv, inv.data(), nEps * sizeof(double), cudaMemcpyHostToDevice, st));
  static_assert(FT_V == FG_THREADS * FG_V, "both sweep kernels cut the distances into the same chunks");
  if (nEps >= 64) MEM_LAUNCH(ctx, k_ferguson_sorted, nBlocks, FT_THREADS, 0, st, d2, (size_t)n, d_inv, nEps, thr, d_part);
  else MEM_LAUNCH(ctx, k_ferguson, nBlocks, FG_THREADS, 0, st, d2, (size_t)n, d_inv, nEps, thr, d_part);
  MEM_LAUNCH(ctx, k_colsum, (nEps + 31) / 32, 1024, 0, st, d_part, nBlocks, nEps, d_out, 0);   // fixed order
  std::vector<double> sums(nEps);
  MEM_CUDA(cudaMemcpyAsync(sums.data(), d_out, nEps * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  for (int e = 0; e < nEps; ++e) out_host[e] = log(sums[e]);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a18  slaplacianonFly.op :57-78, dense storage (absent entries stay 0):
//   W = exp(-d2/sigma^2);  W /= d_i d_j (d = column sums);  W /= sqrt(d'_i) sqrt(d'_j);  L = |W + W^T| / 2
// ---------------------------------------------------------------------------------------------
__global__ void k_lap_weights(const double* __restrict__ M, double* __restrict__ W, size_t nn, double inv_s2) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    const double m = M[e];
    W[e] = (m >= 0.0) ? exp(-m * inv_s2) : 0.0;
  }
}
__global__ void k_lap_scale(double* __restrict__ W, int nS, const double* __restrict__ d) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= nS) return;
  W[(size_t)i * nS + j] /= (d[i] * d[j]);
}
__global__ void k_lap_sym(const double* __restrict__ W, double* __restrict__ L, int nS) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bj + r, j = bi + tx;   // transposed tile
    t[r][tx] = (i < nS && j < nS) ? W[(size_t)i * nS + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + r, j = bj + tx;
    if (i < nS && j < nS) L[(size_t)i * nS + j] = fabs(W[(size_t)i * nS + j] + t[tx][r]) * 0.5;
  }
}

int laplacian_dense_device(mem_ctx* ctx, const double* M, int nS, double sigma, double* L, cudaStream_t st) {
  const size_t nn = (size_t)nS * nS;
  MEM_CHECK(ctx->scratch.ensure(nn * sizeof(double) + (size_t)nS * sizeof(double)));
  double* W = ctx->scratch.as<double>();
  double* d = W + nn;
  const int g1 = (int)std::min<size_t>((nn + 255) / 256, (size_t)ctx->sm_count * 32);
  MEM_LAUNCH(ctx, k_lap_weights, g1, 256, 0, st, M, W, nn, 1.0 / (sigma * sigma));
  MEM_LAUNCH(ctx, k_colsum, (nS + 31) / 32, 1024, 0, st, W, nS, nS, d, 0);
  MEM_LAUNCH(ctx, k_lap_scale, dim3((nS + 255) / 256, nS), 256, 0, st, W, nS, d);
  MEM_LAUNCH(ctx, k_colsum, (nS + 31) / 32, 1024, 0, st, W, nS, nS, d, 1);
  MEM_LAUNCH(ctx, k_lap_scale, dim3((nS + 255) / 256, nS), 256, 0, st, W, nS, d);
  MEM_LAUNCH(ctx, k_lap_sym, dim3((nS + 31) / 32, (nS + 31) / 32), 256, 0, st, W, L, nS);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a19 helper: y = L x for the dense symmetric L that stays on the device.  ARPACK (scipy eigsh,
// sembeddingonFly.py:27) keeps running on the host exactly as in the reference, but its O(nS^2) operator
// application happens here, one warp per row in a fixed summation order (deterministic).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_symv(const double* __restrict__ L, const double* __restrict__ x,
                                              double* __restrict__ y, int nS) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nS) return;
  const double* r = L + (size_t)row * nS;
  double acc = 0.0;
  for (int j = lane; j < nS; j += 32) acc = fma(r[j], x[j], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

int symv_host(mem_ctx* ctx, const double* L, int nS, const double* x_host, double* y_host) {
  cudaStream_t st = ctx->stream;
  MEM_CHECK(ctx->small_out.ensure((size_t)2 * nS * sizeof(double)));
  double* dx = ctx->small_out.as<double>();
  double* dy = dx + nS;
  MEM_CUDA(cudaMemcpyAsync(dx, x_host, (size_t)nS * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_LAUNCH(ctx, k_symv, (nS + 7) / 8, 256, 0, st, L, dx, dy, nS);
  MEM_CUDA(cudaMemcpyAsync(y_host, dy, (size_t)nS * sizeof(double), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// manifoldTrimmingAuto.py:50,63 — D[posPath][:, posPath] on a D that stays on the device: out[a][b] = D[sel[a]][sel[b]].
// The trimming loop re-embeds ever smaller subsets of one PD; only the index list travels, not the matrix.
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_gather_square(const T* __restrict__ D, int nS, const int* __restrict__ sel, int m,
                                                       T* __restrict__ out) {
  const int a = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= m) return;
  out[(size_t)a * m + b] = D[(size_t)sel[a] * nS + sel[b]];
}
int gather_square_device(mem_ctx* ctx, const void* D, int elem_bytes, int nS, const int* sel_host, int m, void* out,
                         cudaStream_t st) {
  if (m < 1 || nS < 1 || (elem_bytes != 4 && elem_bytes != 8)) {
    set_error("gather: need m >= 1, nS >= 1 and 4- or 8-byte elements (m=%d nS=%d bytes=%d)", m, nS, elem_bytes);
    return 1;
  }
  for (int a = 0; a < m; ++a)
    if (sel_host[a] < 0 || sel_host[a] >= nS) {
      set_error("gather: index %d out of range at position %d (nS=%d)", sel_host[a], a, nS);
      return 1;
    }
  MEM_CHECK(ctx->small_out.ensure((size_t)m * sizeof(int)));
  int* d_sel = ctx->small_out.as<int>();
  MEM_CUDA(cudaMemcpyAsync(d_sel, sel_host, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));                       // sel_host may be a temporary of the caller
  const dim3 grid((m + 255) / 256, m);
  if (elem_bytes == 4) MEM_LAUNCH(ctx, k_gather_square<float>, grid, 256, 0, st, (const float*)D, nS, d_sel, m, (float*)out);
  else MEM_LAUNCH(ctx, k_gather_square<double>, grid, 256, 0, st, (const double*)D, nS, d_sel, m, (double*)out);
  return 0;
}

}  // namespace mem
