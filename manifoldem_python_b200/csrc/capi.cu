// capi.cu — extern "C" surface declared in include/manifoldem_b200.h.
#include "common.cuh"
#include <stdlib.h>

#include <math.h>
#include <algorithm>
#include <set>
#include <thread>

namespace mem {
const char* last_error();
int pd_distance_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, cudaStream_t st);
int pd_distance_batch_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, int n_pd, const int* pd_start,
                             const double* psi_p_deg, cudaStream_t st);
int knn_device(mem_ctx* ctx, const double* D, int nS, int k, int* idx, double* val, cudaStream_t st);
int graph_compact_device(mem_ctx* ctx, const double* M, int nS, double* out, long long* count);
int graph_dense_device(mem_ctx* ctx, const int* idx, const double* val, int nS, int k, double* M, cudaStream_t st);
int ferguson_device(mem_ctx* ctx, const double* d2, int64_t n, const double* logEps, int nEps, double thr, double* out);
int laplacian_dense_device(mem_ctx* ctx, const double* M, int nS, double sigma, double* L, cudaStream_t st);
int symv_host(mem_ctx* ctx, const double* L, int nS, const double* x_host, double* y_host);
int s2_assign_host(mem_ctx* ctx, const double* centres, int nG, const double* pts, long long n, int* idx);
int lanczos_steps_device(mem_ctx* ctx, const double* L, int nS, double* V, double* ab, int ld_ab, int j0, int j1,
                         cudaStream_t st);
int lanczos_ritz_device(mem_ctx* ctx, const double* V, int nS, int j, const double* S_host, int k, double* X,
                        cudaStream_t st);
int nlsa_spectra_device(mem_ctx* ctx, const double* img, const double* ctf, int n, int N, double2* H, double* Ch,
                        cudaStream_t st);
int nlsa_cond_device(mem_ctx* ctx, const void* D, int elem_bytes, int nAll, const int* sel, int num, int ConOrder,
                     double* out, cudaStream_t st);
int nlsa_supervectors_device(mem_ctx* ctx, const double2* H, const double* Ch, const int* sel, const double* mu_psi_host,
                             int num, int ConOrder, int E, int N, const double* msk2, double* A, cudaStream_t st);
int nlsa_gram_small_device(mem_ctx* ctx, const double* A, long long rows, int E, double* AtA_host, cudaStream_t st);
int nlsa_project_device(mem_ctx* ctx, const double* A, long long rows, int E, const double* M_host, double* U, int Npix,
                        int ConOrder, double* topo_host, cudaStream_t st);
int nlsa_reconstruct_device(mem_ctx* ctx, const double* U, int Npix, int ConOrder, int E, const double* Q_host, int nI, int nC,
                            double* IMGT, double* D2, cudaStream_t st);
int manifold_fit_host(mem_ctx* ctx, const double* x_host, int nS, double* ab_host, double* tau_host, int max_iter,
                      double da_max, double db_max, int* iters_host, cudaStream_t st);
int s2_pairwise_host(mem_ctx* ctx, const double* U, int nU, const double* V, int nV, double* dot, double* dist);
int ctf_host(mem_ctx* ctx, const mem_pd_params* prm, const double* df, double* out);
int gather_square_device(mem_ctx* ctx, const void* D, int elem_bytes, int nS, const int* sel_host, int m, void* out,
                         cudaStream_t st);
}  // namespace mem

using namespace mem;

static cudaStream_t pick(mem_ctx* ctx, void* stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

extern "C" {

int mem_version(void) { return 100; }
const char* mem_last_error(void) { return mem::last_error(); }

int mem_ctx_create(int device, mem_ctx** out) {
  int n = 0;
  MEM_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) {
    set_error("no CUDA device %d (found %d)", device, n);
    return 1;
  }
  MEM_CUDA(cudaSetDevice(device));
  mem_ctx* ctx = new mem_ctx();
  ctx->device = device;
  MEM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  MEM_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  for (auto& e : ctx->ev) MEM_CUDA(cudaEventCreate(&e));
  for (auto& e : ctx->timer) MEM_CUDA(cudaEventCreate(&e));
  MEM_CUDA(cudaEventCreateWithFlags(&ctx->done, cudaEventBlockingSync | cudaEventDisableTiming));
  *out = ctx;
  return 0;
}

int mem_ctx_destroy(mem_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->plans) {
    cufftDestroy(kv.second.r2c);
    cufftDestroy(kv.second.c2r);
  }
  for (auto& kv : ctx->plans_d) cufftDestroy(kv.second);
  mem::DevBuf* bufs[] = {&ctx->fft_work, &ctx->raw, &ctx->flip, &ctx->shift, &ctx->psi, &ctx->df, &ctx->msk2, &ctx->rot_cs, &ctx->rot_pid, &ctx->rot_pitch_tab, &ctx->batch_aux,
                         &ctx->imgA, &ctx->imgB, &ctx->imgAll, &ctx->imgFlip, &ctx->spec, &ctx->spec2, &ctx->cbin, &ctx->zhi,
                         &ctx->zlo, &ctx->part_cf, &ctx->part_cfw, &ctx->part_c2, &ctx->part_fl, &ctx->part_int, &ctx->avgspec, &ctx->avgimg,
                         &ctx->stats, &ctx->D, &ctx->ctf64, &ctx->small_out, &ctx->contract_ws, &ctx->contract_items, &ctx->clk_probe, &ctx->scratch,
                         &ctx->geom.Gtab, &ctx->geom.bin_of_pix, &ctx->geom.r2_of_bin, &ctx->geom.bin_start,
                         &ctx->geom.bin_pix, &ctx->geom.s3_col, &ctx->geom.special_pix, &ctx->geom.fold_bin,
                         &ctx->geom.fold_start, &ctx->geom.fold_ent, &ctx->knn_ws, &ctx->knn_out};
  for (auto* b : bufs) b->release();
  for (auto& e : ctx->ev) cudaEventDestroy(e);
  for (auto& e : ctx->timer) cudaEventDestroy(e);
  if (ctx->done) cudaEventDestroy(ctx->done);
  for (auto& e : ctx->kev) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

int mem_ctx_sync(mem_ctx* ctx) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  MEM_CUDA(cudaDeviceSynchronize());
  return 0;
}

int mem_ctx_set_option(mem_ctx* ctx, const char* name, int32_t value) {
  if (!ctx || !name) {
    set_error("mem_ctx_set_option: null argument");
    return 1;
  }
  if (!strcmp(name, "legacy_rotate")) ctx->legacy_rotate = value;
  else if (!strcmp(name, "radial_variant")) ctx->radial_variant = value;
  else if (!strcmp(name, "full_sums")) ctx->full_sums = value;
  else if (!strcmp(name, "rowfft_blocks")) ctx->rowfft_blocks = value;
  else if (!strcmp(name, "cufft_a10")) ctx->cufft_a10 = value;
  else if (!strcmp(name, "cufft_lowpass")) ctx->cufft_lowpass = value;
  else if (!strcmp(name, "cufft_rows320")) ctx->cufft_rows320 = value;
  else {
    set_error("mem_ctx_set_option: unknown option '%s'", name);
    return 1;
  }
  return 0;
}

int64_t mem_ctx_launch_count(mem_ctx* ctx, int reset) {
  const int64_t v = ctx->launches;
  if (reset) ctx->launches = 0;
  return v;
}

int mem_ctx_timer_start(mem_ctx* ctx) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaEventRecord(ctx->timer[0], ctx->stream));
  return 0;
}
int mem_ctx_timer_stop(mem_ctx* ctx, float* ms) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaEventRecord(ctx->timer[1], ctx->stream));
  MEM_CUDA(cudaEventSynchronize(ctx->timer[1]));
  MEM_CUDA(cudaEventElapsedTime(ms, ctx->timer[0], ctx->timer[1]));
  return 0;
}
int mem_ctx_kernel_time(mem_ctx* ctx, int reset, double* total_ms, int64_t* launches, int32_t* items, int32_t* k_blocks) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  double tot = 0;
  for (size_t i = 0; i + 1 < ctx->kev_used; i += 2) {
    float ms = 0;
    MEM_CUDA(cudaEventElapsedTime(&ms, ctx->kev[i], ctx->kev[i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)(ctx->kev_used / 2);
  if (items) *items = ctx->last_tc_items;
  if (k_blocks) *k_blocks = ctx->last_tc_nkb;
  if (reset) ctx->kev_used = 0;
  return 0;
}

int mem_ctx_kernel_clock(mem_ctx* ctx, double* sm_mhz, double* kernel_ms) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!ctx->clk_probe.p) {
    set_error("mem_ctx_kernel_clock: no tcgen05 contraction has run on this context");
    return 1;
  }
  unsigned long long v[4];
  MEM_CUDA(cudaMemcpy(v, ctx->clk_probe.p, sizeof(v), cudaMemcpyDeviceToHost));
  const double ns = (double)(v[3] - v[1]), cyc = (double)(v[2] - v[0]);
  if (sm_mhz) *sm_mhz = ns > 0 ? cyc / ns * 1e3 : 0.0;
  if (kernel_ms) *kernel_ms = ns * 1e-6;
  return 0;
}

int mem_gather_rows_host(void* dst, const void* src, const int64_t* rows, int64_t n, size_t row_bytes, int32_t threads) {
  if (!dst || !src || !rows || n < 0) {
    set_error("mem_gather_rows_host: null argument");
    return 1;
  }
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(threads > 0 ? threads : 4, 64), n));
  auto work = [=](int t) {
    const int64_t k0 = n * t / T, k1 = n * (t + 1) / T;
    for (int64_t k = k0; k < k1; ++k)
      memcpy((char*)dst + (size_t)k * row_bytes, (const char*)src + (size_t)rows[k] * row_bytes, row_bytes);
  };
  if (T == 1) {
    work(0);
    return 0;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  return 0;
}

int mem_host_alloc(void** out, size_t bytes) {
  MEM_CUDA(cudaMallocHost(out, bytes));
  return 0;
}
int mem_host_free(void* p) {
  MEM_CUDA(cudaFreeHost(p));
  return 0;
}
// Device arrays of the Python layer come from CUDA's stream-ordered allocator on the context's stream (the default memory
// pool keeps what is freed: a cudaMalloc / cudaFree pair per temporary — each free synchronising the device — was 40 % of a
// 10 ms DMembeddingII.embed call at nS = 2,000).  MANIFOLDEM_B200_ALLOC=sync selects plain cudaMalloc / cudaFree.
static bool alloc_async() {
  static const bool v = [] {
    const char* e = getenv("MANIFOLDEM_B200_ALLOC");
    return !(e && !strcmp(e, "sync"));
  }();
  return v;
}
int mem_dev_alloc(mem_ctx* ctx, void** out, size_t bytes) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  if (alloc_async()) {
    if (!ctx->pool_ready) {
      cudaMemPool_t pool;
      MEM_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
      uint64_t keep = 8ull << 30;                         // hold up to 8 GB of freed blocks for reuse
      MEM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      ctx->pool_ready = 1;
    }
    MEM_CUDA(cudaMallocAsync(out, bytes ? bytes : 1, ctx->stream));
    return 0;
  }
  MEM_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
}
int mem_dev_free(mem_ctx* ctx, void* p) {
  if (ctx) MEM_CUDA(cudaSetDevice(ctx->device));
  if (ctx && alloc_async()) {
    MEM_CUDA(cudaFreeAsync(p, ctx->stream));              // ordered behind the work already enqueued on the context's stream
    return 0;
  }
  MEM_CUDA(cudaFree(p));                                  // also frees stream-ordered allocations (synchronises)
  return 0;
}
int mem_copy_h2d(mem_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int mem_copy_d2h(mem_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int mem_pd_distance_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return pd_distance_device(ctx, prm, io, pick(ctx, stream));
}

int mem_pd_distance_batch_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, int32_t n_pd,
                                 const int32_t* pd_start, const double* psi_p_deg, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return pd_distance_batch_device(ctx, prm, io, n_pd, pd_start, psi_p_deg, pick(ctx, stream));
}

int mem_pd_distance_host(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* h) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t nS = prm->nS, NN = (size_t)prm->N * prm->N;
  mem_pd_io d = {};
  MEM_CHECK(ctx->raw.ensure(nS * NN * sizeof(float)));
  MEM_CHECK(ctx->flip.ensure(nS));
  MEM_CHECK(ctx->psi.ensure(nS * sizeof(double)));
  MEM_CHECK(ctx->df.ensure(nS * sizeof(double)));
  MEM_CUDA(cudaEventRecord(ctx->ev[6], st));
  MEM_CUDA(cudaMemcpyAsync(ctx->raw.p, h->raw, nS * NN * sizeof(float), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(ctx->flip.p, h->flip, nS, cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(ctx->psi.p, h->psi_deg, nS * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(ctx->df.p, h->df, nS * sizeof(double), cudaMemcpyHostToDevice, st));
  d.raw = ctx->raw.as<float>();
  d.flip = ctx->flip.as<uint8_t>();
  d.psi_deg = ctx->psi.as<double>();
  d.df = ctx->df.as<double>();
  if (h->shift) {
    MEM_CHECK(ctx->shift.ensure(nS * 2 * sizeof(double)));
    MEM_CUDA(cudaMemcpyAsync(ctx->shift.p, h->shift, nS * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    d.shift = ctx->shift.as<double>();
  }
  if (h->msk2) {
    MEM_CHECK(ctx->msk2.ensure(NN));
    MEM_CUDA(cudaMemcpyAsync(ctx->msk2.p, h->msk2, NN, cudaMemcpyHostToDevice, st));
    d.msk2 = ctx->msk2.as<uint8_t>();
  }
  MEM_CUDA(cudaEventRecord(ctx->ev[7], st));
  if (h->D && !prm->avg_only) {
    MEM_CHECK(ctx->D.ensure(nS * nS * sizeof(float)));
    d.D = ctx->D.as<float>();
  }
  if (h->imgAll) {
    MEM_CHECK(ctx->imgAll.ensure(nS * NN * sizeof(float)));
    d.imgAll = ctx->imgAll.as<float>();
  }
  if (h->imgAllFlip) {
    MEM_CHECK(ctx->imgFlip.ensure(nS * NN * sizeof(float)));
    d.imgAllFlip = ctx->imgFlip.as<float>();
  }
  if (h->CTF) {
    MEM_CHECK(ctx->ctf64.ensure(nS * NN * sizeof(double)));
    d.CTF = ctx->ctf64.as<double>();
  }
  const bool want_knn = prm->knn_k > 0 && h->knn_idx && h->knn_val && !prm->avg_only;
  if (want_knn) {   // lists: int32 indices behind the float64 values in one buffer
    MEM_CHECK(ctx->knn_out.ensure(nS * (size_t)prm->knn_k * (sizeof(double) + sizeof(int32_t))));
    d.knn_val = ctx->knn_out.as<double>();
    d.knn_idx = reinterpret_cast<int32_t*>(d.knn_val + nS * (size_t)prm->knn_k);
  }
  MEM_CHECK(ctx->small_out.ensure(3 * NN * sizeof(float)));   // own buffer: `stats` is re-sized by the pipeline itself
  float* small = ctx->small_out.as<float>();
  if (h->imgAvg) d.imgAvg = small;
  if (h->imgAvgFlip) d.imgAvgFlip = small + NN;
  if (h->imgAllIntensity) d.imgAllIntensity = small + 2 * NN;
  MEM_CHECK(pd_distance_device(ctx, prm, &d, st));
  MEM_CUDA(cudaEventRecord(ctx->ev[8], st));
  if (d.D) MEM_CUDA(cudaMemcpyAsync(h->D, d.D, nS * nS * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (want_knn) {
    MEM_CUDA(cudaMemcpyAsync(h->knn_idx, d.knn_idx, nS * (size_t)prm->knn_k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MEM_CUDA(cudaMemcpyAsync(h->knn_val, d.knn_val, nS * (size_t)prm->knn_k * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  if (d.imgAll) MEM_CUDA(cudaMemcpyAsync(h->imgAll, d.imgAll, nS * NN * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (d.imgAllFlip) MEM_CUDA(cudaMemcpyAsync(h->imgAllFlip, d.imgAllFlip, nS * NN * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (d.CTF) MEM_CUDA(cudaMemcpyAsync(h->CTF, d.CTF, nS * NN * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (d.imgAvg) MEM_CUDA(cudaMemcpyAsync(h->imgAvg, d.imgAvg, NN * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (d.imgAvgFlip) MEM_CUDA(cudaMemcpyAsync(h->imgAvgFlip, d.imgAvgFlip, NN * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (d.imgAllIntensity) MEM_CUDA(cudaMemcpyAsync(h->imgAllIntensity, d.imgAllIntensity, NN * sizeof(float), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaEventRecord(ctx->ev[9], st));
  // several PDs are in flight per GPU, one host thread each (and one process per GPU): wait without spinning
  MEM_CUDA(cudaEventRecord(ctx->done, st));
  MEM_CUDA(cudaEventSynchronize(ctx->done));
  return 0;
}

int mem_ctf_host(mem_ctx* ctx, const mem_pd_params* prm, const double* df, double* CTF) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return ctf_host(ctx, prm, df, CTF);
}

int mem_pd_last_timings(mem_ctx* ctx, float* ms, int n) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  MEM_CUDA(cudaStreamSynchronize(ctx->stream));
  float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], ctx->ev[i], ctx->ev[i + 1]);
  cudaEventElapsedTime(&t[5], ctx->ev[0], ctx->ev[5]);
  if (cudaEventElapsedTime(&t[6], ctx->ev[6], ctx->ev[7]) != cudaSuccess) t[6] = 0;
  if (cudaEventElapsedTime(&t[7], ctx->ev[8], ctx->ev[9]) != cudaSuccess) t[7] = 0;
  cudaGetLastError();
  for (int i = 0; i < n && i < 8; ++i) ms[i] = t[i];
  return 0;
}

int mem_contract_device(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                        int32_t contraction, int32_t k_chunk_blocks, int32_t split_k, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return contract_run(ctx, shp, Zhi, Zlo, D, contraction, k_chunk_blocks, split_k, pick(ctx, stream));
}

int mem_contract_knn_device(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                            int32_t k, int32_t* knn_idx, double* knn_val, int32_t contraction, int32_t k_chunk_blocks,
                            int32_t split_k, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  KnnOut knn;
  knn.k = k; knn.idx = knn_idx; knn.val = knn_val;
  if (k <= 0 || !knn_idx || !knn_val) {
    set_error("mem_contract_knn_device: k > 0 and both list pointers are required");
    return 1;
  }
  return contract_run(ctx, shp, Zhi, Zlo, D, contraction, k_chunk_blocks, split_k, pick(ctx, stream), &knn);
}

int mem_gather_square_device(mem_ctx* ctx, const void* D, int32_t elem_bytes, int32_t nS, const int32_t* sel, int32_t m,
                             void* out, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return gather_square_device(ctx, D, elem_bytes, nS, sel, m, out, pick(ctx, stream));
}

int mem_knn_mode(int32_t mode) {
  if (mode < 0 || mode > 2) {
    set_error("mem_knn_mode: 0 (auto), 1 (sort) or 2 (selection)");
    return 1;
  }
  knn_set_mode(mode);
  return 0;
}

int mem_operand_shape(mem_ctx* ctx, int32_t N, mem_contract_shape* out) {
  (void)ctx;
  if (N < 4) {
    set_error("bad box size %d", N);
    return 1;
  }
  const int Nh = N / 2 + 1;
  std::set<int> r2;
  for (int ky = 0; ky < N; ++ky) {
    const int fy = ky < (N + 1) / 2 ? ky : ky - N;
    for (int kx = 0; kx < Nh; ++kx) r2.insert(fy * fy + kx * kx);
  }
  const bool even = (N % 2 == 0);
  const int n_special = even ? 4 : 1;
  const int two_k3 = N * N - n_special;
  out->nS = 0;
  out->n1_blocks = ((int)r2.size() + n_special + 31) / 32;
  out->n3_blocks = (two_k3 + 31) / 32;
  out->ldz = 32LL * (2 * out->n1_blocks + out->n3_blocks);
  return 0;
}

int mem_knn_device(mem_ctx* ctx, const double* D, int32_t nS, int32_t k, int32_t* idx, double* val, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return knn_device(ctx, D, nS, k, idx, val, pick(ctx, stream));
}
int mem_knn_device_f32(mem_ctx* ctx, const float* D, int32_t nS, int32_t k, int32_t* idx, double* val, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return knn_device_f32(ctx, D, nS, k, idx, val, pick(ctx, stream));
}
int mem_graph_compact_device(mem_ctx* ctx, const double* M, int32_t nS, double* out, int64_t* count) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  long long c = 0;
  MEM_CHECK(graph_compact_device(ctx, M, nS, out, &c));
  *count = c;
  return 0;
}
int mem_graph_dense_device(mem_ctx* ctx, const int32_t* idx, const double* val, int32_t nS, int32_t k, double* M,
                           void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return graph_dense_device(ctx, idx, val, nS, k, M, pick(ctx, stream));
}
int mem_ferguson_device(mem_ctx* ctx, const double* d2, int64_t n, const double* logEps, int32_t nEps, double thr,
                        double* out) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return ferguson_device(ctx, d2, n, logEps, nEps, thr, out);
}
int mem_laplacian_dense_device(mem_ctx* ctx, const double* M, int32_t nS, double sigma, double* L, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return laplacian_dense_device(ctx, M, nS, sigma, L, pick(ctx, stream));
}
int mem_symv_host(mem_ctx* ctx, const double* L, int32_t nS, const double* x, double* y) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return symv_host(ctx, L, nS, x, y);
}
int mem_s2_assign_host(mem_ctx* ctx, const double* centres, int32_t nG, const double* pts, int64_t n, int32_t* idx) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return s2_assign_host(ctx, centres, nG, pts, (long long)n, idx);
}

int mem_lanczos_steps_device(mem_ctx* ctx, const double* L, int32_t nS, double* V, double* ab, int32_t ld_ab, int32_t j0,
                             int32_t j1, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return lanczos_steps_device(ctx, L, nS, V, ab, ld_ab, j0, j1, pick(ctx, stream));
}

int mem_lanczos_ritz_device(mem_ctx* ctx, const double* V, int32_t nS, int32_t j, const double* S_host, int32_t k, double* X,
                            void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return lanczos_ritz_device(ctx, V, nS, j, S_host, k, X, pick(ctx, stream));
}

int mem_nlsa_spectra_device(mem_ctx* ctx, const double* img, const double* ctf, int32_t n, int32_t N, void* H, double* Ch,
                            void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_spectra_device(ctx, img, ctf, n, N, reinterpret_cast<double2*>(H), Ch, pick(ctx, stream));
}
int mem_nlsa_cond_device(mem_ctx* ctx, const void* D, int32_t elem_bytes, int32_t nAll, const int32_t* sel, int32_t num,
                         int32_t ConOrder, double* ConD, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_cond_device(ctx, D, elem_bytes, nAll, sel, num, ConOrder, ConD, pick(ctx, stream));
}
int mem_nlsa_supervectors_device(mem_ctx* ctx, const void* H, const double* Ch, const int32_t* sel, const double* mu_psi,
                                 int32_t num, int32_t ConOrder, int32_t E, int32_t N, const double* msk2, double* A,
                                 void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_supervectors_device(ctx, reinterpret_cast<const double2*>(H), Ch, sel, mu_psi, num, ConOrder, E, N, msk2, A,
                                  pick(ctx, stream));
}
int mem_nlsa_gram_small_device(mem_ctx* ctx, const double* A, int64_t rows, int32_t E, double* AtA, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_gram_small_device(ctx, A, (long long)rows, E, AtA, pick(ctx, stream));
}
int mem_nlsa_project_device(mem_ctx* ctx, const double* A, int64_t rows, int32_t E, const double* M, double* U, int32_t Npix,
                            int32_t ConOrder, double* topo_mean, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_project_device(ctx, A, (long long)rows, E, M, U, Npix, ConOrder, topo_mean, pick(ctx, stream));
}
int mem_nlsa_reconstruct_device(mem_ctx* ctx, const double* U, int32_t Npix, int32_t ConOrder, int32_t E, const double* Q,
                                int32_t nI, int32_t nC, double* IMGT, double* D2, void* stream) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return nlsa_reconstruct_device(ctx, U, Npix, ConOrder, E, Q, nI, nC, IMGT, D2, pick(ctx, stream));
}

int mem_manifold_fit_host(mem_ctx* ctx, const double* x, int32_t nS, double* ab, double* tau, int32_t max_iter, double da_max,
                          double db_max, int32_t* iters) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return manifold_fit_host(ctx, x, nS, ab, tau, max_iter, da_max, db_max, iters, ctx->stream);
}

int mem_s2_pairwise_host(mem_ctx* ctx, const double* U, int32_t nU, const double* V, int32_t nV, double* dot, double* dist) {
  MEM_CUDA(cudaSetDevice(ctx->device));
  return s2_pairwise_host(ctx, U, nU, V, nV, dot, dist);
}

}  // extern "C"
