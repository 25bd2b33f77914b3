// common.cuh — context, workspace arena and error plumbing shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <vector>
#include <string>

#include "../../include/manifoldem_b200.h"

namespace mem {

void set_error(const char* fmt, ...);

#define MEM_CUDA(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      mem::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define MEM_CUFFT(x)                                                                         \
  do {                                                                                       \
    cufftResult r_ = (x);                                                                    \
    if (r_ != CUFFT_SUCCESS) {                                                               \
      mem::set_error("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #x, (int)r_);         \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define MEM_CHECK(x)            \
  do {                          \
    if ((x) != 0) return 1;     \
  } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define MEM_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                              \
  do {                                                                                       \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                                      \
    (ctx)->launches++;                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess) {                                                                 \
      mem::set_error("%s:%d launch %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(e_)); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes);
  void release();
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Per-box-size tables (device), built once per context and N / filter.
struct Geometry {
  int N = 0, Nh = 0;          // box, half-spectrum width N/2+1
  int Kh = 0;                 // N*Nh half-spectrum pixels
  int Kr = 0;                 // distinct |k|^2 values (radial bins)
  int n_special = 0;          // purely-real self-conjugate pixels (4 for even N, 1 for odd N)
  int n1_blocks = 0;          // 32-column blocks in S1 (and S2): ceil((Kr+n_special)/32)
  int K3 = 0;                 // complex entries in S3
  int n3_blocks = 0;          // ceil(2*K3/32)
  int64_t ldz = 0;            // Z row pitch (floats) = 32*(2*n1_blocks+n3_blocks)
  int filter_type = -1, filter_order = 0;
  double filter_Qc = 0;
  DevBuf Gtab;                // [Kh] float  low-pass * 1/N^2
  DevBuf bin_of_pix;          // [Kh] int32
  DevBuf r2_of_bin;           // [Kr] int32
  DevBuf bin_start;           // [Kr+1] int32
  DevBuf bin_pix;             // [Kh] int32 pixels sorted by bin
  DevBuf s3_col;              // [Kh] int32: column (in floats, relative to Z row) of the re part, or -1
  DevBuf special_pix;         // [4] int32
  // rows ky and N-ky share |k|^2: the folded quadrant [Na][Nh], Na = N/2+1, is what the radial sums gather from
  int Na = 0;
  DevBuf fold_bin;            // [Na*Nh] int32 bin of the folded entry
  DevBuf fold_start;          // [Kr+1] int32
  DevBuf fold_ent;            // [Na*Nh] int32 folded entries sorted by bin
};

struct FftPlan { cufftHandle r2c = 0, c2r = 0; };

}  // namespace mem

struct mem_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  int sm_count = 148;
  mem::Geometry geom;
  std::map<long long, mem::FftPlan> plans;   // key = N * 2^20 + batch
  std::map<long long, cufftHandle> plans_d;  // float64 2-D plans of the NLSA stage: key = (type, N, batch)
  mem::DevBuf fft_work;
  // workspace for one PD
  mem::DevBuf raw, flip, shift, psi, df, msk2, rot_cs, rot_pid, rot_pitch_tab, batch_aux, mspec_all;
  mem::DevBuf imgA, imgB, imgAll, imgFlip, spec, spec2, cbin, zhi, zlo;
  mem::DevBuf part_cf, part_cfw, part_c2, part_fl, part_int, avgspec, avgimg, stats;
  mem::DevBuf D, ctf64, small_out;
  mem::DevBuf contract_ws;   // split-K partial tiles
  mem::DevBuf contract_items;   // work-item table of the last contraction shape
  mem::DevBuf clk_probe;        // {clock64, globaltimer} at the start and end of CTA 0 of the last k_contract_tc2
  long long items_key[4] = {-1, -1, -1, -1};   // nS, nkb, split, count
  mem::DevBuf scratch;       // misc (ferguson partials, knn)
  mem::DevBuf knn_ws;        // (key, index) rows of the chunked sort, only for nS > 16,384
  mem::DevBuf knn_out;       // kNN lists of the host-buffer entry point
  cudaEvent_t ev[10] = {};
  cudaEvent_t timer[2] = {};
  cudaEvent_t done = nullptr;     // blocking-sync event: host-buffer calls sleep instead of spinning while the GPU works
  std::vector<cudaEvent_t> kev;   // event pairs around every contraction launch since the last reset
  size_t kev_used = 0;
  int last_tc_items = 0, last_tc_nkb = 0;   // geometry of the last tcgen05 launch (executed-flop accounting)
  float timings[8] = {};
  void* tmap_encode = nullptr;   // cuTensorMapEncodeTiled entry point
  int pool_ready = 0;            // the device's default memory pool has its release threshold set
  int full_sums = 0;             // 1: the spectrum sums always run over every image (tests: M from all images)
  int cufft_rows320 = 0;         // 1: the row passes at N = 320 through cuFFT's 1-D plans + generic ingest / prefilter (comparison)
  int cufft_lowpass = 0;         // 1: ingest / low-pass / a10 through the generic kernels + cuFFT even where own FFT kernels exist (tests)
  int cufft_a10 = 0;             // 1: the a10 transform through cuFFT's 2-D plan even for N = 256 (tests / comparison)
  int rowfft_blocks = 0;         // experiments: CTAs per SM the row FFT kernels are compiled for (0 = default 4)
  int radial_variant = 0;        // experiments: thread count / unroll of k_operands_radial_sm
  int legacy_rotate = 0;         // tests: 1 = the generic k_rotate for every box (mem_ctx_set_option)
};

namespace mem {
int geometry_prepare(mem_ctx* ctx, int N, int filter_type, int filter_order, double Qc);
int fft_get(mem_ctx* ctx, int N, int batch, FftPlan* out, bool rows_only = false);
// column pass of the low-pass on the row-transformed half spectra: FFT along ky, * G, inverse FFT, in place.
// Returns false when the box size has no specialised kernel (the caller keeps the 2-D cuFFT path).
bool colfilter_supported(int N);
bool colpass_supported(int N);   // own column pass only (rows through cuFFT 1-D plans)
int colpass_run(mem_ctx* ctx, float2* spec, const float* G, const float2* stats, int nS, int N, int fwd_only, cudaStream_t st);
int rows320_forward_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float2* spec, float2* stats, int nS, int transposed,
                        int plain, cudaStream_t st);
int rows320_inverse_run(mem_ctx* ctx, const float2* spec, float* out, int nS, cudaStream_t st);
int colfilter_run(mem_ctx* ctx, float2* spec, const float* G, const float2* stats, int nS, int N, cudaStream_t st);
// a2 + a3 + row R2C in one pass for the same box sizes: raw particles -> row-transformed half spectra + (mean, 1/std)
int ingest_rowfft_run(mem_ctx* ctx, const float* raw, const uint8_t* flip, float2* spec, float2* stats, int nS, int N,
                      int transposed, cudaStream_t st);
// a10 forward transform for the same box sizes with the library's own FFT kernels (rows R2C, then columns)
int fft2_forward_run(mem_ctx* ctx, const float* img, float2* spec, int nS, int N, cudaStream_t st);
// inverse row pass (C2R) + annular mask + row pass of the periodic spline prefilter: half spectra -> out [nS][N][N]
int rowifft_prefilter_run(mem_ctx* ctx, const float2* spec, float* out, int nS, int N, cudaStream_t st);
// kNN lists wanted from the contraction (a15 fused behind a12): idx [nS][k] int32, val [nS][k] float64, device.
// With D == nullptr the nS x nS matrix is never assembled: the lists are selected from the split-K partial tiles.
struct KnnOut { int k = 0; int* idx = nullptr; double* val = nullptr; };
int contract_run(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                 int contraction, int k_chunk_blocks, int split_k, cudaStream_t st, const KnnOut* knn = nullptr);
int contract_tc(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                int k_chunk_blocks, int split_k, cudaStream_t st, int two_cta, const KnnOut* knn = nullptr);
int knn_device_f32(mem_ctx* ctx, const float* D, int nS, int k, int* idx, double* val, cudaStream_t st);
int knn_from_workspace(mem_ctx* ctx, const float* ws, int ldw, int nslices, int nS, int k, int* idx, double* val,
                       cudaStream_t st);
void knn_set_mode(int mode);
// rotate.cu: the two rotations of a7 for boxes that are a multiple of 32 (other boxes: k_rotate in align.cu)
bool rotate_fast_supported(int N);
int rotate_angles_run(mem_ctx* ctx, const double* psi_deg, double psi_p_deg, double2* cs, uint8_t* pid, int nS,
                      cudaStream_t st);
int rotate_angles_batch_run(mem_ctx* ctx, const int* pd_of, const double* psi_p_deg, double2* cs2, uint8_t* pid2, int nS,
                            cudaStream_t st);
int contract_tc_grouped(mem_ctx* ctx, const mem_contract_shape* shp, int n_pd, const int* pd_start, const float* Zhi,
                        const float* Zlo, float* const* D, cudaStream_t st);
int rotate_img_run(mem_ctx* ctx, const float* coef, float* out, const double2* cs, const uint8_t* pid, int nS, int N,
                   cudaStream_t st);
int rotate_common_run(mem_ctx* ctx, const float* coef, float* out, const double2* cs_common, double angle_deg, int nS,
                      int N, const uint8_t* msk2, float* out_masked, cudaStream_t st);
}  // namespace mem
