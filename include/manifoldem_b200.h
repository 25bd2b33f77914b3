/* manifoldem_b200.h — C ABI of the B200-native ManifoldEM hot path.
 *
 * The reference (evanseitz/ManifoldEM_Python) has no FFI layer: its boundary
 * for this path is three Python callables.  This header is the
 * native surface those callables bind through ctypes; every entry point cites
 * the reference code it replaces.  Plain pointers and sizes only — no torch
 * types.  All functions return 0 on success, non-zero on error (message from
 * mem_last_error()).  "device" pointers are CUDA device pointers on the
 * context's device; "host" pointers are ordinary (ideally pinned) host memory.
 *
 * Image layout: [nS][N][N] float32, row-major, picture orientation
 * (row = first NumPy axis).  Spectra: cuFFT R2C layout [nS][N][N/2+1] complex64.
 */
#ifndef MANIFOLDEM_B200_H
#define MANIFOLDEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mem_ctx mem_ctx;

/* ---- lifecycle ------------------------------------------------------------------ */
int         mem_version(void);
const char* mem_last_error(void);
int         mem_ctx_create(int device, mem_ctx** out);
int         mem_ctx_destroy(mem_ctx* ctx);
int         mem_ctx_sync(mem_ctx* ctx);
/* tuning / test switches of a context, by name; unknown names are an error.
 *   "legacy_rotate"  1 = the generic rotation kernel for every box size (default 0: boxes that are a multiple of 32
 *                    take the float4-staged / four-images-per-tap kernels of rotate.cu; results are bit-identical)
 *   "full_sums"      1 = the per-pixel spectrum sums always run over every image.  Default 0: when only D / the neighbour
 *                    lists are requested the sums serve nothing but the common component M removed from the contraction
 *                    operands (D is invariant under any common M), and PDs of >= 1,024 images take every 8th image
 *   "cufft_a10"      1 = the a10 forward transform through cuFFT's 2-D plan also at N = 128 / 256 (default 0: the library's own
 *                    row / column FFT kernels; other boxes always take cuFFT)
 *   "cufft_lowpass"  1 = ingest, low-pass and a10 through the generic kernels + cuFFT's 2-D plans also at N = 128 / 256 / 320
 *                    (comparison / tests; default 0: own row and column kernels at 128 / 256 / 320)
 *   "cufft_rows320"  1 = at N = 320 only the column pass is the library's own, rows through cuFFT's batched 1-D plans (comparison)
 *   "radial_variant", "rowfft_blocks"   experiment switches (thread count of the operand writer, CTAs per SM the row FFT kernels
 *                    are compiled for); 0 = the measured defaults */
int         mem_ctx_set_option(mem_ctx* ctx, const char* name, int32_t value);
/* kernels launched by this library on ctx since the last reset (bench.py gpu_launches) */
int64_t     mem_ctx_launch_count(mem_ctx* ctx, int reset);
/* CUDA-event stopwatch on the context's stream (bench.py times its steps on the launching stream) */
int         mem_ctx_timer_start(mem_ctx* ctx);
int         mem_ctx_timer_stop(mem_ctx* ctx, float* ms);
/* summed CUDA-event duration of every tcgen05 contraction launch since the last reset, their count, and
 * the CTA count (each CTA = one 128x256 tile of one K slice) and K blocks per CTA of the last one
 * (executed-flop accounting for the roofline) */
int         mem_ctx_kernel_time(mem_ctx* ctx, int reset, double* total_ms, int64_t* launches, int32_t* items,
                                int32_t* k_blocks);
/* SM clock (MHz) the LAST tcgen05 contraction launch on ctx actually ran at, from a clock64 / globaltimer probe inside
 * the kernel (CTA 0), and the probe's own wall time of that CTA in ms.  Synchronises. */
int         mem_ctx_kernel_clock(mem_ctx* ctx, double* sm_mhz, double* kernel_ms);
/* pinned host memory for the host-buffer entry points */
int         mem_host_alloc(void** out, size_t bytes);
int         mem_host_free(void* p);
/* a2 gather (getDistanceCTF_local_Conj9combinedS2.py:246-262 reads the members of a PD one by one from the memory-mapped
 * stack): dst[k] = src[rows[k]], rows of row_bytes bytes, copied by `threads` host threads (<= 0: 4) — e.g. straight
 * from the mapped file into pinned memory.  Plain host memory on both sides; no CUDA call. */
int         mem_gather_rows_host(void* dst, const void* src, const int64_t* rows, int64_t n, size_t row_bytes,
                                 int32_t threads);
/* plain device memory on the context's device + copies (so a ctypes host needs nothing else to stage data);
 * every entry point that takes a ctx makes ctx's device current in the calling thread first */
int         mem_dev_alloc(mem_ctx* ctx, void** out, size_t bytes);
int         mem_dev_free(mem_ctx* ctx, void* p);
int         mem_copy_h2d(mem_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int         mem_copy_d2h(mem_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- per-PD distance stage -------------------------------------------------------
 * Replaces getDistanceCTF_local_Conj9combinedS2.op
 * (modules/getDistanceCTF_local_Conj9combinedS2.py:216-420) from the point where
 * the particle images of the PD have been gathered, up to (not including) the
 * pickle dump. Host-side scalars (psi angles, PD) are computed by the Python host
 * exactly as the reference does (q2Spider / least_squares). */
typedef struct {
  int32_t nS;            /* particles in this PD                                   (:223) */
  int32_t N;             /* box size p.nPix                                        (:225) */
  int32_t transposed;    /* 1: SPIDER raw layout, picture = raw^T                  (:254-258) */
  int32_t relion_shift;  /* 1: apply shift(order=3,mode='wrap') by (shy-.5,shx-.5) (:263-264) */
  int32_t filter_type;   /* 0 = 'Butter', 1 = 'Gauss'                              (:156-167) */
  int32_t filter_order;  /* filterPar['N']                                                    */
  double  filter_Qc;     /* filterPar['Qc']                                                   */
  double  pix_size;      /* p.pix_size [A]                                         (:337) */
  double  Cs;            /* p.Cs [mm]                                                        */
  double  EkV;           /* p.EkV [kV]                                                       */
  double  gaussEnv;      /* p.gaussEnv (inf => envelope 1)                                   */
  double  AmpContrast;   /* p.AmpContrast                                                    */
  double  psi_p_deg;     /* psi_ang(PD), degrees                                   (:313) */
  int32_t avg_only;      /* options['avgOnly']: skip D                             (:376) */
  int32_t contraction;   /* 0 = tcgen05 3xTF32, CTA-pair tiles (product); 1 = SIMT fp64-accumulate checker; 2 = tcgen05 single-CTA tiles */
  int32_t k_chunk_blocks;/* tcgen05: K blocks (of 32) accumulated in TMEM before promotion to FP32 registers;
                            low byte = period inside the S1/S2 columns, bits 8.. = period inside S3 (0 = same);
                            0 = default (CTA pairs: 2 | 4 << 8, single CTA: 1) */
  int32_t split_k;       /* tcgen05: K slices per tile; 0 = auto (fill 148 SMs)                  */
  int32_t knn_k;         /* > 0 with io.knn_idx / io.knn_val: the k nearest neighbours of every particle
                            (DMembeddingII.initialize, modules/DMembeddingII.py:43-57) straight behind the
                            contraction; with io.D == NULL the nS x nS matrix is never assembled */
  int32_t reserved0;
} mem_pd_params;

typedef struct {
  /* inputs */
  const float*   raw;        /* [nS][N*N] gathered particle images as stored on disk           */
  const uint8_t* flip;       /* [nS] 1 = conjugate member (np.flipud after reading)  (:271-275) */
  const double*  shift;      /* [nS][2] (shy-0.5, shx-0.5) per member, or NULL                 */
  const double*  psi_deg;    /* [nS] first in-plane rotation angle, degrees = -(180/pi)*Psi (:326) */
  const double*  df;         /* [nS] defocus [A]                                               */
  const uint8_t* msk2;       /* [N][N] projected volume mask or NULL (msk2 = 1)      (:304-310) */
  /* outputs, any may be NULL */
  float*  D;                 /* [nS][nS] squared distances                           (:391-397) */
  float*  imgAll;            /* [nS][N][N] aligned images                            (:349) */
  float*  imgAllFlip;        /* [nS][N][N] phase-flipped images                      (:346-347) */
  double* CTF;               /* [nS][N*N]  CTF, ifftshift-ed, float64                (:339) */
  float*  imgAvg;            /* [N][N] Wiener-filtered average                       (:353-366) */
  float*  imgAvgFlip;        /* [N][N] average of phase-flipped images               (:367) */
  float*  imgAllIntensity;   /* [N][N] mean(imgAllFlip^2)                            (:400) */
  int32_t* knn_idx;          /* [nS][knn_k] neighbour indices, self first   (DMembeddingII.py:43-57) */
  double*  knn_val;          /* [nS][knn_k] squared distances, [i][0] = 0                            */
} mem_pd_io;

/* all pointers in io are DEVICE pointers; work is enqueued on `stream` (a cudaStream_t, NULL = the
 * context's stream); no host synchronisation. */
int mem_pd_distance_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, void* stream);
/* A GROUP of PDs in one call (GetDistancesS2.py:94-120 hands its workers one PD at a time; the tessellation's PDs hold
 * 117..2,000 particles, too few to fill the GPU one by one).  prm->nS = images of all PDs, io->raw / flip / psi_deg / df
 * = the PDs' arrays back to back (PD g = images pd_start[g] .. pd_start[g+1]-1, pd_start [n_pd + 1] on the HOST, from 0 to
 * nS), psi_p_deg [n_pd] on the HOST (prm->psi_p_deg is ignored).  io->D receives the n_pd matrices back to back (PD g at
 * float offset sum_{h<g} nS_h^2); io->imgAll is optional; every other output, msk2 and the RELION shift are not available
 * here.  Device pointers; the per-image stages run once over the concatenated stack and one grouped tcgen05 launch
 * contracts all PDs.  Synchronises twice (two small host tables are uploaded). */
int mem_pd_distance_batch_device(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io, int32_t n_pd,
                                 const int32_t* pd_start, const double* psi_p_deg, void* stream);
/* all pointers in io are HOST pointers; stages H2D, runs, copies results back, synchronises. */
int mem_pd_distance_host(mem_ctx* ctx, const mem_pd_params* prm, const mem_pd_io* io);
/* the CTF field alone (ctemh_cryoFrank.op, modules/ctemh_cryoFrank.py:24-44, as :339 / :393 store it): uses nS, N,
 * pix_size, Cs, EkV, gaussEnv, AmpContrast of prm; df [nS] and CTF [nS][N*N] float64 are HOST pointers.  Same kernel as
 * inside mem_pd_distance_*, so a record that keeps df instead of the field reads back bit-identical values. */
int mem_ctf_host(mem_ctx* ctx, const mem_pd_params* prm, const double* df, double* CTF);
/* CUDA-event timings (ms) of the last mem_pd_distance_* call on ctx:
 * [0] ingest+lowpass  [1] align (2x prefilter+rotate)  [2] FFT+CTF+operands  [3] flip/averages
 * [4] contraction     [5] total device                 [6] h2d   [7] d2h            */
int mem_pd_last_timings(mem_ctx* ctx, float* ms, int n);

/* ---- contraction alone (operands already on device) ------------------------------------
 * D = 4 * ( S1 S2^T + S2 S1^T - S3 S3^T ) over rows of Z = [S1 | S2 | S3]   
 * == |C|^2 (|F|^2)^T + its transpose - 2 Re(A A^H)  of :391-397. */
typedef struct {
  int32_t nS;        /* rows                                          */
  int32_t n1_blocks; /* 32-column blocks in S1 (== blocks in S2)      */
  int32_t n3_blocks; /* 32-column blocks in S3                        */
  int64_t ldz;       /* row pitch of Zhi/Zlo in floats (multiple of 4) */
} mem_contract_shape;
int mem_contract_device(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo,
                        float* D, int32_t contraction, int32_t k_chunk_blocks, int32_t split_k, void* stream);
/* the same with the kNN lists (a15) taken from the split-K partial tiles; D may be NULL (never assembled) */
int mem_contract_knn_device(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo,
                            float* D, int32_t k, int32_t* knn_idx, double* knn_val, int32_t contraction,
                            int32_t k_chunk_blocks, int32_t split_k, void* stream);
/* operand layout for a box size (columns of Z): fills n1_blocks, n3_blocks, ldz */
int mem_operand_shape(mem_ctx* ctx, int32_t N, mem_contract_shape* out);

/* ---- diffusion-map front end (DMembeddingII.op, modules/DMembeddingII.py:86-185) ---------- */
/* a15 kNN (DMembeddingII.initialize :43-57): for every column i of the symmetric matrix D (float64,
 * device), the k smallest entries with the diagonal forced first; idx [nS][k] int32, val [nS][k] f64
 * (val[i][0] = 0).  D is not modified. */
int mem_knn_device(mem_ctx* ctx, const double* D, int32_t nS, int32_t k, int32_t* idx, double* val, void* stream);
/* a15 on the float32 D that mem_pd_distance_device leaves on the device (D never visits the host) */
int mem_knn_device_f32(mem_ctx* ctx, const float* D, int32_t nS, int32_t k, int32_t* idx, double* val, void* stream);
/* D[posPath][:, posPath] of the trimming loop (modules/manifoldTrimmingAuto.py:50,63) on a D that stays on the device:
 * out[a][b] = D[sel[a]][sel[b]]; D [nS][nS] and out [m][m] are device arrays of 4-byte (float32) or 8-byte (float64)
 * elements, sel [m] int32 is a HOST array (copied before the call returns). */
int mem_gather_square_device(mem_ctx* ctx, const void* D, int32_t elem_bytes, int32_t nS, const int32_t* sel, int32_t m,
                             void* out, void* stream);
/* which kNN kernel the three kNN paths use (process-wide): 0 = automatic (radix selection of the k winners when
 * 4 k <= nS and the row fits shared memory, else a full bitonic sort of the row), 1 = always sort, 2 = selection
 * whenever it fits.  Both kernels return identical lists (ties by index); the switch exists for tests and timing. */
int mem_knn_mode(int32_t mode);
/* a16 OR-symmetrised kNN graph (DMembeddingII.op :113-140) in dense form: M [nS][nS] float64 device,
 * M[i][j] = d^2 of the union graph, 0 for the 'zero' (self) entries, -1 where there is no edge. */
int mem_graph_dense_device(mem_ctx* ctx, const int32_t* idx, const double* val, int32_t nS, int32_t k, double* M,
                           void* stream);
/* row-major compaction of the graph entries (M >= 0) into out[*count] (device, capacity >= number of edges;
 * nS*nS always suffices): lets the Ferguson sweep run over the edges only when k < nS.  Synchronises. */
int mem_graph_compact_device(mem_ctx* ctx, const double* M, int32_t nS, double* out, int64_t* count);
/* a17 Ferguson sweep (fergusonE.op :36-43): out[e] = log sum_{d2/(2 eps_e) < thr} exp(-d2/(2 eps_e)),
 * d2 [n] float64 device (negative entries = no edge, skipped), logEps [nEps] float64 HOST,
 * out [nEps] float64 HOST.  Synchronises. */
int mem_ferguson_device(mem_ctx* ctx, const double* d2, int64_t n, const double* logEps, int32_t nEps,
                        double thr, double* out);
/* a18 dense Gaussian-kernel Laplacian (slaplacianonFly.op :57-78) for the k = nS graph:
 * W = exp(-M/sigma^2) on the graph support (M >= 0 entries; negative = no edge), alpha = 1
 * normalisation, symmetric normalisation, L = |L + L^T|/2.  M, L: [nS][nS] float64 device. */
int mem_laplacian_dense_device(mem_ctx* ctx, const double* M, int32_t nS, double sigma, double* L, void* stream);
/* a19 operator application for the host eigen-solver (sembeddingonFly.op :27, scipy ARPACK eigsh): y = L x with
 * L [nS][nS] float64 on the device, x and y [nS] float64 on the HOST.  Synchronises. */
int mem_symv_host(mem_ctx* ctx, const double* L, int32_t nS, const double* x, double* y);

/* a19 on the device (replaces sembeddingonFly.py:27 eigsh(l, k = nEigs + 1, maxiter = 300)): Lanczos
 * with full re-orthogonalisation on the resident Laplacian.  mem_lanczos_steps_device enqueues steps j0 .. j1 - 1
 * (j0 == 0 also builds the deterministic start vector) without synchronising: V [ld_ab][nS] float64 Krylov basis (row per
 * vector), ab [2][ld_ab] float64 = alpha_j in row 0, beta_j (= ||w|| after step j - 1) in row 1; j1 < ld_ab.  The host reads
 * `ab` between blocks of steps (eigen-decomposition of the j x j tridiagonal matrix, residual estimates) and finally asks
 * for the Ritz vectors X [k][nS] = S^T V of the k <= 32 wanted pairs, S [j][k] float64 on the HOST. */
int mem_lanczos_steps_device(mem_ctx* ctx, const double* L, int32_t nS, double* V, double* ab, int32_t ld_ab, int32_t j0,
                             int32_t j1, void* stream);
int mem_lanczos_ritz_device(mem_ctx* ctx, const double* V, int32_t nS, int32_t j, const double* S_host, int32_t k, double* X,
                            void* stream);

/* ---- NLSA / psi analysis (modules/NLSA.py:23-158, get_wiener.py, svdRF.py, L2_distance.py) -------------
 * All float64, device pointers unless marked HOST.  `sel` [num] int32 (device) = posPath[PosPsi1], the snapshot order of the
 * psi being analysed; nI = num - ConOrder; Nh = N/2 + 1; E = psiTrunc.  The CTF planes must be even (CTF(-k) = CTF(k), true
 * of every CTF the distance stage writes): the Wiener-weighted sums are taken on the Hermitian half plane. */
/* once per PD: H [n][N][Nh] complex128 = rfft2(img[i]) * CTF[i], Ch [n][N][Nh] = CTF half planes (NLSA.py:73-76 hoisted) */
int mem_nlsa_spectra_device(mem_ctx* ctx, const double* img, const double* ctf, int32_t n, int32_t N, void* H, double* Ch,
                            void* stream);
/* NLSA.py:30-33: ConD [nI][nI] = sum_{i < ConOrder} DD[r + i][c + i], DD = D[sel][:, sel]; D [nAll][nAll] float32 or float64 */
int mem_nlsa_cond_device(mem_ctx* ctx, const void* D, int32_t elem_bytes, int32_t nAll, const int32_t* sel, int32_t num,
                         int32_t ConOrder, double* ConD, void* stream);
/* NLSA.py:66-86 + get_wiener.py: A [ConOrder N^2][E] = the Wiener-filtered, masked snapshot stack times mu_psi [nI][E] (HOST),
 * row ii N^2 + c N + r for picture pixel (r, c); msk2 [N][N] float64 or NULL (= 1).  Synchronises once (mu_psi upload). */
int mem_nlsa_supervectors_device(mem_ctx* ctx, const void* H, const double* Ch, const int32_t* sel, const double* mu_psi,
                                 int32_t num, int32_t ConOrder, int32_t E, int32_t N, const double* msk2, double* A,
                                 void* stream);
/* svdRF.py:21: AtA [E][E] (HOST) = A^T A.  Synchronises. */
int mem_nlsa_gram_small_device(mem_ctx* ctx, const double* A, int64_t rows, int32_t E, double* AtA, void* stream);
/* svdRF.py:25 + NLSA.py:95-103: U [rows][E] = A M with M [E][E] (HOST) = V S^-1; topo_mean [Npix][E] (HOST) = mean over the
 * ConOrder blocks of U.  Synchronises. */
int mem_nlsa_project_device(mem_ctx* ctx, const double* A, int64_t rows, int32_t E, const double* M, double* U, int32_t Npix,
                            int32_t ConOrder, double* topo_mean, void* stream);
/* NLSA.py:106-144: IMGT [nC][Npix] = frames rebuilt from the first two singular triplets (Q [2][nI] HOST = diag(s) V^T psiC^T),
 * each normalised to mean 0 / std 1; D2 [nC][nC] = L2_distance(IMGT, IMGT)**2 (NULL: skip). */
int mem_nlsa_reconstruct_device(mem_ctx* ctx, const double* U, int32_t Npix, int32_t ConOrder, int32_t E, const double* Q,
                                int32_t nI, int32_t nC, double* IMGT, double* D2, void* stream);

/* fit_1D_open_manifold_3D.op (modules/fit_1D_open_manifold_3D.py:66-144 with solve_d_R_d_tau_p_3D.py / R_p.py): the alternating
 * fit x_ij = a_j cos(j pi tau_i) + b_j of the three leading diffusion coordinates x [nS][3], started from ab = (a_1..3, b_1..3)
 * (get_fit_1D_open_manifold_3D_param.op, computed by the caller), at most max_iter iterations, stopping when the largest relative
 * change of a and of b (in percent) falls below da_max / db_max.  HOST pointers: ab [6] in / out, tau [nS] out, iters [1] out.
 * Synchronises. */
int mem_manifold_fit_host(mem_ctx* ctx, const double* x, int32_t nS, double* ab, double* tau, int32_t max_iter, double da_max,
                          double db_max, int32_t* iters);

/* ---- upstream of the distance stage: S2 tessellation (modules/S2tessellation.py) ---------- */
/* classS2 (:59-63): for every particle direction pts[i] (unit 3-vectors, [n][3] float64) the index of the nearest
 * bin centre (centres [nG][3] float64, Euclidean distance in float64, smallest index on a tie) -> idx [n] int32.
 * HOST pointers; synchronises. */
int mem_s2_assign_host(mem_ctx* ctx, const double* centres, int32_t nG, const double* pts, int64_t n, int32_t* idx);
/* FindCCGraph.CalcPairwiseDistS2 (modules/FindCCGraph.py:227-273): U [nU][3], V [nV][3] float64 = the selected columns
 * of the 3 x N matrix X, one point per row -> dot [nU][nV] = U V^T and dist [nU][nV] = sqrt(Dsq), Dsq < 1e-6 -> 0, with
 * Dsq[i][j] = (|u_j|^2 + |v_j|^2) - 2 u_i.v_j exactly as the reference's NumPy broadcast evaluates :267 (nU == nV
 * required, like the broadcast).  HOST pointers; synchronises. */
int mem_s2_pairwise_host(mem_ctx* ctx, const double* U, int32_t nU, const double* V, int32_t nV, double* dot, double* dist);

#ifdef __cplusplus
}
#endif
#endif /* MANIFOLDEM_B200_H */
