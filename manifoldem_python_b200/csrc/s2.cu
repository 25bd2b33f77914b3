// s2.cu — nearest tessellation bin for every particle orientation (S2tessellation.classS2,
// modules/S2tessellation.py:59-63: NearestNeighbors(n_neighbors=1).fit(bin centres).kneighbors(directions)).
// Brute force in float64: one thread per direction, the bin centres streamed through shared memory in tiles; the
// squared distance is accumulated term by term without FMA contraction (the value a float64 host loop gives) and
// the smallest index wins a tie.  n x nG x 8 flop: 4,071 bins x 10^6 particles is ~2 ms of the FP64 pipe where the
// reference builds and walks a ball tree on the host.
#include "common.cuh"

namespace mem {

constexpr int S2_TILE = 1024, S2_THREADS = 256;

__global__ void __launch_bounds__(S2_THREADS) k_s2_assign(const double* __restrict__ centres, int nG,
                                                          const double* __restrict__ pts, long long n,
                                                          int* __restrict__ idx) {
  __shared__ double c[S2_TILE * 3];
  const long long i = (long long)blockIdx.x * S2_THREADS + threadIdx.x;
  double x = 0, y = 0, z = 0;
  if (i < n) { x = pts[3 * i]; y = pts[3 * i + 1]; z = pts[3 * i + 2]; }
  double best = INFINITY;
  int arg = 0;
  for (int g0 = 0; g0 < nG; g0 += S2_TILE) {
    const int m = min(S2_TILE, nG - g0);
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * m; t += S2_THREADS) c[t] = centres[3 * (size_t)g0 + t];
    __syncthreads();
    for (int g = 0; g < m; ++g) {
      const double dx = x - c[3 * g], dy = y - c[3 * g + 1], dz = z - c[3 * g + 2];
      const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (d < best) { best = d; arg = g0 + g; }
    }
  }
  if (i < n) idx[i] = arg;
}

// centres [nG][3], pts [n][3] float64 and idx [n] int32 are HOST pointers; synchronises.
int s2_assign_host(mem_ctx* ctx, const double* centres, int nG, const double* pts, long long n, int* idx) {
  if (nG < 1 || n < 0) {
    set_error("s2_assign: need nG >= 1 and n >= 0 (nG=%d n=%lld)", nG, n);
    return 1;
  }
  if (n == 0) return 0;
  cudaStream_t st = ctx->stream;
  const size_t cb = (size_t)nG * 3 * sizeof(double), pb = (size_t)n * 3 * sizeof(double);
  const size_t pb_al = (pb + 255) & ~(size_t)255, cb_al = (cb + 255) & ~(size_t)255;
  MEM_CHECK(ctx->scratch.ensure(cb_al + pb_al + (size_t)n * sizeof(int)));
  uint8_t* base = ctx->scratch.as<uint8_t>();
  double* d_c = reinterpret_cast<double*>(base);
  double* d_p = reinterpret_cast<double*>(base + cb_al);
  int* d_i = reinterpret_cast<int*>(base + cb_al + pb_al);
  MEM_CUDA(cudaMemcpyAsync(d_c, centres, cb, cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(d_p, pts, pb, cudaMemcpyHostToDevice, st));
  MEM_LAUNCH(ctx, k_s2_assign, (unsigned)((n + S2_THREADS - 1) / S2_THREADS), S2_THREADS, 0, st, d_c, nG, d_p, n, d_i);
  MEM_CUDA(cudaMemcpyAsync(idx, d_i, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// FindCCGraph.CalcPairwiseDistS2 (modules/FindCCGraph.py:227-273): U = X[:, UIdxs], V = X[:, VIdxs] (3 x nU, 3 x nV);
//   pwDotProd = U^T V;   Dsq = sum(U*U, 0) + sum(V*V, 0) - 2 U^T V;   Dsq[Dsq < 1e-6] = 0;   pwDist = sqrt(Dsq).
// NumPy adds the two norm vectors ELEMENTWISE (both are 1-D, the `.T` of :267 is a no-op) and broadcasts the sum along
// the rows, so Dsq[i][j] = (|u_j|^2 + |v_j|^2) - 2 u_i.v_j — the reference's expression, kept as it stands (for the unit
// vectors the stage feeds it the two readings coincide); nU must equal nV, as NumPy's broadcast demands.
// One thread per pair, products and sums in the order a float64 loop gives (no FMA contraction).
__global__ void __launch_bounds__(256) k_s2_pairwise(const double* __restrict__ U, const double* __restrict__ V, int nU,
                                                     int nV, double* __restrict__ dot, double* __restrict__ dist) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (i >= nU || j >= nV) return;
  const double ux = U[3 * i], uy = U[3 * i + 1], uz = U[3 * i + 2];
  const double vx = V[3 * j], vy = V[3 * j + 1], vz = V[3 * j + 2];
  const double d = __dadd_rn(__dadd_rn(__dmul_rn(ux, vx), __dmul_rn(uy, vy)), __dmul_rn(uz, vz));
  const double wx = U[3 * j], wy = U[3 * j + 1], wz = U[3 * j + 2];     // column j of U: the broadcast of :267
  const double uu = __dadd_rn(__dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)), __dmul_rn(wz, wz));
  const double vv = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
  double q = __dsub_rn(__dadd_rn(uu, vv), __dmul_rn(2.0, d));
  if (q < 1e-6) q = 0.0;
  dot[(size_t)i * nV + j] = d;
  dist[(size_t)i * nV + j] = sqrt(q);
}

// U [nU][3], V [nV][3] (the selected columns of X, transposed), dot / dist [nU][nV]: HOST pointers; synchronises.
int s2_pairwise_host(mem_ctx* ctx, const double* U, int nU, const double* V, int nV, double* dot, double* dist) {
  if (nU != nV) {
    set_error("CalcPairwiseDistS2: the reference adds sum(U*U,0) and sum(V*V,0) elementwise, nU (%d) must equal nV (%d)", nU, nV);
    return 1;
  }
  if (nU <= 0) return 0;
  cudaStream_t st = ctx->stream;
  const size_t ub = ((size_t)nU * 3 * sizeof(double) + 255) & ~(size_t)255, ob = (size_t)nU * nV * sizeof(double);
  MEM_CHECK(ctx->scratch.ensure(2 * ub + 2 * ob));
  uint8_t* base = ctx->scratch.as<uint8_t>();
  double* d_u = reinterpret_cast<double*>(base);
  double* d_v = reinterpret_cast<double*>(base + ub);
  double* d_dot = reinterpret_cast<double*>(base + 2 * ub);
  double* d_dist = reinterpret_cast<double*>(base + 2 * ub + ob);
  MEM_CUDA(cudaMemcpyAsync(d_u, U, (size_t)nU * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaMemcpyAsync(d_v, V, (size_t)nV * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  MEM_LAUNCH(ctx, k_s2_pairwise, dim3((nV + 31) / 32, (nU + 7) / 8), 256, 0, st, d_u, d_v, nU, nV, d_dot, d_dist);
  MEM_CUDA(cudaMemcpyAsync(dot, d_dot, ob, cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaMemcpyAsync(dist, d_dist, ob, cudaMemcpyDeviceToHost, st));
  MEM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace mem
