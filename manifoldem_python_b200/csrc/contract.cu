// contract.cu — dispatcher for the distance contraction (row a12) plus the SIMT fp64-accumulate
// checker kernel.  The checker is NOT a fallback: the product path (contraction == 0) is the tcgen05
// kernel in contract_tc.cu and fails loudly if it cannot run; contraction == 1 must be asked for
// explicitly (tests use it to separate operand errors from tensor-core accumulation errors).
//
//   D = 4 * ( S1 S2^T + S2 S1^T - S3 S3^T ),   Z = [S1 | S2 | S3] = Zhi + Zlo
//
// which equals |C|^2 (|F|^2)^T + transpose - 2 Re(A A^H) of getDistanceCTF...py:391-397 with
//   S1 = C^2/4 per radial bin, S2 = radial power of F per bin, S3 = one representative of each
//   conjugate pair of A = C F.
#include "common.cuh"

namespace mem {

// 16x16 pairs per CTA, K streamed in 32-column blocks through shared memory; fp64 accumulation.
__global__ void __launch_bounds__(256) k_contract_simt(const float* __restrict__ zhi, const float* __restrict__ zlo,
                                                       float* __restrict__ D, int nS, int n1, int n3, int64_t ldz) {
  __shared__ float sa[16][33], sb[16][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  if (j0 + 15 < i0) return;   // strictly below the diagonal: mirrored later
  double acc = 0;
  const int nkb = 2 * n1 + n3;
  for (int kb = 0; kb < nkb; ++kb) {
    int ca, cb;
    double sign = 1.0;
    if (kb < n1) { ca = kb; cb = kb + n1; }
    else if (kb < 2 * n1) { ca = kb; cb = kb - n1; }
    else { ca = kb; cb = kb; sign = -1.0; }
    __syncthreads();
    for (int e = threadIdx.x; e < 512; e += 256) {
      const int r = e >> 5, c = e & 31;
      const int ia = i0 + r, jb = j0 + r;
      sa[r][c] = ia < nS ? zhi[(size_t)ia * ldz + 32 * ca + c] + zlo[(size_t)ia * ldz + 32 * ca + c] : 0.0f;
      sb[r][c] = jb < nS ? zhi[(size_t)jb * ldz + 32 * cb + c] + zlo[(size_t)jb * ldz + 32 * cb + c] : 0.0f;
    }
    __syncthreads();
    double part = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) part += (double)sa[ty][c] * (double)sb[tx][c];
    acc += sign * part;
  }
  const int i = i0 + ty, j = j0 + tx;
  if (i < nS && j < nS && i <= j) {
    const float v = (float)(4.0 * acc);
    D[(size_t)i * nS + j] = v;
    D[(size_t)j * nS + i] = v;
  }
}

int contract_run(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                 int contraction, int k_chunk_blocks, int split_k, cudaStream_t st, const KnnOut* knn) {
  const bool want_knn = knn && knn->k > 0 && knn->idx && knn->val;
  if (!D && !want_knn) {
    set_error("contraction: neither D nor kNN lists requested");
    return 1;
  }
  if (shp->nS <= 0 || shp->ldz < 32LL * (2 * shp->n1_blocks + shp->n3_blocks) || (shp->ldz & 3)) {
    set_error("bad contraction shape");
    return 1;
  }
  if (contraction == 1) {
    if (!D) {                                         // the checker always assembles D
      MEM_CHECK(ctx->D.ensure((size_t)shp->nS * shp->nS * sizeof(float)));
      D = ctx->D.as<float>();
    }
    dim3 grid((shp->nS + 15) / 16, (shp->nS + 15) / 16);
    MEM_LAUNCH(ctx, k_contract_simt, grid, 256, 0, st, Zhi, Zlo, D, shp->nS, shp->n1_blocks, shp->n3_blocks, shp->ldz);
    if (want_knn) MEM_CHECK(knn_device_f32(ctx, D, shp->nS, knn->k, knn->idx, knn->val, st));
    return 0;
  }
  if (contraction != 0 && contraction != 2) {
    set_error("unknown contraction kind %d", contraction);
    return 1;
  }
  // 0: product path, CTA-pair tiles (cta_group::2);  2: single-CTA tiles (kept for comparison and tests)
  return contract_tc(ctx, shp, Zhi, Zlo, D, k_chunk_blocks, split_k, st, contraction == 0 ? 1 : 0,
                     want_knn ? knn : nullptr);
}

}  // namespace mem
