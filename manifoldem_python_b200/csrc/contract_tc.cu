// contract_tc.cu — the distance contraction (row a12) on the 5th-generation tensor cores.
//
//   D = 4 * ( S1 S2^T + S2 S1^T - S3 S3^T )            rows of Z = [S1 | S2 | S3], Z = Zhi + Zlo
//
// One GEMM-shaped pass over a K axis of 32-column blocks.  A k-block in S1 pairs A = S1 block with
// B = the matching S2 block, a k-block in S2 pairs A = S2 with B = S1, and the S3 blocks pair with
// themselves with the A operand negated in the instruction descriptor, so the whole of
// |C|^2 (|F|^2)^T + transpose - 2 Re(A A^H)  (getDistanceCTF...py:391-397) accumulates in ONE
// TMEM accumulator.  FP32 accuracy comes from 3xTF32: hi*hi + hi*lo + lo*hi per k-step, and from
// promoting the TMEM accumulator into FP32 registers every `chunk` k-blocks (tensor-core
// accumulation error grows with the number of accumulate steps; SURVEY §7 hard part 1).
//
// CTA = one (128 x 256 tile of D, K-slice) work item, upper triangle of tiles only.
//   warp 0      TMA producer  (cp.async.bulk.tensor, 128B swizzle, 2 stages x 96 KB)
//   warp 1      tcgen05.mma issuer (one lane), owns the TMEM allocation (512 columns = 2 accumulators)
//   warps 2..9  epilogue: tcgen05.ld the finished chunk, add into registers, release the accumulator;
//               after the last chunk store the partial tile to the split-K workspace
// k_contract_finalize sums the K-slices in a fixed order, scales by 4 and mirrors to the lower triangle.
#include "common.cuh"

#include <cuda.h>
#include <algorithm>
#include <vector>

namespace mem {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int STAGES = 2;
constexpr int BOX_ROWS = 128;
constexpr int BOX_BYTES = BOX_ROWS * BK * 4;            // 16 KB
constexpr int STAGE_BYTES = 6 * BOX_BYTES;              // A_hi, A_lo, B_hi(2), B_lo(2) = 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 320;

// row0 / col0: tile origin inside its PD; zbase: first row of that PD in the (possibly concatenated) operand matrix;
// nrows / ldw / ws_off: the PD's size, workspace pitch and workspace offset (floats) — a grouped launch carries the tiles
// of several PDs in one work list (pd_distance_batch_device); pad0 / pad1: pacing group and its size.
struct WorkItem { int row0, col0, kb0, kb1, slice, pad0, pad1, zbase, nrows, ldw; long long ws_off; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a broken pipeline traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows at 128 B pitch, 8-row groups 1024 B apart (SBO), LBO unused (=1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, K-major A and B, M=128, N=256; bit 13 negates A
__device__ __forceinline__ uint32_t make_idesc(bool negate_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((negate_a ? 1u : 0u) << 13) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_contract_tc(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
              const WorkItem* __restrict__ items, float* __restrict__ ws, int nS, int ldw, int n1, int chunk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barrier slots (8 B each): full[0..1], empty[2..3], tmem_full[4..5], tmem_empty[6..7]; tmem base at +64
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WorkItem it = items[blockIdx.x];
  // promotion period: `chunk & 255` k-blocks inside the S1/S2 columns, `chunk >> 8` (if set) inside S3
  const int c12 = chunk & 255, c3 = (chunk >> 8) ? (chunk >> 8) : c12;
  const int kmid = min(max(2 * n1, it.kb0), it.kb1);
  const int nchunks = (kmid - it.kb0 + c12 - 1) / c12 + (it.kb1 - kmid + c3 - 1) / c3;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (2 + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (4 + b), 1);
      mbar_init(bars + 8 * (6 + b), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        int ca, cb;
        if (kb < n1) { ca = kb; cb = kb + n1; }
        else if (kb < 2 * n1) { ca = kb; cb = kb - n1; }
        else { ca = kb; cb = kb; }
        mbar_wait(bars + 8 * (2 + stage), phase ^ 1);
        const uint32_t full = bars + 8 * stage;
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t s0 = base + stage * STAGE_BYTES;
        tma_load_2d(s0 + 0 * BOX_BYTES, &map_hi, full, ca * BK, it.row0);
        tma_load_2d(s0 + 1 * BOX_BYTES, &map_lo, full, ca * BK, it.row0);
        tma_load_2d(s0 + 2 * BOX_BYTES, &map_hi, full, cb * BK, it.col0);
        tma_load_2d(s0 + 3 * BOX_BYTES, &map_hi, full, cb * BK, it.col0 + BOX_ROWS);
        tma_load_2d(s0 + 4 * BOX_BYTES, &map_lo, full, cb * BK, it.col0);
        tma_load_2d(s0 + 5 * BOX_BYTES, &map_lo, full, cb * BK, it.col0 + BOX_ROWS);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      int stage = 0;
      uint32_t phase = 0;
      int kb = it.kb0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(bars + 8 * (6 + buf), (((uint32_t)(c >> 1)) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * BN;
        const int kend = kb < kmid ? min(kmid, kb + c12) : min(it.kb1, kb + c3);
        bool first = true;
        for (; kb < kend; ++kb) {
          const uint32_t idesc = make_idesc(kb >= 2 * n1);
          mbar_wait(bars + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t s0 = base + stage * STAGE_BYTES;
          // the small cross terms (2^-11 of the products) first: on the first block of a chunk they accumulate from
          // zero, where the truncating accumulator loses nothing that matters; then the four hi*hi steps
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t a_hi = make_smem_desc(s0 + 0 * BOX_BYTES + ks * 32);
            const uint64_t a_lo = make_smem_desc(s0 + 1 * BOX_BYTES + ks * 32);
            const uint64_t b_hi = make_smem_desc(s0 + 2 * BOX_BYTES + ks * 32);
            const uint64_t b_lo = make_smem_desc(s0 + 4 * BOX_BYTES + ks * 32);
            umma_tf32(tacc, a_lo, b_hi, idesc, first ? 0u : 1u);
            umma_tf32(tacc, a_hi, b_lo, idesc, 1u);
            first = false;
          }
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t a_hi = make_smem_desc(s0 + 0 * BOX_BYTES + ks * 32);
            const uint64_t b_hi = make_smem_desc(s0 + 2 * BOX_BYTES + ks * 32);
            umma_tf32(tacc, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(bars + 8 * (2 + stage));   // smem slot free when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(bars + 8 * (4 + buf));       // accumulator `buf` holds a finished chunk
      }
    }
  } else {
    // ===== epilogue: 8 warps, lane quarter q = warp % 4, column half h =====
    const int q = warp & 3, h = (warp - 2) >> 2;
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bars + 8 * (4 + buf), ((uint32_t)(c >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + h * 128;
#pragma unroll
      for (int gI = 0; gI < 4; ++gI) {
        uint32_t v[32];
        tmem_ld32(taddr + gI * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[gI * 32 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (6 + buf));
    }
    const int row = it.row0 + q * 32 + lane;
    const int col0 = it.col0 + h * 128;
    if (row < nS) {
      float* dst = ws + ((size_t)it.slice * ldw + row) * ldw + col0;
#pragma unroll
      for (int j = 0; j < 128; j += 4) {
        if (col0 + j < ldw)
          *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}


// ------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  CTA r holds rows
// [128 r, 128 r + 128) of the A operand and rows [128 r, +128) of the B operand (the N side) in its own shared
// memory, so every SM reads only 4 KB + 4 KB per 256x256x8 MMA instead of 4 + 8 KB: the 1-CTA kernel is
// bound by shared-memory bandwidth (operand reads + TMA fill ~ 156 B/clk/SM), this one needs ~ 106 B/clk/SM.
//   stage (per CTA) = A_hi, A_lo, B_hi, B_lo boxes of 128 rows x 32 columns = 64 KB, 3 stages
//   full[s]        leader's barrier; both CTAs' TMA complete_tx on it (peer bit of the address cleared)
//   empty[s]       each CTA's own barrier, released by the leader's tcgen05.commit multicast
//   tmem_full[b]   each CTA's own barrier (commit multicast); tmem_empty[b] leader's, 32 arrivals (16 warps x 2 CTAs)
// 16 epilogue warps per CTA (lane quarter x 64-column group) keep the per-chunk TMEM drain to two
// tcgen05.ld round trips, so the accumulator turns around well inside one chunk of MMAs.
// ------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 3;
constexpr int PACE_M = 32, PACE_W = 2;                  // milestone spacing (k-blocks) and allowed lead (milestones)
constexpr int STAGE2_BYTES = 4 * BOX_BYTES;             // 64 KB per CTA
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr int EPI_WARPS2 = 16;
constexpr int NUM_THREADS2 = 64 + 32 * EPI_WARPS2;      // TMA warp, MMA warp, 16 epilogue warps

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// M = 256 (two CTAs x 128), N = 256
__device__ __forceinline__ uint32_t make_idesc2(bool negate_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((negate_a ? 1u : 0u) << 13) | ((uint32_t)(256 >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
k_contract_tc2(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const WorkItem* __restrict__ items, float* __restrict__ ws, int nS, int ldw, int n1, int chunk,
               int* __restrict__ prog, int n_mil, unsigned long long* __restrict__ clk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES2 * STAGE2_BYTES;
  // clock probe (bench.py roofline): SM cycles and wall nanoseconds over the life of CTA 0 -> the SM clock this launch
  // actually ran at (the kernel is power-capped well below the clock nvidia-smi samples between launches)
  if (clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    clk[0] = (unsigned long long)clock64();
    clk[1] = t;
  }
  // 8-byte slots: full[0..2], empty[3..5], tmem_full[6..7], tmem_empty[8..9]; tmem base at +96
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES2 * STAGE2_BYTES + 96);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const WorkItem it = items[blockIdx.x >> 1];
  // promotion period: `chunk & 255` k-blocks inside the S1/S2 columns, `chunk >> 8` (if set) inside S3
  const int c12 = chunk & 255, c3 = (chunk >> 8) ? (chunk >> 8) : c12;
  const int kmid = min(max(2 * n1, it.kb0), it.kb1);
  const int nchunks = (kmid - it.kb0 + c12 - 1) / c12 + (it.kb1 - kmid + c3 - 1) / c3;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (3 + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (6 + b), 1);
      mbar_init(bars + 8 * (8 + b), 2 * EPI_WARPS2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs; each loads its own halves, transaction bytes land on the leader's barrier) =====
      int stage = 0;
      uint32_t phase = 0;
      const int arow = it.zbase + it.row0 + (int)rank * BOX_ROWS, brow = it.zbase + it.col0 + (int)rank * BOX_ROWS;
      // Pacing (multi-wave launches only): the tiles of a super-block share operand panels through the L2 only while
      // they read the same k-blocks at about the same time.  Every PACE_M k-blocks the leader's producer counts
      // itself in at a milestone and does not run more than PACE_W milestones ahead of the slowest tile of its
      // group (bounded wait: a group that is not fully resident falls back to free running).
      bool pace = prog != nullptr && rank == 0 && it.pad1 > 1;
      int* gp = prog ? prog + (size_t)it.pad0 * n_mil : nullptr;
      for (int kb = it.kb0; kb < it.kb1; ++kb) {
        if (pace) {
          const int rel = kb - it.kb0;
          if (rel > 0 && (rel & (PACE_M - 1)) == 0) {
            const int m = rel / PACE_M;
            atomicAdd(gp + m, 1);
            if (m > PACE_W) {                           // milestone 0 is the start and is never counted
              const volatile int* w = gp + (m - PACE_W);
              const long long t0 = clock64();
              while (*w < it.pad1) {
                if (clock64() - t0 > 1000000LL) { pace = false; break; }
              }
            }
          }
        }
        int ca, cb;
        if (kb < n1) { ca = kb; cb = kb + n1; }
        else if (kb < 2 * n1) { ca = kb; cb = kb - n1; }
        else { ca = kb; cb = kb; }
        mbar_wait(bars + 8 * (3 + stage), phase ^ 1);
        const uint32_t full = bars + 8 * stage;
        if (rank == 0) mbar_expect_tx(full, 2 * STAGE2_BYTES);
        const uint32_t s0 = base + stage * STAGE2_BYTES;
        tma_load_2d_2sm(s0 + 0 * BOX_BYTES, &map_hi, full, ca * BK, arow);
        tma_load_2d_2sm(s0 + 1 * BOX_BYTES, &map_lo, full, ca * BK, arow);
        tma_load_2d_2sm(s0 + 2 * BOX_BYTES, &map_hi, full, cb * BK, brow);
        tma_load_2d_2sm(s0 + 3 * BOX_BYTES, &map_lo, full, cb * BK, brow);
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader CTA only) =====
      int stage = 0;
      uint32_t phase = 0;
      int kb = it.kb0;
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        mbar_wait(bars + 8 * (8 + buf), (((uint32_t)(c >> 1)) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + buf * BN;
        const int kend = kb < kmid ? min(kmid, kb + c12) : min(it.kb1, kb + c3);
        bool first = true;
        for (; kb < kend; ++kb) {
          const uint32_t idesc = make_idesc2(kb >= 2 * n1);
          mbar_wait(bars + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t s0 = base + stage * STAGE2_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {          // cross terms first (see the single-CTA kernel)
            const uint64_t a_hi = make_smem_desc(s0 + 0 * BOX_BYTES + ks * 32);
            const uint64_t a_lo = make_smem_desc(s0 + 1 * BOX_BYTES + ks * 32);
            const uint64_t b_hi = make_smem_desc(s0 + 2 * BOX_BYTES + ks * 32);
            const uint64_t b_lo = make_smem_desc(s0 + 3 * BOX_BYTES + ks * 32);
            umma_tf32_2sm(tacc, a_lo, b_hi, idesc, first ? 0u : 1u);
            umma_tf32_2sm(tacc, a_hi, b_lo, idesc, 1u);
            first = false;
          }
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t a_hi = make_smem_desc(s0 + 0 * BOX_BYTES + ks * 32);
            const uint64_t b_hi = make_smem_desc(s0 + 2 * BOX_BYTES + ks * 32);
            umma_tf32_2sm(tacc, a_hi, b_hi, idesc, 1u);
          }
          umma_commit_2sm(bars + 8 * (3 + stage));   // both CTAs' smem slots free when these MMAs retire
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(bars + 8 * (6 + buf));       // accumulator `buf` (both CTAs' halves) holds a finished chunk
      }
    }
  } else {
    // ===== epilogue: 16 warps per CTA on this CTA's 128 accumulator rows; warp = (lane quarter q, 64-column group) =====
    const int q = warp & 3, cg = (warp - 2) >> 2;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bars + 8 * (6 + buf), ((uint32_t)(c >> 1)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + cg * 64;
#pragma unroll
      for (int gI = 0; gI < 2; ++gI) {
        uint32_t v[32];
        tmem_ld32(taddr + gI * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (gI == 1) {                       // all TMEM reads of this chunk are done: release the accumulator first
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cta0(bars + 8 * (8 + buf));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[gI * 32 + j] += __uint_as_float(v[j]);
      }
    }
    const int row = it.row0 + (int)rank * BOX_ROWS + q * 32 + lane;
    const int col0 = it.col0 + cg * 64;
    if (row < it.nrows) {
      float* dst = ws + it.ws_off + ((size_t)it.slice * it.ldw + row) * it.ldw + col0;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        if (col0 + j < it.ldw)
          *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();     // the peer may still be reading its TMEM half / arriving on the leader's barriers
  if (clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    clk[2] = (unsigned long long)clock64();
    clk[3] = t;
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// D[i][j] = D[j][i] = 4 * sum_s ws[s][i][j]  (i <= j), 32x32 tiles, coalesced both ways
__global__ void __launch_bounds__(256) k_contract_finalize(const float* __restrict__ ws, float* __restrict__ D, int nS,
                                                           int ldw, int nslices) {
  __shared__ float t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    float v = 0.0f;
    if (i < nS && j < nS)
      for (int s = 0; s < nslices; ++s) v += ws[((size_t)s * ldw + i) * ldw + j];
    t[r][tx] = 4.0f * v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < nS && j < nS) {
      if (bi != bj) D[(size_t)i * nS + j] = t[r][tx];
      else D[(size_t)i * nS + j] = (r <= tx) ? t[r][tx] : t[tx][r];
    }
    if (bi != bj) {
      const int i2 = bj * 32 + r, j2 = bi * 32 + tx;   // mirrored tile, row of the column block
      if (i2 < nS && j2 < nS) D[(size_t)i2 * nS + j2] = t[tx][r];
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(mem_ctx* ctx, CUtensorMap* map, const float* Z, int nS, int64_t ldz) {
  if (!ctx->tmap_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MEM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return 1;
    }
    ctx->tmap_encode = fn;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)ldz, (cuuint64_t)nS};
  cuuint64_t gstr[1] = {(cuuint64_t)ldz * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BOX_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)ctx->tmap_encode)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)Z, gdim, gstr, box,
                                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d", (int)r);
    return 1;
  }
  return 0;
}

int contract_tc(mem_ctx* ctx, const mem_contract_shape* shp, const float* Zhi, const float* Zlo, float* D,
                int k_chunk_blocks, int split_k, cudaStream_t st, int two_cta, const KnnOut* knn) {
  const int nS = shp->nS, n1 = shp->n1_blocks, n3 = shp->n3_blocks;
  const int nkb = 2 * n1 + n3;
  if (((uintptr_t)Zhi & 15) || ((uintptr_t)Zlo & 15)) {
    set_error("contraction operands must be 16-byte aligned");
    return 1;
  }
  int dev_cc_major = 0;
  MEM_CUDA(cudaDeviceGetAttribute(&dev_cc_major, cudaDevAttrComputeCapabilityMajor, ctx->device));
  if (dev_cc_major != 10) {
    set_error("tcgen05 contraction needs an sm_100a device (found compute capability major %d); there is no fallback", dev_cc_major);
    return 1;
  }
  // default promotion period: 1 k-block for single-CTA tiles; 2 for CTA pairs, whose accumulator hand-over crosses
  // the cluster twice per chunk (commit multicast + remote arrive ~ 650 clk) and needs the longer chunk to hide it
  // The S3 columns (mixed-sign products, 78 % of K) tolerate 4 blocks per promotion at the same measured accuracy
  // (tests/tools/gpu_debug.py chunk3): code = period(S1,S2) | period(S3) << 8.
  const int chunk = k_chunk_blocks > 0 ? k_chunk_blocks : (two_cta ? (2 | 4 << 8) : 1);
  // tiles of the upper triangle (any element with col >= row)
  std::vector<std::pair<int, int>> tiles;
  std::vector<int> tile_sb, sb_count;                    // super-block of every tile, tiles per super-block
  const int TM = two_cta ? 2 * BM : BM;                  // tile rows: a CTA pair covers 256
  const int units = two_cta ? ctx->sm_count / 2 : ctx->sm_count;
  const int tm = (nS + TM - 1) / TM, tn = (nS + BN - 1) / BN;
  // Order: compact super-blocks of about `units` tiles.  The CTAs of a wave start together and consume K at the same
  // rate, so at any moment they read the same k-block of sbi row panels + sbj column panels (17 for CTA pairs)
  // instead of one column panel + 74 row panels: at C5 size (20,000 x 320^2, 247 MB per panel, far beyond L2) the
  // column-by-column order needs ~4.9 TB/s of HBM and the contraction turns memory-bound.
  const int sbi = std::max(1, (int)floor(sqrt((double)units))), sbj = std::max(1, units / sbi);
  for (int sj = 0; sj < tn; sj += sbj)
    for (int si = 0; si < tm; si += sbi) {
      int cnt = 0;
      for (int bj = sj; bj < std::min(tn, sj + sbj); ++bj)
        for (int bi = si; bi < std::min(tm, si + sbi); ++bi)
          if (bj * BN + BN - 1 >= bi * TM) {
            tiles.push_back({bi, bj});
            tile_sb.push_back((int)sb_count.size());
            ++cnt;
          }
      if (cnt) sb_count.push_back(cnt);
    }
  const int T = (int)tiles.size();
  int split = split_k;
  if (split <= 0) {
    const int min_kb = 4 * (chunk & 255);
    double best = -1;
    split = 1;
    for (int s = 1; s <= 32; ++s) {
      if (s > 1 && nkb / s < min_kb) break;
      const int items = T * s;
      const int waves = (items + units - 1) / units;
      const double eff = (double)items / ((double)waves * units);
      if (eff > best + 0.02) { best = eff; split = s; }
    }
  }
  split = std::max(1, std::min(split, nkb));
  const int ldw = ((nS + 3) / 4) * 4;
  std::vector<WorkItem> items;
  for (int s = 0; s < split; ++s) {
    const int kb0 = (int)((long long)nkb * s / split), kb1 = (int)((long long)nkb * (s + 1) / split);
    for (size_t ti = 0; ti < tiles.size(); ++ti)
      items.push_back({tiles[ti].first * TM, tiles[ti].second * BN, kb0, kb1, s, 0, 0, 0, nS, ldw, 0});
  }
  // pacing groups = waves: `units` consecutive work items are resident together (all items have the same length)
  const int n_groups = ((int)items.size() + units - 1) / units;
  for (size_t k = 0; k < items.size(); ++k) {
    const int g = (int)k / units;
    items[k].pad0 = g;
    items[k].pad1 = std::min(units, (int)items.size() - g * units);
  }
  // pacing counters: only when the launch takes more than one wave of CTA pairs (C3 / C5-sized PDs)
  int* prog = nullptr;
  const int n_mil = (nkb + split - 1) / split / PACE_M + 2;
  if (two_cta && (int)items.size() > units) {
    const size_t bytes = (size_t)n_groups * n_mil * sizeof(int);
    MEM_CHECK(ctx->scratch.ensure(bytes));
    prog = ctx->scratch.as<int>();
    MEM_CUDA(cudaMemsetAsync(prog, 0, bytes, st));
  }
  MEM_CHECK(ctx->clk_probe.ensure(4 * sizeof(unsigned long long)));
  MEM_CHECK(ctx->contract_ws.ensure((size_t)split * ldw * ldw * sizeof(float)));
  float* ws = ctx->contract_ws.as<float>();
  const long long key3 = (long long)items.size() * 2 + two_cta;
  if (ctx->items_key[0] != nS || ctx->items_key[1] != nkb || ctx->items_key[2] != split || ctx->items_key[3] != key3) {
    MEM_CHECK(ctx->contract_items.ensure(items.size() * sizeof(WorkItem)));
    MEM_CUDA(cudaMemcpyAsync(ctx->contract_items.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st));
    MEM_CUDA(cudaStreamSynchronize(st));   // items is a host temporary; cached per shape afterwards
    ctx->items_key[0] = nS; ctx->items_key[1] = nkb; ctx->items_key[2] = split; ctx->items_key[3] = key3;
  }
  WorkItem* d_items = ctx->contract_items.as<WorkItem>();
  CUtensorMap map_hi, map_lo;
  MEM_CHECK(make_map(ctx, &map_hi, Zhi, nS, shp->ldz));
  MEM_CHECK(make_map(ctx, &map_lo, Zlo, nS, shp->ldz));
  MEM_CUDA(cudaFuncSetAttribute(k_contract_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  MEM_CUDA(cudaFuncSetAttribute(k_contract_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  // per-launch device timing (read back by mem_ctx_kernel_time; bench.py roofline)
  if (ctx->kev_used + 2 > ctx->kev.size()) {
    for (int i = 0; i < 64; ++i) {
      cudaEvent_t e;
      MEM_CUDA(cudaEventCreate(&e));
      ctx->kev.push_back(e);
    }
  }
  MEM_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used], st));
  if (two_cta)
    MEM_LAUNCH(ctx, k_contract_tc2, 2 * (int)items.size(), NUM_THREADS2, SMEM2_BYTES, st, map_hi, map_lo, d_items, ws, nS, ldw, n1, chunk,
               prog, n_mil, ctx->clk_probe.as<unsigned long long>());
  else
    MEM_LAUNCH(ctx, k_contract_tc, (int)items.size(), NUM_THREADS, SMEM_BYTES, st, map_hi, map_lo, d_items, ws, nS, ldw, n1, chunk);
  MEM_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used + 1], st));
  ctx->kev_used += 2;
  ctx->last_tc_items = (two_cta ? 2 : 1) * (int)items.size();   // CTAs, each computing a 128 x 256 tile of its slice
  ctx->last_tc_nkb = (nkb + split - 1) / split;                   // K blocks per CTA
  // a15 behind a12: the neighbour lists are selected straight from the partial tiles; D is assembled only when
  // the caller asked for it, or when the list is too long a part of the row for the selection kernel
  bool knn_done = false;
  if (knn) {
    const int rc = knn_from_workspace(ctx, ws, ldw, split, nS, knn->k, knn->idx, knn->val, st);
    if (rc == 1) return 1;
    knn_done = (rc == 0);
    if (!knn_done && !D) {
      MEM_CHECK(ctx->D.ensure((size_t)nS * nS * sizeof(float)));
      D = ctx->D.as<float>();
    }
  }
  if (D) {
    dim3 fgrid((nS + 31) / 32, (nS + 31) / 32);
    MEM_LAUNCH(ctx, k_contract_finalize, fgrid, 256, 0, st, ws, D, nS, ldw, split);
  }
  if (knn && !knn_done) MEM_CHECK(knn_device_f32(ctx, D, nS, knn->k, knn->idx, knn->val, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Grouped launch: the tiles of SEVERAL PDs in one work list (GetDistancesS2.py:94-120 produces PDs of 117..2,000
// particles: one such PD is 1-36 tiles, far too few for 74 SM pairs, and a launch + work-table upload per PD costs more
// than its tiles).  Zhi / Zlo hold the operand rows of all PDs back to back (PD g = rows pd_start[g] .. pd_start[g+1]);
// one tensor map covers them; a tile that hangs over the end of its PD reads the next PD's rows (finite numbers that
// land in rows / columns the epilogue never stores), past the last PD the TMA zero-fills.  K is cut into slices of one
// common length so every work item costs the same and the list fills whole waves of CTA pairs.  D[g] = device pointer
// of PD g's nS_g x nS_g output.
int contract_tc_grouped(mem_ctx* ctx, const mem_contract_shape* shp, int n_pd, const int* pd_start, const float* Zhi,
                        const float* Zlo, float* const* D, cudaStream_t st) {
  const int n1 = shp->n1_blocks, n3 = shp->n3_blocks, nkb = 2 * n1 + n3;
  const int units = ctx->sm_count / 2, TM = 2 * BM;
  const int chunk = 2 | 4 << 8;
  int total_tiles = 0;
  for (int g = 0; g < n_pd; ++g) {
    const int n = pd_start[g + 1] - pd_start[g];
    if (n < 1) {
      set_error("contract_tc_grouped: PD %d is empty", g);
      return 1;
    }
    const int tm = (n + TM - 1) / TM;
    total_tiles += tm * (tm + 1) / 2;
  }
  // slices per tile: enough items for >= 4 waves (or as many as the minimum slice length allows), whole waves preferred
  const int min_kb = 4 * (chunk & 255) * 8;
  int split = 1;
  double best = -1;
  for (int s2 = 1; s2 <= 64; ++s2) {
    if (s2 > 1 && nkb / s2 < min_kb) break;
    const long long items = (long long)total_tiles * s2;
    const long long waves = (items + units - 1) / units;
    const double eff = (double)items / ((double)waves * units);
    if (eff > best + 0.02) { best = eff; split = s2; }
  }
  std::vector<WorkItem> items;
  std::vector<long long> ws_off(n_pd);
  long long ws_total = 0;
  for (int g = 0; g < n_pd; ++g) {
    const int n = pd_start[g + 1] - pd_start[g];
    const int ldw = ((n + 3) / 4) * 4;
    ws_off[g] = ws_total;
    ws_total += (long long)split * ldw * ldw;
  }
  for (int s2 = 0; s2 < split; ++s2) {
    const int kb0 = (int)((long long)nkb * s2 / split), kb1 = (int)((long long)nkb * (s2 + 1) / split);
    for (int g = 0; g < n_pd; ++g) {
      const int n = pd_start[g + 1] - pd_start[g];
      const int tm = (n + TM - 1) / TM, ldw = ((n + 3) / 4) * 4;
      for (int bj = 0; bj < tm; ++bj)
        for (int bi = 0; bi <= bj; ++bi)
          items.push_back({bi * TM, bj * BN, kb0, kb1, s2, 0, 1, pd_start[g], n, ldw, ws_off[g]});
    }
  }
  MEM_CHECK(ctx->contract_ws.ensure((size_t)ws_total * sizeof(float)));
  MEM_CHECK(ctx->clk_probe.ensure(4 * sizeof(unsigned long long)));
  MEM_CHECK(ctx->contract_items.ensure(items.size() * sizeof(WorkItem)));
  MEM_CUDA(cudaMemcpyAsync(ctx->contract_items.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st));
  MEM_CUDA(cudaStreamSynchronize(st));                     // items is a host temporary: one upload per GROUP of PDs
  ctx->items_key[0] = -1;                                  // the single-PD cache no longer describes the table
  float* ws = ctx->contract_ws.as<float>();
  const int rows_all = pd_start[n_pd];
  CUtensorMap map_hi, map_lo;
  MEM_CHECK(make_map(ctx, &map_hi, Zhi, rows_all, shp->ldz));
  MEM_CHECK(make_map(ctx, &map_lo, Zlo, rows_all, shp->ldz));
  MEM_CUDA(cudaFuncSetAttribute(k_contract_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
  if (ctx->kev_used + 2 > ctx->kev.size()) {
    for (int i = 0; i < 64; ++i) {
      cudaEvent_t e;
      MEM_CUDA(cudaEventCreate(&e));
      ctx->kev.push_back(e);
    }
  }
  MEM_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used], st));
  MEM_LAUNCH(ctx, k_contract_tc2, 2 * (int)items.size(), NUM_THREADS2, SMEM2_BYTES, st, map_hi, map_lo,
             ctx->contract_items.as<WorkItem>(), ws, rows_all, 0, n1, chunk, (int*)nullptr, 0,
             ctx->clk_probe.as<unsigned long long>());
  MEM_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used + 1], st));
  ctx->kev_used += 2;
  ctx->last_tc_items = 2 * (int)items.size();
  ctx->last_tc_nkb = (nkb + split - 1) / split;
  for (int g = 0; g < n_pd; ++g) {
    const int n = pd_start[g + 1] - pd_start[g], ldw = ((n + 3) / 4) * 4;
    dim3 fgrid((n + 31) / 32, (n + 31) / 32);
    MEM_LAUNCH(ctx, k_contract_finalize, fgrid, 256, 0, st, ws + ws_off[g], D[g], n, ldw, split);
  }
  return 0;
}

}  // namespace mem
