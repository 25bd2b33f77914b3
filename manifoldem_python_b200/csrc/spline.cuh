// spline.cuh — the periodic cubic B-spline prefilter recursion shared by align.cu and lowpass.cu
// (what scipy.ndimage's spline_filter amounts to on the 3x3-tiled image of rotatefill.py:21-25):
//   c+[k] = 6 s[k] + z c+[k-1],   c[k] = z (c[k+1] - c+[k]),   z = sqrt(3) - 2.
#pragma once

#define SPL_Z (-0.26794919243112270647f)
#define SPL_REACH 20   // samples after which z^n is dropped (z^20 < 4e-12)

namespace mem {

// One line of 32 * E samples held by a warp: lane l owns v[0..E) = 6 * samples [l E, l E + E).  Every lane runs the
// causal recursion from a zero carry; the true carry is the z-weighted sum of the H = SPL_REACH / E + 2 previous
// lanes' local ends (periodic, through shuffles); then the same backwards.  zE = z^E.
template <int E>
__device__ __forceinline__ void spline_line_warp(float (&v)[E], float zE, int lane) {
  constexpr int H = SPL_REACH / E + 2;
  float run = 0.0f;
#pragma unroll
  for (int j = 0; j < E; ++j) { run = fmaf(SPL_Z, run, v[j]); v[j] = run; }
  float carry = 0.0f, f = 1.0f;
#pragma unroll
  for (int h = 1; h <= H; ++h) {
    carry = fmaf(f, __shfl_sync(0xffffffffu, run, (lane - h) & 31), carry);
    f *= zE;
  }
  float zp = SPL_Z * carry;
#pragma unroll
  for (int j = 0; j < E; ++j) { v[j] += zp; zp *= SPL_Z; }
  run = 0.0f;
#pragma unroll
  for (int j = E - 1; j >= 0; --j) { run = SPL_Z * (run - v[j]); v[j] = run; }
  carry = 0.0f; f = 1.0f;
#pragma unroll
  for (int h = 1; h <= H; ++h) {
    carry = fmaf(f, __shfl_sync(0xffffffffu, run, (lane + h) & 31), carry);
    f *= zE;
  }
  zp = SPL_Z * carry;
#pragma unroll
  for (int j = E - 1; j >= 0; --j) { v[j] += zp; zp *= SPL_Z; }
}

}  // namespace mem
