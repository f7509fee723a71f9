// Device-side building blocks shared by the dNMF kernels (sm_100a).
// Math spec: SURVEY.md Appendix A; reference lines cited are relative to the reference tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dnmf {

constexpr int kWarpX = 8;   // warp footprint in x (lanes & 7)
constexpr int kWarpY = 4;   // warp footprint in y (lanes >> 3)
constexpr int kBasis = 10;  // [1,x,y,z,x2,y2,z2,xy,xz,yz]  (Demix/dNMF.py:47-51)
constexpr int kNumPartials = 32;  // 30 gradient moments + sse + pad
constexpr int kWarpScratch = 96;  // floats per warp of the fused kernel's cross-warp reduction area

// Un-normalised sample coordinate with the reference's fp32 op order:
//   u  = fl(fl(fl(2 q) / (s-1)) - 1)                      Demix/dNMF.py:55
//   ix = fl(fl(fl(u + 1) / 2) * (s-1))                    ATen grid_sampler unnormalize
// True division, no FMA contraction, no reciprocal.  A singleton axis (s == 1) makes the
// reference divide by zero; here it samples at q unchanged (documented deviation).
__device__ __forceinline__ float sample_coord(float q, float sm1) {
  if (sm1 == 0.f) return q;
  float u = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, q), sm1), 1.f);
  return __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), sm1);
}

// un-normalised sample coordinate, fast form (see verify_coord_kernel).  Input is x2 = 2q (the main loop
// gets it for free by doubling the Horner coefficients: scaling by 2 is exact).  The division by s-1 is the
// exact 3-instruction sequence, and the final fl(fl(w*0.5)*(s-1)) is folded into one multiply by
// (s-1)/2, which is exact because w*0.5 is exact and (s-1)/2 is representable.
__device__ __forceinline__ float sample_coord_fast(float x2, float sm1, float rcp, float half_sm1) {
  const float q0 = __fmul_rn(x2, rcp);
  const float r = __fmaf_rn(-q0, sm1, x2);
  const float v = __fmaf_rn(r, rcp, q0);
  const float u = __fsub_rn(v, 1.f);
  return __fmul_rn(__fadd_rn(u, 1.f), half_sm1);
}

// floor + fraction of a coordinate already clamped to [-2, s]; table index domain is i in [-2, s].
__device__ __forceinline__ void split_coord(float ix, int s, int& i, float& f) {
  float c = fminf(fmaxf(ix, -2.f), (float)s);
  float fl = floorf(c);
  i = (int)fl;
  f = c - fl;
}

// Conservative window of table-entry indices touched by the inclusive voxel box, one axis.
// Bit-exact twin of oracle.dnmf_oracle.tile_window (fp32 interval arithmetic, fixed order).
// `b` points at beta[a*3 + d] entries of this frame with stride `bs` between rows a.
// `clipped` reports whether the clamp to the table domain [-2, s] cut the window: when it did not, every
// sample of the box is known to fall inside [wlo, whi] (the +-1 margin dominates the rounding of the
// interval sums and of the coordinate chain), so the consumer may index the staged slices unclamped.
__device__ __forceinline__ void tile_window_axis(const float* b, int bs, float x0, float y0, float z0,
                                                 float x1, float y1, float z1, int s, int& wlo,
                                                 int& whi, bool& clipped) {
  const float mlo[kBasis] = {1.f, x0, y0, z0, __fmul_rn(x0, x0), __fmul_rn(y0, y0), __fmul_rn(z0, z0),
                             __fmul_rn(x0, y0), __fmul_rn(x0, z0), __fmul_rn(y0, z0)};
  const float mhi[kBasis] = {1.f, x1, y1, z1, __fmul_rn(x1, x1), __fmul_rn(y1, y1), __fmul_rn(z1, z1),
                             __fmul_rn(x1, y1), __fmul_rn(x1, z1), __fmul_rn(y1, z1)};
  float lo = b[0], hi = b[0], mag = fabsf(b[0]);
#pragma unroll
  for (int a = 1; a < kBasis; ++a) {
    float c = b[a * bs];
    float p1 = __fmul_rn(c, mlo[a]);
    float p2 = __fmul_rn(c, mhi[a]);
    lo = __fadd_rn(lo, fminf(p1, p2));
    hi = __fadd_rn(hi, fmaxf(p1, p2));
    mag = fmaxf(mag, fmaxf(fabsf(p1), fabsf(p2)));
  }
  lo = fminf(fmaxf(lo, -4.f), (float)(s + 4));
  hi = fminf(fmaxf(hi, -4.f), (float)(s + 4));
  int l = (int)floorf(lo) - 1;
  int h = (int)floorf(hi) + 1;
  wlo = min(max(l, -2), s);
  whi = min(max(h, -2), s);
  // terms beyond 2^14 could carry a summed rounding error that is not small against the +-1 margin
  clipped = (l < -2) || (h > s) || !(lo <= hi) || !(mag <= 16384.f);
}

__device__ __forceinline__ void tile_window_axis(const float* b, int bs, float x0, float y0, float z0,
                                                 float x1, float y1, float z1, int s, int& wlo,
                                                 int& whi) {
  bool clipped;
  tile_window_axis(b, bs, x0, y0, z0, x1, y1, z1, s, wlo, whi, clipped);
}

// Neuron k (ranges r[0..5] = lo0,hi0,lo1,hi1,lo2,hi2) touches the window?
__device__ __forceinline__ bool neuron_in_window(const int* __restrict__ r, const int* wlo, const int* whi) {
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int lo = r[2 * d], hi = r[2 * d + 1];
    ok = ok && (lo <= hi) && (lo <= whi[d] + 1) && (hi >= wlo[d]);
  }
  return ok;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dnmf
