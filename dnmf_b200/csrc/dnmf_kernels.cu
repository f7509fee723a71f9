// dnmf_b200 -- hand-written sm_100a kernels of the dNMF fit hot path and their C ABI.
//
//   kernel 1a  build_tables_kernel      per-axis truncated Gaussian tables + integer ranges
//   kernel 1b  bin_count/scan/fill      deterministic neuron-to-tile binning (stand-alone form;
//                                       the fused kernel runs the same device code in its prologue)
//   kernel 2   fit_tile_kernel          fused forward + residual + loss + analytic beta-gradient
//              reduce_partials_kernel   fixed-order second-stage reduction (bit-reproducible)
//   kernel 3a  adam_kernel              dense Adam over all 30*T deformation coefficients
//   kernel 3b  mu_* kernels             trace statistics + multiplicative non-negative sweeps
//
// Reference lines (Demix/dNMF.py, demo.py) are cited next to each piece; math in SURVEY.md App. A.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dnmf_b200.h"
#include "dnmf_device.cuh"

#ifndef DNMF_ZUNROLL
#define DNMF_ZUNROLL 1  // z steps interleaved per lane in the fused kernel's main loop
#endif
namespace dnmf {
constexpr int kZUnroll = DNMF_ZUNROLL;
}
#ifndef DNMF_UNROLLED_MARCH
#define DNMF_UNROLLED_MARCH 0  // 1: one fully unrolled main loop per slot-pair count (more code than the I-cache holds)
#endif
#ifndef DNMF_MERGE_TAIL01
#define DNMF_MERGE_TAIL01 1  // 1: the specialised main loops keep "single slot" apart and merge even / odd lists
                             // (4 bodies instead of 6; measured at cfg2: identity beta 2.65 ms either way, a different
                             // deformation per frame 3.40 -> 2.94 ms per 1000 frames)
#endif
#ifndef DNMF_ALWAYS_SAFE
#define DNMF_ALWAYS_SAFE 0  // 1: every tile takes the clamped main loop (one loop body fewer in the instruction cache)
#endif
#ifndef DNMF_MU_MINB
#define DNMF_MU_MINB 10  // the same for the trace-statistics variant (MODE 3) of the single-warp layout: 167 registers
                         // (measured at cfg2, ms per 1000 frames: 16 -> 5.08, 14 -> 5.12 (128 regs, spills), 12 -> 4.66, 10 -> 4.35)
#endif
#ifndef DNMF_MINB
#define DNMF_MINB 16  // resident single-warp CTAs per SM the fused kernel is compiled for (124 registers used; 18 / 20 CTAs
                      // per SM compile to 94 registers without spills but measured 2 % slower at cfg2: 3.67e5 vs 3.75e5)
#endif

namespace dnmf {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& s) {
  g_err = s;
  return 1;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +    \
                  std::to_string(__LINE__) + ")");                                                \
  } while (0)

// ------------------------------------------------------------------------------------------------
// kernel 1a: tables.  entry(i) = (G[i], G[i+1]-G[i]) for i = -2..s, G = exp(-(i-pos)^2/sigma^2)
// inside [lo,hi], 0 outside (zero padding of grid_sample + cutoff).  Demix/dNMF.py:39-40.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gauss_node(int i, float pos, float sigma, int lo, int hi) {
  if (i < lo || i > hi) return 0.f;
  float d = __fsub_rn((float)i, pos);
  float q = __fdiv_rn(__fmul_rn(d, d), __fmul_rn(sigma, sigma));
  return expf(-q);
}

__global__ void build_ranges_kernel(const float* __restrict__ pos, const float* __restrict__ sigma,
                                    int K, int X, int Y, int Z, float cutoff, int* __restrict__ rng) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int sz[3] = {X, Y, Z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int s = sz[d];
    int lo = 0, hi = s - 1;
    if (cutoff > 0.f && isfinite(cutoff)) {
      float rad = __fmul_rn(cutoff, sigma[k]);
      float lo_f = ceilf(__fsub_rn(pos[k * 3 + d], rad));
      float hi_f = floorf(__fadd_rn(pos[k * 3 + d], rad));
      lo_f = fminf(fmaxf(lo_f, 0.f), (float)s);
      hi_f = fminf(fmaxf(hi_f, -1.f), (float)(s - 1));
      lo = (int)lo_f;
      hi = (int)hi_f;
    }
    rng[k * 6 + 2 * d] = lo;
    rng[k * 6 + 2 * d + 1] = hi;
  }
}

// d/dpos and d/dsigma of a node value (extension: learnable positions / widths); same truncation window
__device__ __forceinline__ void gauss_node_derivs(int i, float pos, float sigma, int lo, int hi, float& dpos,
                                                  float& dsig) {
  dpos = 0.f;
  dsig = 0.f;
  if (i < lo || i > hi) return;
  const float d = __fsub_rn((float)i, pos);
  const float s2 = __fmul_rn(sigma, sigma);
  const float g = expf(-__fdiv_rn(__fmul_rn(d, d), s2));
  dpos = g * 2.f * d / s2;
  dsig = g * 2.f * d * d / (s2 * sigma);
}

__global__ void build_tables_kernel(const float* __restrict__ pos, const float* __restrict__ sigma,
                                    const int* __restrict__ rng, int K, int s, int axis,
                                    float2* __restrict__ tab, float2* __restrict__ tab_dpos,
                                    float2* __restrict__ tab_dsig) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;  // entry within a neuron's row, i = e - 2
  int k = blockIdx.y;
  if (e >= s + 3) return;
  int i = e - 2;
  int lo = rng[k * 6 + 2 * axis], hi = rng[k * 6 + 2 * axis + 1];
  float p = pos[k * 3 + axis], sg = sigma[k];
  float g0 = gauss_node(i, p, sg, lo, hi);
  float g1 = gauss_node(i + 1, p, sg, lo, hi);
  tab[(size_t)k * (s + 3) + e] = make_float2(g0, __fsub_rn(g1, g0));
  if (tab_dpos != nullptr) {
    float p0, s0, p1, s1;
    gauss_node_derivs(i, p, sg, lo, hi, p0, s0);
    gauss_node_derivs(i + 1, p, sg, lo, hi, p1, s1);
    tab_dpos[(size_t)k * (s + 3) + e] = make_float2(p0, p1 - p0);
    tab_dsig[(size_t)k * (s + 3) + e] = make_float2(s0, s1 - s0);
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 1b: stand-alone binning (count -> exclusive scan -> fill), one warp per (frame, tile).
// ------------------------------------------------------------------------------------------------
struct Geom {
  int X, Y, Z, K, T;
  int tx, ty, tz, ntx, nty, ntz;
};

__device__ __forceinline__ void tile_box(const Geom& g, int tile, int& x0, int& y0, int& z0, int& x1,
                                         int& y1, int& z1) {
  int bx = tile % g.ntx;
  int by = (tile / g.ntx) % g.nty;
  int bz = tile / (g.ntx * g.nty);
  x0 = bx * g.tx;
  y0 = by * g.ty;
  z0 = bz * g.tz;
  x1 = min(x0 + g.tx, g.X) - 1;
  y1 = min(y0 + g.ty, g.Y) - 1;
  z1 = min(z0 + g.tz, g.Z) - 1;
}

template <bool FILL>
__global__ void bin_tiles_kernel(Geom g, const float* __restrict__ beta, const int* __restrict__ frame_ids,
                                 int B, const int* __restrict__ rng, int* __restrict__ counts,
                                 const long long* __restrict__ offsets, int* __restrict__ windows,
                                 int* __restrict__ ids, long long ids_capacity, int expand) {
  const int lane = threadIdx.x & 31;
  const int nt = g.ntx * g.nty * g.ntz;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (long long)B * nt) return;
  const int b = (int)(item / nt), tile = (int)(item - (long long)b * nt);
  const int t = frame_ids[b];
  int x0, y0, z0, x1, y1, z1;
  tile_box(g, tile, x0, y0, z0, x1, y1, z1);
  int wlo[3], whi[3];
  const int sz[3] = {g.X, g.Y, g.Z};
#pragma unroll
  for (int d = 0; d < 3; ++d)
    tile_window_axis(beta + (size_t)d * g.T + t, 3 * g.T, (float)x0, (float)y0, (float)z0, (float)x1,
                     (float)y1, (float)z1, sz[d], wlo[d], whi[d]);
#pragma unroll
  for (int d = 0; d < 3; ++d) {  // expand > 0 only when building the static candidate lists
    wlo[d] -= expand;
    whi[d] += expand;
  }
  long long base = FILL ? offsets[item] : 0;
  int cnt = 0;
  for (int k0 = 0; k0 < g.K; k0 += 32) {
    int k = k0 + lane;
    bool ok = (k < g.K) && neuron_in_window(rng + (size_t)k * 6, wlo, whi);
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (FILL && ok) {
      long long pos = base + cnt + __popc(m & ((1u << lane) - 1u));
      if (pos < ids_capacity) ids[pos] = k;
    }
    cnt += __popc(m);
  }
  if (!FILL && lane == 0) {
    counts[item] = cnt;
    if (windows) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        windows[item * 6 + 2 * d] = wlo[d];
        windows[item * 6 + 2 * d + 1] = whi[d];
      }
    }
  }
}

// Single-block exclusive scan of n int counts into n+1 int64 offsets; also the maximum count.
__global__ void scan_counts_kernel(const int* __restrict__ counts, long long n, long long* __restrict__ offsets,
                                   int* __restrict__ max_out) {
  __shared__ long long s_sum[1024];
  __shared__ int s_max[1024];
  const int tid = threadIdx.x, nth = blockDim.x;
  const long long per = (n + nth - 1) / nth;
  const long long i0 = min(n, per * tid), i1 = min(n, i0 + per);
  long long acc = 0;
  int mx = 0;
  for (long long i = i0; i < i1; ++i) {
    acc += counts[i];
    mx = max(mx, counts[i]);
  }
  s_sum[tid] = acc;
  s_max[tid] = mx;
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    int m = 0;
    for (int i = 0; i < nth; ++i) {
      long long v = s_sum[i];
      s_sum[i] = run;
      run += v;
      m = max(m, s_max[i]);
    }
    offsets[n] = run;
    if (max_out) *max_out = m;
  }
  __syncthreads();
  long long run = s_sum[tid];
  for (long long i = i0; i < i1; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: fused forward / residual / loss / analytic gradient.
//
// One CTA = NWX*NWY warps = one spatial tile (8*NWX) x (4*NWY) x tz of one frame.  The frame tile is
// read ONCE from HBM: one bulk async copy (TMA, cp.async.bulk + mbarrier) per contiguous run of
// ty*Z floats, landing in shared memory while the prologue runs.  Prologue: beta_t -> conservative
// sample window -> neuron list (ascending k, ballot compaction) -> the listed neurons' table slices
// staged in shared memory, entry-major [entry][slot] so that consecutive slots are adjacent (one
// LDS.128 serves two neurons, slot offsets are immediates) with C[k,t] folded into the x slice.
// Main loop: each lane owns one (x,y) column and marches along z; per (voxel, neuron) pair
// 1.5 LDS + 10 FP32 ops give Yhat and dYhat/dix (no N x K footprint matrix, no transcendental).
// The 30 gradient entries are accumulated as z-moments per lane, expanded with the lane's (x,y)
// monomials and reduced with a transposing butterfly (31 shuffles for 32 values) -> CTA partial;
// a second kernel sums partials in fixed order (bit-reproducible).
// Math: Demix/dNMF.py:54-58 + F.mse_loss (:188) + autograd of grid_sample wrt grid.
// ------------------------------------------------------------------------------------------------
struct FitParams {
  const float* frames;
  const int* frame_ids;
  const float* beta;
  const float* C;
  const float2* tab0;
  const float2* tab1;
  const float2* tab2;
  const int* rng;
  float* partials;
  float* yhat;   // MODE 1: Yhat output; MODE 2: residual output
  float bg;      // MODE 2: scalar background added to Yhat
  int frames_are_batch;
  int X, Y, Z, K, T;
  int tz, ntx, nty, ntz;
  int cap;  // staged slot capacity, even
  int wmax0, wmax1, wmax2;
  int full_depth;
  int bulk_ok;   // tile rows may be fetched with cp.async.bulk (16 B alignment holds)
  int fast_div;  // exact 3-instruction division verified for all three axes
  float rcp0, rcp1, rcp2;
  const long long* cand_off;  // static per-tile candidate lists (identity windows expanded by cand_expand)
  const int* cand_ids;
  int cand_expand;
  int cand_cap;  // shared-memory capacity for one tile's candidates (>= the longest static list when possible)
  int B;         // frames in this launch
  int fpc;       // consecutive frames walked by one CTA (<= 32)
  double* muG;   // MODE 3 (trace statistics): G_t[K][K], b_t[K] of every frame, accumulated with fp64 atomics
  double* mub;
  int* mu_overflow;  // MODE 3: set when a tile's list is not fully staged (the caller reruns the generic kernel)
  int dyn_tail;  // != 0: main loop with the run-time tail kind (one loop body per SAFE; see march_rolled TAIL 3)
  unsigned* restage_count;  // [32] frames whose slices were rebuilt, counted per CTA (MODE 0; may be NULL)
  int y_pitch;   // floats between x rows of the Y tile in shared memory (>= ty * tile depth)
  int z_skew;    // != 0: lane (lx, ly) starts its z march at ((ly * z_skew) & 3), see march_rolled<SKEW>
  int tmap_ok;   // the frame tile can be fetched with ONE tensor TMA copy (3-D map over [frame][x][y*Z])
  int b_base;    // index of this launch's first frame in the buffer the tensor map describes
  alignas(64) CUtensorMap tmap;
};

struct FitSmem {
  int tab_f2;    // float2 count of the staged-table region
  int y_f;       // float count of the Y tile
  int list_u16;  // uint16 count of the list
  size_t bytes;
};

static FitSmem fit_smem_layout(int nw, int tx, int ty, int tz, int cap, int wsum, int K, int wmax0, int cand_cap,
                               int y_pitch) {
  FitSmem s;
  s.tab_f2 = cap * (wsum + wmax0);  // live slices + the x slice without traces
  s.y_f = tx * std::max(y_pitch, ty * tz) + 4;
  s.list_u16 = (K + 7) & ~7;
  s.bytes = (((size_t)s.tab_f2 * 8 + 127) & ~(size_t)127) + (size_t)s.y_f * 4 + (size_t)nw * kNumPartials * 4 +
            80 * 4 + 16 +
            (size_t)((cap + 5) & ~3) * 4 + (size_t)cand_cap * 28 + (size_t)((cand_cap + 7) & ~7) * 2 + (size_t)((cap + 7) & ~7) * 2 +
            (size_t)s.list_u16 * 2;
  return s;
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

// Exhaustive check over all 2^32 float bit patterns that the fast form equals the reference op
// sequence (true division) for this axis size; the fast path is enabled only when no pattern differs.
__global__ void verify_coord_kernel(float sm1, float rcp, unsigned long long* __restrict__ mismatches) {
  const unsigned base = (blockIdx.x * blockDim.x + threadIdx.x) * 256u;
  unsigned bad = 0;
  for (unsigned i = 0; i < 256u; ++i) {
    const unsigned bits = base + i;
    if (((bits >> 23) & 0xffu) >= 253u) continue;  // |q| >= 2^126, inf, NaN: 2q overflows, reference is UB there
    const float q = __uint_as_float(bits);
    const float a = sample_coord(q, sm1);
    const float b = sample_coord_fast(__fadd_rn(q, q), sm1, rcp, __fmul_rn(0.5f, sm1));
    const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (isnan(a) && isnan(b));
    bad += same ? 0u : 1u;
  }
  if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

__device__ __forceinline__ void pair_accumulate(float ex_g, float ex_d, float ey_g, float ey_d, float ez_g,
                                                float ez_d, float f0, float f1, float f2, float& yh,
                                                float& g0, float& g1, float& g2) {
  const float ca0 = fmaf(f0, ex_d, ex_g);
  const float a1 = fmaf(f1, ey_d, ey_g);
  const float a2 = fmaf(f2, ez_d, ez_g);
  const float t12 = a1 * a2;
  yh = fmaf(ca0, t12, yh);
  g0 = fmaf(ex_d, t12, g0);
  g1 = fmaf(ca0 * a2, ey_d, g1);
  g2 = fmaf(ca0 * a1, ez_d, g2);
}

// Two neurons at once with Blackwell's packed FP32x2 instructions (FFMA2 / FMUL2: two IEEE fp32
// results per lane per issue slot).  Operands are (slot j, slot j+1) pairs straight out of one LDS.128.
__device__ __forceinline__ void pair2_accumulate(float2 exG, float2 exD, float2 eyG, float2 eyD, float2 ezG,
                                                 float2 ezD, float2 f0, float2 f1, float2 f2, float2& yh,
                                                 float2& g0, float2& g1, float2& g2) {
  const float2 ca0 = __ffma2_rn(f0, exD, exG);
  const float2 a1 = __ffma2_rn(f1, eyD, eyG);
  const float2 a2 = __ffma2_rn(f2, ezD, ezG);
  const float2 t12 = __fmul2_rn(a1, a2);
  yh = __ffma2_rn(ca0, t12, yh);
  g0 = __ffma2_rn(exD, t12, g0);
  g1 = __ffma2_rn(__fmul2_rn(ca0, a2), eyD, g1);
  g2 = __ffma2_rn(__fmul2_rn(ca0, a1), ezD, g2);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Loop-invariant values the compiler would otherwise rematerialise inside the hot loop (constant-bank
// reloads, int->float conversions, address arithmetic): routing them through an opaque move pins
// them in a register.
__device__ __forceinline__ float pin(float v) {
  float r;
  asm volatile("mov.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ unsigned pin(unsigned v) {
  unsigned r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ int pin(int v) {
  int r;
  asm volatile("mov.s32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// Sum 32 per-lane values across the warp with 31 shuffles: afterwards lane l holds the total of v[l].
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------
// Specialised main loop of the fused kernel for the common case (two y-adjacent sub-tiles per warp, the
// whole list staged): each lane carries voxel A = (x, y, z) and voxel B = (x, y+4, z) through the z march
// TOGETHER, so all per-voxel arithmetic that is not neuron-pair math (Horner q, the F2 coordinate chain,
// fraction, residual, loss, the nine gradient z-moments) runs as packed FP32x2 over (A, B), while the
// neuron-pair math stays packed over (slot j, slot j+1).  The number of staged slot pairs NP is a template
// parameter: the pair loop is fully unrolled with immediate LDS offsets (no loop counter, no address
// increments, no zeroed accumulators), selected per tile by a uniform switch.  SAFE = false additionally
// drops the window clamp and the out-of-volume lane mask; it is chosen only for full tiles whose
// conservative window was not clipped by the table domain (tile_window_axis), where every sample is known to
// index inside the staged slices.  Same IEEE operations per value as the generic loop below.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxNP = 8;  // unrolled variants up to 16 staged neurons; longer lists take the generic loop

struct MarchArgs {
  float2 c0[3], c1[3];      // Horner coefficients of 2q for (A, B), per axis
  float c2[3];              // z^2 coefficient (shared by A and B)
  float sm1[3], rcp[3], hsm1[3];
  int wl[3], wm1[3];
  unsigned base[3];         // shared-memory byte address of each axis' slice region
  unsigned strideB;         // bytes per table entry (CAP slots of 8 B)
  unsigned yaddrA, yoffB;   // byte address of A's Y column; B's column is yoffB bytes further
  float zf0;
  int nz;
  bool validA, validB;
  float bg;                 // MODE 2: scalar background
  float oz;                 // 0.0f the compiler cannot see (loaded from shared memory, sK[11])
  int zskew;                // SKEW variants: this lane starts its z march at z0 + zskew and wraps around
};

struct MarchOut {
  float2 S0[3], S1[3], S2[3];  // z-moments of r * dYhat/dix_d for (A, B)
  float2 sse, sum_r;
};

__device__ __forceinline__ float4 lds128r(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int OFF>
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr), "n"(OFF));
  return v;
}

// One slot pair (two neurons) for one voxel.  FIRST: the accumulators are produced, not updated.
template <int OFF, bool FIRST>
__device__ __forceinline__ void slot_pair(unsigned ax, unsigned ay, unsigned az, float2 f0, float2 f1, float2 f2,
                                          float2& yh, float2& g0, float2& g1, float2& g2) {
  const float4 ex = lds128<OFF>(ax), ey = lds128<OFF>(ay), ez = lds128<OFF>(az);
  const float2 exG = make_float2(ex.x, ex.y), exD = make_float2(ex.z, ex.w);
  const float2 eyG = make_float2(ey.x, ey.y), eyD = make_float2(ey.z, ey.w);
  const float2 ezG = make_float2(ez.x, ez.y), ezD = make_float2(ez.z, ez.w);
  const float2 ca0 = __ffma2_rn(f0, exD, exG);
  const float2 a1 = __ffma2_rn(f1, eyD, eyG);
  const float2 a2 = __ffma2_rn(f2, ezD, ezG);
  const float2 t12 = __fmul2_rn(a1, a2);
  if (FIRST) {
    yh = __fmul2_rn(ca0, t12);
    g0 = __fmul2_rn(exD, t12);
    g1 = __fmul2_rn(__fmul2_rn(ca0, a2), eyD);
    g2 = __fmul2_rn(__fmul2_rn(ca0, a1), ezD);
  } else {
    yh = __ffma2_rn(ca0, t12, yh);
    g0 = __ffma2_rn(exD, t12, g0);
    g1 = __ffma2_rn(__fmul2_rn(ca0, a2), eyD, g1);
    g2 = __ffma2_rn(__fmul2_rn(ca0, a1), ezD, g2);
  }
}

// All NP slot pairs of one voxel -> (Yhat, dYhat/dix_0..2).
template <int NP>
__device__ __forceinline__ void voxel_pairs(unsigned ax, unsigned ay, unsigned az, float f0s, float f1s, float f2s,
                                            float& yh, float& g0, float& g1, float& g2) {
  static_assert(NP >= 1 && NP <= 8, "unrolled slot pairs");
  const float2 f0 = make_float2(f0s, f0s), f1 = make_float2(f1s, f1s), f2 = make_float2(f2s, f2s);
  float2 y2, a2, b2, c2;
  slot_pair<0, true>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 1) slot_pair<16, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 2) slot_pair<32, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 3) slot_pair<48, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 4) slot_pair<64, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 5) slot_pair<80, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 6) slot_pair<96, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  if (NP > 7) slot_pair<112, false>(ax, ay, az, f0, f1, f2, y2, a2, b2, c2);
  yh = y2.x + y2.y;
  g0 = a2.x + a2.y;
  g1 = b2.x + b2.y;
  g2 = c2.x + c2.y;
}

__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// MODE as in fit_tile_kernel: 0 fit, 1 forward only (Yhat replaces the Y tile in shared memory), 2 fit with
// scalar background, residual written back to the Y tile.
template <int NP, bool SAFE, int MODE>
__device__ __forceinline__ void march_pairs(const MarchArgs& a, MarchOut& o) {
  // accumulators start from an opaque zero (read back from shared memory): with a literal 0 ptxas peels the first z step
  // into a straight-line copy of the whole loop body (fold of 0 + x), which costs instruction-cache footprint
  const float oz = a.oz;
  const float2 zero2 = make_float2(oz, oz);
#pragma unroll
  for (int d = 0; d < 3; ++d) o.S0[d] = o.S1[d] = o.S2[d] = zero2;
  float2 sse = zero2, sum_r = zero2;
  unsigned yaddr = a.yaddrA;
  if constexpr (NP == 0) {  // empty list: Yhat = 0, no gradient; only the loss term
#pragma unroll 1
    for (int zz = 0; zz < a.nz; ++zz, yaddr += 4u) {
      if (MODE == 1) {
        sts32(yaddr, 0.f);
        sts32(yaddr + a.yoffB, 0.f);
        continue;
      }
      float2 r = make_float2(-lds32(yaddr), -lds32(yaddr + a.yoffB));
      if (MODE == 2) r = __fadd2_rn(r, make_float2(a.bg, a.bg));
      if (SAFE) {
        r.x = a.validA ? r.x : 0.f;
        r.y = a.validB ? r.y : 0.f;
      }
      if (MODE == 2) {
        sts32(yaddr, r.x);
        sts32(yaddr + a.yoffB, r.y);
        sum_r = __fadd2_rn(sum_r, r);
      }
      sse = __ffma2_rn(r, r, sse);
    }
  } else {
    // address of entry i of axis d: (i - wl) * strideB + base  ==  i * strideB + bias   (mod 2^32)
    unsigned bias[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) bias[d] = a.base[d] - (unsigned)a.wl[d] * a.strideB;
    float zf = a.zf0;
#pragma unroll 1
    for (int zz = 0; zz < a.nz; ++zz, zf += 1.f, yaddr += 4u) {
      const float2 z2 = make_float2(zf, zf);
      float2 ix[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float2 rcp2 = make_float2(a.rcp[d], a.rcp[d]);
        const float2 q = __ffma2_rn(z2, __ffma2_rn(z2, make_float2(a.c2[d], a.c2[d]), a.c1[d]), a.c0[d]);  // = 2q
        const float2 t0 = __fmul2_rn(q, rcp2);
        const float2 r = __ffma2_rn(make_float2(-t0.x, -t0.y), make_float2(a.sm1[d], a.sm1[d]), q);
        const float2 v = __ffma2_rn(r, rcp2, t0);  // = fl(2q / (s-1)), verified exact (verify_coord_kernel)
        const float2 u = __fadd2_rn(v, make_float2(-1.f, -1.f));
        ix[d] = __fmul2_rn(__fadd2_rn(u, make_float2(1.f, 1.f)), make_float2(a.hsm1[d], a.hsm1[d]));
      }
      unsigned adA[3], adB[3];
      float2 f[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int iA = __float2int_rd(ix[d].x), iB = __float2int_rd(ix[d].y);
        f[d] = __fadd2_rn(ix[d], make_float2(-(float)iA, -(float)iB));
        if (SAFE) {
          adA[d] = (unsigned)min(max(iA - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
          adB[d] = (unsigned)min(max(iB - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
        } else {
          adA[d] = (unsigned)iA * a.strideB + bias[d];
          adB[d] = (unsigned)iB * a.strideB + bias[d];
        }
      }
      float2 yh, g[3];
      voxel_pairs<NP>(adA[0], adA[1], adA[2], f[0].x, f[1].x, f[2].x, yh.x, g[0].x, g[1].x, g[2].x);
      voxel_pairs<NP>(adB[0], adB[1], adB[2], f[0].y, f[1].y, f[2].y, yh.y, g[0].y, g[1].y, g[2].y);
      if (MODE == 1) {
        sts32(yaddr, yh.x);
        sts32(yaddr + a.yoffB, yh.y);
        continue;
      }
      if (MODE == 2) yh = __fadd2_rn(yh, make_float2(a.bg, a.bg));
      float2 r = __fadd2_rn(yh, make_float2(-lds32(yaddr), -lds32(yaddr + a.yoffB)));
      if (SAFE) {
        r.x = a.validA ? r.x : 0.f;
        r.y = a.validB ? r.y : 0.f;
      }
      if (MODE == 2) {
        sts32(yaddr, r.x);
        sts32(yaddr + a.yoffB, r.y);
        sum_r = __fadd2_rn(sum_r, r);
      }
      sse = __ffma2_rn(r, r, sse);
      const float zq = zf * zf;
      const float2 zq2 = make_float2(zq, zq);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float2 h = __fmul2_rn(r, g[d]);
        o.S0[d] = __fadd2_rn(o.S0[d], h);
        o.S1[d] = __ffma2_rn(z2, h, o.S1[d]);
        o.S2[d] = __ffma2_rn(zq2, h, o.S2[d]);
      }
    }
  }
  o.sse = sse;
  o.sum_r = sum_r;
}

// Rolled form of march_pairs: the slot-pair count is a run-time (tile-uniform) value and the pair loop is a
// real loop, A and B interleaved inside it.  A few more integer/branch instructions per pair than the
// unrolled variants, but ONE loop body for every list length: the kernel's hot code then fits the 32 KB
// L1.5 instruction cache (the unrolled family did not, and the SMs starved on instruction fetch).

// Last slot of an odd list, for voxels A and B at once: the slot's (G, D) entries of A and of B are loaded as
// scalars into adjacent registers, so the ten operations run packed over (A, B) and land directly on the
// (A, B)-packed Yhat / gradient values.  Half the FP32-pipe cycles of a zero-padded slot pair.
// FIRST: Yhat is produced as fma(ca0, t12, oz) with the opaque zero `oz`, not as a product: ptxas 12.9 fuses
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with --fmad=false), which would round Yhat - Y differently from
// the forward-only instantiation that stores Yhat, and break "fit of the model's own output has zero residual".
template <bool FIRST>
__device__ __forceinline__ void tail_slot(const unsigned (&adA)[3], const unsigned (&adB)[3], unsigned off,
                                          const float2 (&f)[3], float oz, float2& yh, float2 (&g)[3]) {
  const float2 exG = make_float2(lds32(adA[0] + off), lds32(adB[0] + off));
  const float2 exD = make_float2(lds32(adA[0] + off + 8u), lds32(adB[0] + off + 8u));
  const float2 eyG = make_float2(lds32(adA[1] + off), lds32(adB[1] + off));
  const float2 eyD = make_float2(lds32(adA[1] + off + 8u), lds32(adB[1] + off + 8u));
  const float2 ezG = make_float2(lds32(adA[2] + off), lds32(adB[2] + off));
  const float2 ezD = make_float2(lds32(adA[2] + off + 8u), lds32(adB[2] + off + 8u));
  const float2 ca0 = __ffma2_rn(f[0], exD, exG);
  const float2 a1 = __ffma2_rn(f[1], eyD, eyG);
  const float2 a2 = __ffma2_rn(f[2], ezD, ezG);
  const float2 t12 = __fmul2_rn(a1, a2);
  if (FIRST) {
    yh = __ffma2_rn(ca0, t12, make_float2(oz, oz));
    g[0] = __fmul2_rn(exD, t12);
    g[1] = __fmul2_rn(__fmul2_rn(ca0, a2), eyD);
    g[2] = __fmul2_rn(__fmul2_rn(ca0, a1), ezD);
  } else {
    yh = __ffma2_rn(ca0, t12, yh);
    g[0] = __ffma2_rn(exD, t12, g[0]);
    g[1] = __ffma2_rn(__fmul2_rn(ca0, a2), eyD, g[1]);
    g[2] = __ffma2_rn(__fmul2_rn(ca0, a1), ezD, g[2]);
  }
}

// TAIL 0: even list, np >= 1 full slot pairs.  TAIL 1: np >= 1 full pairs and one last slot.  TAIL 2: a single
// slot (np == 0).  TAIL 3: the kind is the run-time value `tail` (warp-uniform branches inside the z loop): one
// loop body per SAFE instead of three.  TAIL 4: np >= 1 and a run-time choice between kinds 0 and 1 only.
// Why: with three specialised bodies per SAFE, once every frame of a CTA has its own deformation the per-frame
// list / restage code joins the hot set, the 32 KB instruction cache thrashes (stall_no_instruction 1.7 per issue,
// profiles/README.md) and the kernel loses 20-25 %.  Measured at cfg2, ms per 1000 frames, identity beta / a
// different deformation per frame: TAIL {0,1,2} 2.65 / 3.40, TAIL 3 2.77 / 3.05, TAIL {4,2} 2.65 / 2.94 (default,
// DNMF_MERGE_TAIL01).  FitParams::dyn_tail selects TAIL 3 per launch (DNMF_DYN_TAIL=1, or -1: from the counters).
// SKEW: the lanes of a warp walk z in rotated order (lane-dependent start, wrap-around), which spreads their
// reads of the Y tile over the banks when the tile's y/x pitches are multiples of 32 floats (Z = 32).
template <bool SAFE, int MODE, int TAIL, bool SKEW>
__device__ __forceinline__ void march_rolled(const MarchArgs& a, int np, MarchOut& o, int tail = 0) {
  const float oz = a.oz;
  const float2 zero2 = make_float2(oz, oz);
#pragma unroll
  for (int d = 0; d < 3; ++d) o.S0[d] = o.S1[d] = o.S2[d] = zero2;
  float2 sse = zero2, sum_r = zero2;
  unsigned bias[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) bias[d] = a.base[d] - (unsigned)a.wl[d] * a.strideB;
  const unsigned pair_bytes = (unsigned)np * 16u;
  int zi = SKEW ? a.zskew : 0;
  float zf = a.zf0 + (float)zi;
  unsigned yaddr = a.yaddrA + 4u * (unsigned)zi;
  const float nzf = (float)a.nz;
  const unsigned nz4 = 4u * (unsigned)a.nz;
  auto advance = [&]() {
    zf += 1.f;
    yaddr += 4u;
    if (SKEW) {
      if (++zi == a.nz) {  // wrap with uniform decrements: per-lane start values would be rematerialised in the loop
        zi = 0;
        zf -= nzf;
        yaddr -= nz4;
      }
    }
  };
#pragma unroll 1
  for (int zz = 0; zz < a.nz; ++zz, advance()) {
    const float2 z2 = make_float2(zf, zf);
    float2 ix[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float2 rcp2 = make_float2(a.rcp[d], a.rcp[d]);
      const float2 q = __ffma2_rn(z2, __ffma2_rn(z2, make_float2(a.c2[d], a.c2[d]), a.c1[d]), a.c0[d]);  // = 2q
      const float2 t0 = __fmul2_rn(q, rcp2);
      const float2 r = __ffma2_rn(make_float2(-t0.x, -t0.y), make_float2(a.sm1[d], a.sm1[d]), q);
      const float2 v = __ffma2_rn(r, rcp2, t0);  // = fl(2q / (s-1)), verified exact (verify_coord_kernel)
      const float2 u = __fadd2_rn(v, make_float2(-1.f, -1.f));
      ix[d] = __fmul2_rn(__fadd2_rn(u, make_float2(1.f, 1.f)), make_float2(a.hsm1[d], a.hsm1[d]));
    }
    unsigned adA[3], adB[3];
    float2 f[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int iA = __float2int_rd(ix[d].x), iB = __float2int_rd(ix[d].y);
      f[d] = __fadd2_rn(ix[d], make_float2(-(float)iA, -(float)iB));
      if (SAFE) {
        adA[d] = (unsigned)min(max(iA - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
        adB[d] = (unsigned)min(max(iB - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
      } else {
        adA[d] = (unsigned)iA * a.strideB + bias[d];
        adB[d] = (unsigned)iB * a.strideB + bias[d];
      }
    }
    const float2 fA0 = make_float2(f[0].x, f[0].x), fA1 = make_float2(f[1].x, f[1].x), fA2 = make_float2(f[2].x, f[2].x);
    const float2 fB0 = make_float2(f[0].y, f[0].y), fB1 = make_float2(f[1].y, f[1].y), fB2 = make_float2(f[2].y, f[2].y);
    float2 yh, g[3];
    if (TAIL == 3 ? (tail != 2) : (TAIL != 2)) {  // TAIL 4: np >= 1, run-time choice between kinds 0 and 1
      // first slot pair produces the accumulators, the rest of the list updates them
      float2 yA, gA0, gA1, gA2, yB, gB0, gB1, gB2;
      slot_pair<0, true>(adA[0], adA[1], adA[2], fA0, fA1, fA2, yA, gA0, gA1, gA2);
      slot_pair<0, true>(adB[0], adB[1], adB[2], fB0, fB1, fB2, yB, gB0, gB1, gB2);
#pragma unroll 1
      for (unsigned off = 16u; off < pair_bytes; off += 16u) {
        {
          const float4 ex = lds128r(adA[0] + off), ey = lds128r(adA[1] + off), ez = lds128r(adA[2] + off);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(ey.x, ey.y),
                           make_float2(ey.z, ey.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fA0, fA1, fA2,
                           yA, gA0, gA1, gA2);
        }
        {
          const float4 ex = lds128r(adB[0] + off), ey = lds128r(adB[1] + off), ez = lds128r(adB[2] + off);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(ey.x, ey.y),
                           make_float2(ey.z, ey.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fB0, fB1, fB2,
                           yB, gB0, gB1, gB2);
        }
      }
      yh = make_float2(yA.x + yA.y, yB.x + yB.y);
      g[0] = make_float2(gA0.x + gA0.y, gB0.x + gB0.y);
      g[1] = make_float2(gA1.x + gA1.y, gB1.x + gB1.y);
      g[2] = make_float2(gA2.x + gA2.y, gB2.x + gB2.y);
      if ((TAIL == 3 || TAIL == 4) ? (tail == 1) : (TAIL == 1)) tail_slot<false>(adA, adB, pair_bytes, f, oz, yh, g);
    } else {
      tail_slot<true>(adA, adB, 0u, f, oz, yh, g);
    }
    if (MODE == 1) {
      sts32(yaddr, yh.x);
      sts32(yaddr + a.yoffB, yh.y);
      continue;
    }
    if (MODE == 2) yh = __fadd2_rn(yh, make_float2(a.bg, a.bg));
    float2 r = __fadd2_rn(yh, make_float2(-lds32(yaddr), -lds32(yaddr + a.yoffB)));
    if (SAFE) {
      r.x = a.validA ? r.x : 0.f;
      r.y = a.validB ? r.y : 0.f;
    }
    if (MODE == 2) {
      sts32(yaddr, r.x);
      sts32(yaddr + a.yoffB, r.y);
      sum_r = __fadd2_rn(sum_r, r);
    }
    sse = __ffma2_rn(r, r, sse);
    const float zq = zf * zf;
    const float2 zq2 = make_float2(zq, zq);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float2 h = __fmul2_rn(r, g[d]);
      o.S0[d] = __fadd2_rn(o.S0[d], h);
      o.S1[d] = __ffma2_rn(z2, h, o.S1[d]);
      o.S2[d] = __ffma2_rn(zq2, h, o.S2[d]);
    }
  }
  o.sse = sse;
  o.sum_r = sum_r;
}

// ------------------------------------------------------------------------------------------------
// MODE 3: trace statistics G_t = A_t^T A_t, b_t = A_t^T Y_t (Demix/dNMF.py:141-142) on the fused kernel's
// machinery (same tiles, lists, staged slices WITHOUT the traces, packed (A, B) coordinate chain).  Per voxel the
// footprint values of NP "row" slot pairs are formed packed over (slot 2p, slot 2p+1); every column slot value
// is broadcast against them: acc[p][c] += (a_2p, a_2p+1) * a_c is one FFMA2 per voxel.  Accumulators stay in
// registers over the tile-frame: a block of 3 row pairs x 6 column slots at a time (longer lists make several
// passes over the tile, one per block pair on or above the diagonal).  Flushed per tile-frame with the
// transposing warp reduction and fp64 atomics.
// ------------------------------------------------------------------------------------------------
// footprint values of slots (2p, 2p+1) at one voxel from the three staged slices
__device__ __forceinline__ float2 slot_values(unsigned ax, unsigned ay, unsigned az, unsigned off, float f0, float f1,
                                              float f2) {
  const float4 ex = lds128r(ax + off), ey = lds128r(ay + off), ez = lds128r(az + off);
  const float2 a0 = __ffma2_rn(make_float2(f0, f0), make_float2(ex.z, ex.w), make_float2(ex.x, ex.y));
  const float2 a1 = __ffma2_rn(make_float2(f1, f1), make_float2(ey.z, ey.w), make_float2(ey.x, ey.y));
  const float2 a2 = __ffma2_rn(make_float2(f2, f2), make_float2(ez.z, ez.w), make_float2(ez.x, ez.y));
  return __fmul2_rn(__fmul2_rn(a0, a1), a2);
}

// NP row pairs starting at byte offset rowoff (rowp of them real), and -- TWO -- three column pairs at coloff
// (colp real); without TWO the columns are the rows.  G[p][c]: (row slots 2p, 2p+1) x column slot c.
template <int NP, bool TWO>
__device__ __forceinline__ void march_stats(const MarchArgs& a, unsigned rowoff, unsigned coloff, int rowp, int colp,
                                            float2 (&G)[3][6], float2 (&bv)[3]) {
  constexpr int NC = TWO ? 3 : NP;  // column pairs
  int zi = a.zskew;
  float zf = a.zf0 + (float)zi;
  unsigned yaddr = a.yaddrA + 4u * (unsigned)zi;
  const float nzf = (float)a.nz;
  const unsigned nz4 = 4u * (unsigned)a.nz;
  unsigned roff[NP], coff[NC];
  float rw[NP], cw[NC];  // 0 for padding pairs (they re-read a real pair)
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    roff[i] = rowoff + 16u * (unsigned)min(i, rowp - 1);
    rw[i] = i < rowp ? 1.f : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    coff[i] = coloff + 16u * (unsigned)min(i, colp - 1);
    cw[i] = i < colp ? 1.f : 0.f;
  }
#pragma unroll 1
  for (int zz = 0; zz < a.nz; ++zz) {
    const float2 z2 = make_float2(zf, zf);
    float2 ix[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float2 rcp2 = make_float2(a.rcp[d], a.rcp[d]);
      const float2 q = __ffma2_rn(z2, __ffma2_rn(z2, make_float2(a.c2[d], a.c2[d]), a.c1[d]), a.c0[d]);  // = 2q
      const float2 t0 = __fmul2_rn(q, rcp2);
      const float2 r = __ffma2_rn(make_float2(-t0.x, -t0.y), make_float2(a.sm1[d], a.sm1[d]), q);
      const float2 v = __ffma2_rn(r, rcp2, t0);
      const float2 u = __fadd2_rn(v, make_float2(-1.f, -1.f));
      ix[d] = __fmul2_rn(__fadd2_rn(u, make_float2(1.f, 1.f)), make_float2(a.hsm1[d], a.hsm1[d]));
    }
    unsigned adA[3], adB[3];
    float2 f[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int iA = __float2int_rd(ix[d].x), iB = __float2int_rd(ix[d].y);
      f[d] = __fadd2_rn(ix[d], make_float2(-(float)iA, -(float)iB));
      adA[d] = (unsigned)min(max(iA - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
      adB[d] = (unsigned)min(max(iB - a.wl[d], 0), a.wm1[d]) * a.strideB + a.base[d];
    }
    const float yA = a.validA ? lds32(yaddr) : 0.f, yB = a.validB ? lds32(yaddr + a.yoffB) : 0.f;
    const float mA = a.validA ? 1.f : 0.f, mB = a.validB ? 1.f : 0.f;
    float2 rA[NP], rB[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 vA = slot_values(adA[0], adA[1], adA[2], roff[i], f[0].x, f[1].x, f[2].x);
      const float2 vB = slot_values(adB[0], adB[1], adB[2], roff[i], f[0].y, f[1].y, f[2].y);
      rA[i] = __fmul2_rn(vA, make_float2(rw[i] * mA, rw[i] * mA));
      rB[i] = __fmul2_rn(vB, make_float2(rw[i] * mB, rw[i] * mB));
      bv[i] = __ffma2_rn(rA[i], make_float2(yA, yA), bv[i]);
      bv[i] = __ffma2_rn(rB[i], make_float2(yB, yB), bv[i]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float2 cA, cB;
      if (TWO) {
        cA = slot_values(adA[0], adA[1], adA[2], coff[c], f[0].x, f[1].x, f[2].x);
        cB = slot_values(adB[0], adB[1], adB[2], coff[c], f[0].y, f[1].y, f[2].y);
        cA = __fmul2_rn(cA, make_float2(cw[c], cw[c]));
        cB = __fmul2_rn(cB, make_float2(cw[c], cw[c]));
      } else {
        cA = rA[c];
        cB = rB[c];
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        G[i][2 * c] = __ffma2_rn(rA[i], make_float2(cA.x, cA.x), G[i][2 * c]);
        G[i][2 * c] = __ffma2_rn(rB[i], make_float2(cB.x, cB.x), G[i][2 * c]);
        G[i][2 * c + 1] = __ffma2_rn(rA[i], make_float2(cA.y, cA.y), G[i][2 * c + 1]);
        G[i][2 * c + 1] = __ffma2_rn(rB[i], make_float2(cB.y, cB.y), G[i][2 * c + 1]);
      }
    }
    zf += 1.f;
    yaddr += 4u;
    if (++zi == a.nz) {
      zi = 0;
      zf -= nzf;
      yaddr -= nz4;
    }
  }
}

// Warp-reduce the accumulators of one block pair and add them to G_t / b_t.  Row slots start at slot 6*pb,
// column slots at 6*lb; off-diagonal blocks are mirrored, b_t is accumulated on diagonal blocks only.
template <int NP, bool TWO>
__device__ __forceinline__ void flush_stats(const float2 (&G)[3][6], const float2 (&bv)[3], int pb, int lb, int nst,
                                            const unsigned short* sList, double* Gt, double* bt, int K, int lane) {
  constexpr int NCS = TWO ? 6 : 2 * NP;        // column slots
  constexpr int NG = 2 * NP * NCS;              // G outputs, index = (p*NCS + c)*2 + half
  constexpr int NOUT = NG + (TWO ? 0 : 2 * NP);  // + b outputs
#pragma unroll
  for (int r0 = 0; r0 < NOUT; r0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int idx = r0 + i;
      float x = 0.f;
      if (idx < NG) {
        const int pc = idx >> 1, pp = pc / NCS, cc = pc % NCS;
        x = (idx & 1) ? G[pp][cc].y : G[pp][cc].x;
      } else if (idx < NOUT) {
        const int q = idx - NG;
        x = (q & 1) ? bv[q >> 1].y : bv[q >> 1].x;
      }
      v[i] = x;
    }
    const float tot = warp_transpose_sum(v, lane);
    const int idx = r0 + lane;
    if (idx < NG) {
      const int pc = idx >> 1, pp = pc / NCS, cc = pc % NCS;
      const int j = 6 * pb + 2 * pp + (idx & 1), l = 6 * lb + cc;
      if (j < nst && l < nst) {
        const int kj = sList[j], kl = sList[l];
        atomicAdd(Gt + (size_t)kj * K + kl, (double)tot);
        if (TWO) atomicAdd(Gt + (size_t)kl * K + kj, (double)tot);
      }
    } else if (idx < NOUT) {
      const int q = idx - NG, j = 6 * pb + q;
      if (j < nst) atomicAdd(bt + sList[j], (double)tot);
    }
  }
}

// Generic main loop (any list length, overflow slots from the L2-resident tables, SUB = 1 or 2, true
// division when the fast form is not verified): one sub-tile after the other, slot pairs in a rolled loop.
struct GenericArgs {
  const FitParams* p;
  const float* sBeta;
  const unsigned short* sList;
  const float* sY;
  int t, L, nst;
  int x0, y0, z0, nz;
  int lx, ly0, RS, zs;
  int wl[3], wm1[3];
  unsigned base[3], strideB;
  float bg;
};

template <int SUB, int MODE, bool FAST_DIV>
__device__ __forceinline__ void march_generic(const GenericArgs& a, float (&S0)[SUB][3], float (&S1)[SUB][3],
                                              float (&S2)[3], float& sse, float& sum_r) {
  constexpr bool WRITE_YHAT = MODE == 1;
  constexpr bool WRITE_RES = MODE == 2;
  const FitParams& p = *a.p;
  const float* sBeta = a.sBeta;
  const int gx = a.x0 + a.lx;
  const float xf = (float)gx;
  const float sm1x = pin((float)(p.X - 1)), sm1y = pin((float)(p.Y - 1)), sm1z = pin((float)(p.Z - 1));
  const float rcpx = pin(p.rcp0), rcpy = pin(p.rcp1), rcpz = pin(p.rcp2);
  const float hsm1x = pin(0.5f * sm1x), hsm1y = pin(0.5f * sm1y), hsm1z = pin(0.5f * sm1z);
  const float2 sm1xy = make_float2(sm1x, sm1y), rcpxy = make_float2(rcpx, rcpy), hsm1xy = make_float2(hsm1x, hsm1y);
  const unsigned strideB = pin(a.strideB);
  const unsigned baseX = pin(a.base[0]), baseY = pin(a.base[1]), baseZ = pin(a.base[2]);
  const int W0m1 = pin(a.wm1[0]), W1m1 = pin(a.wm1[1]), W2m1 = pin(a.wm1[2]);
  const int wl0 = pin(a.wl[0]), wl1 = pin(a.wl[1]), wl2 = pin(a.wl[2]);
  const int nst = a.nst, L = a.L, t = a.t;
  const int nquad = nst >> 2;
  const bool has_overflow = L > nst;
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
  const float bg = a.bg;
  float2 S0xy[SUB], S1xy[SUB], S2xy = make_float2(0.f, 0.f);
  S2[0] = S2[1] = S2[2] = 0.f;
#pragma unroll
  for (int h = 0; h < SUB; ++h) {
    const int ly = a.ly0 + h * kWarpY;
    const int gy = a.y0 + ly;
    const bool valid = (gx < p.X) && (gy < p.Y);
    const float yf = (float)gy;
    float c0[3], c1[3], c2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = sBeta[d];
      v = fmaf(sBeta[3 + d], xf, v);
      v = fmaf(sBeta[6 + d], yf, v);
      v = fmaf(sBeta[12 + d], xf * xf, v);
      v = fmaf(sBeta[15 + d], yf * yf, v);
      v = fmaf(sBeta[21 + d], xf * yf, v);
      c0[d] = v;
      c1[d] = fmaf(sBeta[27 + d], yf, fmaf(sBeta[24 + d], xf, sBeta[9 + d]));
      c2[d] = sBeta[18 + d];
      if (FAST_DIV) {  // exact doubling: the main loop then evaluates 2q directly
        c0[d] += c0[d];
        c1[d] += c1[d];
        c2[d] += c2[d];
      }
      S0[h][d] = 0.f;
      S1[h][d] = 0.f;
    }
    S0xy[h] = make_float2(0.f, 0.f);
    S1xy[h] = make_float2(0.f, 0.f);
    const float2 c0xy = make_float2(c0[0], c0[1]), c1xy = make_float2(c1[0], c1[1]), c2xy = make_float2(c2[0], c2[1]);
    unsigned yaddr = smem_u32(a.sY + a.lx * a.RS + ly * a.zs);  // the lane's column of the Y tile, 4 B per z step
    float zf = (float)a.z0;
#pragma unroll kZUnroll
    for (int zz = 0; zz < a.nz; ++zz, zf += 1.f, yaddr += 4u) {
      // with FAST_DIV the Horner coefficients are pre-doubled, so q* below is 2q exactly; the x and y axes
      // go through the chain as one packed FP32x2 stream (same IEEE roundings per half), z stays scalar
      float ix0, ix1, ix2;
      if (FAST_DIV) {
        const float2 zz2 = make_float2(zf, zf);
        const float2 qxy = __ffma2_rn(zz2, __ffma2_rn(zz2, c2xy, c1xy), c0xy);
        const float q2 = fmaf(zf, fmaf(zf, c2[2], c1[2]), c0[2]);
        const float2 t0 = __fmul2_rn(qxy, rcpxy);
        const float2 r = __ffma2_rn(make_float2(-t0.x, -t0.y), sm1xy, qxy);
        const float2 v = __ffma2_rn(r, rcpxy, t0);
        const float2 u = __fadd2_rn(v, make_float2(-1.f, -1.f));
        const float2 ixy = __fmul2_rn(__fadd2_rn(u, make_float2(1.f, 1.f)), hsm1xy);
        ix0 = ixy.x;
        ix1 = ixy.y;
        ix2 = sample_coord_fast(q2, sm1z, rcpz, hsm1z);
      } else {
        const float q0 = fmaf(zf, fmaf(zf, c2[0], c1[0]), c0[0]);
        const float q1 = fmaf(zf, fmaf(zf, c2[1], c1[1]), c0[1]);
        const float q2 = fmaf(zf, fmaf(zf, c2[2], c1[2]), c0[2]);
        ix0 = sample_coord(q0, sm1x);
        ix1 = sample_coord(q1, sm1y);
        ix2 = sample_coord(q2, sm1z);
      }
      // floor / fraction.  The clamp to the (conservative) window bounds every table access; samples
      // below / above the table domain [-2, s] land on its first / last entry, which are zero, so no
      // float clamp is needed.
      const int i0 = __float2int_rd(ix0), i1 = __float2int_rd(ix1), i2 = __float2int_rd(ix2);
      const float f0 = ix0 - (float)i0, f1 = ix1 - (float)i1, f2 = ix2 - (float)i2;
      const unsigned o0 = (unsigned)min(max(i0 - wl0, 0), W0m1);
      const unsigned o1 = (unsigned)min(max(i1 - wl1, 0), W1m1);
      const unsigned o2 = (unsigned)min(max(i2 - wl2, 0), W2m1);
      float yh, g0, g1, g2;
      {
        unsigned ax = o0 * strideB + baseX, ay = o1 * strideB + baseY, az = o2 * strideB + baseZ;
        const float2 ff0 = make_float2(f0, f0), ff1 = make_float2(f1, f1), ff2 = make_float2(f2, f2);
        float2 yh2 = make_float2(0.f, 0.f), g02 = yh2, g12 = yh2, g22 = yh2;
#pragma unroll 1
        for (int j = 0; j < nquad; ++j, ax += 32u, ay += 32u, az += 32u) {  // four neurons per iteration
          const float4 ex = lds128<0>(ax), ey = lds128<0>(ay), ez = lds128<0>(az);
          const float4 fx = lds128<16>(ax), fy = lds128<16>(ay), fz = lds128<16>(az);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(ey.x, ey.y),
                           make_float2(ey.z, ey.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), ff0, ff1, ff2,
                           yh2, g02, g12, g22);
          pair2_accumulate(make_float2(fx.x, fx.y), make_float2(fx.z, fx.w), make_float2(fy.x, fy.y),
                           make_float2(fy.z, fy.w), make_float2(fz.x, fz.y), make_float2(fz.z, fz.w), ff0, ff1, ff2,
                           yh2, g02, g12, g22);
        }
        if (nst & 2) {
          const float4 ex = lds128<0>(ax), ey = lds128<0>(ay), ez = lds128<0>(az);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(ey.x, ey.y),
                           make_float2(ey.z, ey.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), ff0, ff1, ff2,
                           yh2, g02, g12, g22);
          ax += 16u;
          ay += 16u;
          az += 16u;
        }
        yh = yh2.x + yh2.y;
        g0 = g02.x + g02.y;
        g1 = g12.x + g12.y;
        g2 = g22.x + g22.y;
        if (nst & 1) {  // odd tail slot: lanes (x, z) of its float4 hold (G, D)
          const float4 ex = lds128<0>(ax), ey = lds128<0>(ay), ez = lds128<0>(az);
          pair_accumulate(ex.x, ex.z, ey.x, ey.z, ez.x, ez.z, f0, f1, f2, yh, g0, g1, g2);
        }
      }
      if (has_overflow) {  // slots beyond the staged capacity: straight from the L2-resident tables
        const int j0 = o0 + wl0 + 2, j1 = o1 + wl1 + 2, j2 = o2 + wl2 + 2;
        for (int j = nst; j < L; ++j) {
          const int k = a.sList[j];
          const float ck = __ldg(p.C + (size_t)k * p.T + t);
          const float2 ex = __ldg(p.tab0 + (size_t)k * sX3 + j0);
          const float2 ey = __ldg(p.tab1 + (size_t)k * sY3 + j1);
          const float2 ez = __ldg(p.tab2 + (size_t)k * sZ3 + j2);
          pair_accumulate(ex.x * ck, ex.y * ck, ey.x, ey.y, ez.x, ez.y, f0, f1, f2, yh, g0, g1, g2);
        }
      }
      const float yv = lds32(yaddr);
      if (WRITE_YHAT) sts32(yaddr, yh);
      const float r = valid ? (WRITE_RES ? ((yh + bg) - yv) : (yh - yv)) : 0.f;
      if (WRITE_RES) {
        sts32(yaddr, r);
        sum_r += r;
      }
      sse = fmaf(r, r, sse);
      // gradient moments: axes (x, y) as one packed FP32x2 stream, z scalar
      const float zf2 = zf * zf;
      const float2 h01 = __fmul2_rn(make_float2(r, r), make_float2(g0, g1));
      const float h2 = r * g2;
      S0xy[h] = __fadd2_rn(S0xy[h], h01);
      S1xy[h] = __ffma2_rn(make_float2(zf, zf), h01, S1xy[h]);
      S2xy = __ffma2_rn(make_float2(zf2, zf2), h01, S2xy);
      S0[h][2] += h2;
      S1[h][2] = fmaf(zf, h2, S1[h][2]);
      S2[2] = fmaf(zf2, h2, S2[2]);
    }
    S0[h][0] = S0xy[h].x;
    S0[h][1] = S0xy[h].y;
    S1[h][0] = S1xy[h].x;
    S1[h][1] = S1xy[h].y;
  }  // sub-tiles
  S2[0] = S2xy.x;
  S2[1] = S2xy.y;
}

// SUB = y-adjacent 8x4 sub-tiles per warp: they share the tile prologue (window, list, staging, TMA) and the
// reduction epilogue, and the unrolled march carries them as one packed stream.
// MODE 0: fit (loss + gradient).  MODE 1: forward only, writes Yhat.  MODE 2: fit with a scalar background
// added to Yhat, and the residual written out for the shared-parameter gradient kernel (extension).
//
// One CTA owns one spatial tile and walks `fpc` consecutive frames of the batch through it.  What depends on
// the tile only -- the static candidate neurons and their node ranges -- is fetched from global memory once
// and kept in shared memory.  What depends on the frame and has a known address -- beta_t, C[candidates, t],
// the tile of Y_t (TMA) -- is requested one frame ahead, while the current frame is in its main loop.  The
// staged table slices are kept across frames: they are gathered again from the L2-resident tables only when
// the frame's window or neuron list differs from the previous frame's; otherwise only the x slice is rescaled
// by the frame's traces.  In the steady state no global-memory latency sits between two main loops.
template <int NWX, int NWY, int SUB, int MODE, bool FAST_DIV>
__global__ void __launch_bounds__(32 * NWX * NWY, (NWX * NWY == 1) ? (MODE == 3 ? DNMF_MU_MINB : DNMF_MINB) : 1) fit_tile_kernel(const __grid_constant__ FitParams p) {
  constexpr bool WRITE_YHAT = MODE == 1;
  constexpr bool WRITE_RES = MODE == 2;
  constexpr int NW = NWX * NWY;
  constexpr int NT = 32 * NW;
  constexpr int TX = kWarpX * NWX, TY = kWarpY * NWY * SUB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int wsum = p.wmax0 + p.wmax1 + p.wmax2;
  const int CAP = p.cap;
  float2* sTab = reinterpret_cast<float2*>(smem_raw);            // [entry][slot], slot stride 8 B
  float2* sXraw = sTab + (size_t)CAP * wsum;                     // x slice before the traces are folded in
  float* sY = reinterpret_cast<float*>(smem_raw + ((((size_t)CAP * (wsum + p.wmax0)) * 8 + 127) & ~(size_t)127));  // 128 B: TMA
  const int zs = p.full_depth ? p.Z : p.tz;  // smem z-stride between y rows
  const int RS = p.y_pitch;                  // smem stride between x rows: TY*zs, padded when that pitch would
                                             // put the lanes of a warp on the same bank (configure_tiling_fixed)
  float* sRed = sY + (TX * RS + 4);
  float* sBeta = sRed + NW * kNumPartials;          // 32 floats
  int* sInt = reinterpret_cast<int*>(sBeta + 32);   // 32 ints: win[6], cnt[NW], flags
  float* sK = reinterpret_cast<float*>(sInt + 32);  // 16 floats: main-loop constants (see below)
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sK + 16);  // 16 B
  float* sCk = reinterpret_cast<float*>(sBar + 2);  // CAP + 2 floats: trace of each staged slot
  int* sCandRng = reinterpret_cast<int*>(sCk + ((CAP + 5) & ~3));       // [cand_cap][6]
  float* sCandC = reinterpret_cast<float*>(sCandRng + (size_t)p.cand_cap * 6);  // [cand_cap]
  unsigned short* sCand = reinterpret_cast<unsigned short*>(sCandC + p.cand_cap);  // [cand_cap (even)]
  unsigned short* sSlotCand = sCand + ((p.cand_cap + 7) & ~7);  // [CAP]: candidate index of each staged slot
  unsigned short* sList = sSlotCand + ((CAP + 7) & ~7);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto cta_sync = [&]() {
    if (NW == 1) __syncwarp(); else __syncthreads();
  };
  // grid = (ntx, nty, chunks*ntz): no integer division on the common ntz == 1 path
  const int bx = blockIdx.x, by = blockIdx.y;
  const int chunk = p.ntz == 1 ? (int)blockIdx.z : (int)blockIdx.z / p.ntz;
  const int bz = p.ntz == 1 ? 0 : (int)blockIdx.z - chunk * p.ntz;
  const int b_first = chunk * p.fpc;
  const int nb = min(p.fpc, p.B - b_first);  // frames this CTA walks (<= 32)
  const int x0 = bx * TX, y0 = by * TY, z0 = bz * p.tz;
  const int nx = min(TX, p.X - x0), ny = min(TY, p.Y - y0), nz = min(p.tz, p.Z - z0);
  const size_t Nvox = (size_t)p.X * p.Y * p.Z;
  const int my_frame = lane < nb ? p.frame_ids[b_first + lane] : 0;

  // ---- once per CTA: the tile's static candidates and their node ranges -> shared memory ----
  int cand_r0 = 0, ncand = 0;
  bool have_cand = false;
  if (p.cand_off != nullptr) {
    const int tile = (bz * p.nty + by) * p.ntx + bx;
    cand_r0 = (int)p.cand_off[tile];
    ncand = (int)p.cand_off[tile + 1] - cand_r0;
    have_cand = ncand <= p.cand_cap;
  }
  if (have_cand) {
    for (int i = tid; i < ncand; i += NT) {
      const int k = p.cand_ids[cand_r0 + i];
      sCand[i] = (unsigned short)k;
#pragma unroll
      for (int q = 0; q < 6; ++q) sCandRng[i * 6 + q] = p.rng[(size_t)k * 6 + q];
    }
  }
  const bool track_c = have_cand && ncand <= 2 * NT;  // candidate traces carried in shared memory
  const bool prefetch_c = track_c && MODE != 3;       // the trace statistics do not read C
  // Loop constants of the march go through shared memory: values the compiler can trace back to kernel
  // parameters are rematerialised inside the z loop (constant-bank loads, int->float conversions, address
  // arithmetic: ~25 instructions per z step), values loaded from shared memory stay in registers.
  if (tid == 0) {
    const float s0 = (float)(p.X - 1), s1 = (float)(p.Y - 1), s2 = (float)(p.Z - 1);
    sK[0] = p.rcp0, sK[1] = p.rcp1, sK[2] = p.rcp2;
    sK[3] = s0, sK[4] = s1, sK[5] = s2;
    sK[6] = 0.5f * s0, sK[7] = 0.5f * s1, sK[8] = 0.5f * s2;
    sK[9] = __uint_as_float((unsigned)CAP * 8u);
    sK[10] = __uint_as_float((unsigned)(kWarpY * zs) * 4u);
    sK[11] = 0.f;
    const unsigned bX = smem_u32(sTab), sB = (unsigned)CAP * 8u;
    sK[12] = __uint_as_float(bX);
    sK[13] = __uint_as_float(bX + (unsigned)p.wmax0 * sB);
    sK[14] = __uint_as_float(bX + (unsigned)(p.wmax0 + p.wmax1) * sB);
    sK[15] = 0.f;
  }

  // ---- frame-tile loads: one bulk async copy (TMA) per contiguous run of ty*Z floats ----
  const int run = ny * p.Z;
  const bool bulk = !WRITE_YHAT && p.bulk_ok && p.full_depth && ((run & 3) == 0);
  const unsigned bar = smem_u32(sBar);
  if (bulk && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cta_sync();
  auto load_tile = [&](int fi) {  // frame fi of this CTA -> sY
    const int b = b_first + fi;
    const int t = __shfl_sync(0xffffffffu, my_frame, fi);
    const float* __restrict__ frame = p.frames + (size_t)(p.frames_are_batch ? b : t) * Nvox;
    if (bulk && p.tmap_ok) {
      if (tid == 0) {  // the whole 8 x (ty*Z) tile in one tensor copy; rows/columns past the volume are zero-filled
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TX * RS * 4)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                "r"(smem_u32(sY)), "l"(reinterpret_cast<unsigned long long>(&p.tmap)), "r"(y0 * p.Z), "r"(x0),
            "r"(p.frames_are_batch ? p.b_base + b : t), "r"(bar)
            : "memory");
      }
    } else if (bulk) {
      if (tid == 0) {  // one thread arms the barrier and issues every row copy (uniform-datapath instructions)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nx * run * 4)
                     : "memory");
        const float* src = frame + ((size_t)x0 * p.Y + y0) * p.Z;
        unsigned dst = smem_u32(sY);
        const size_t src_step = (size_t)p.Y * p.Z;
        for (int r = 0; r < nx; ++r, src += src_step, dst += (unsigned)RS * 4u)
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
              "l"(src), "r"(run * 4), "r"(bar)
              : "memory");
      }
    } else if (p.full_depth) {
      for (int lx = warp; lx < nx; lx += NW) {
        const float* src = frame + ((size_t)(x0 + lx) * p.Y + y0) * p.Z;
        for (int e = lane; e < run; e += 32) sY[lx * RS + e] = __ldg(src + e);
      }
    } else {
      for (int row = warp; row < nx * TY; row += NW) {
        int lx = row / TY, ly = row - lx * TY;
        if (ly < ny) {
          const float* src = frame + ((size_t)(x0 + lx) * p.Y + (y0 + ly)) * p.Z + z0;
          for (int e = lane; e < nz; e += 32) sY[lx * RS + ly * zs + e] = __ldg(src + e);
        }
      }
    }
  };
  if (!WRITE_YHAT && bulk) load_tile(0);

  // ---- requested one frame ahead: beta_t and the candidates' traces ----
  float beta_next = 0.f, cc_next[2] = {0.f, 0.f};
  auto prefetch_frame = [&](int fi) {
    const int t = __shfl_sync(0xffffffffu, my_frame, fi);
    if (tid < 30) beta_next = p.beta[(size_t)tid * p.T + t];
    if (prefetch_c) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = tid + u * NT;
        if (i < ncand) cc_next[u] = p.C[(size_t)sCand[i] * p.T + t];
      }
    }
  };
  prefetch_frame(0);

  // lane geometry (frame independent)
  const int lx = (warp % NWX) * kWarpX + (lane & 7);
  const int ly0 = (warp / NWX) * (kWarpY * SUB) + (lane >> 3);
  const int gx = x0 + lx;
  const float xf = (float)gx;
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
  const unsigned strideB = (unsigned)CAP * 8u;
  const unsigned baseX = smem_u32(sTab);
  const unsigned baseY = baseX + (unsigned)p.wmax0 * strideB;
  const unsigned baseZ = baseY + (unsigned)p.wmax1 * strideB;
  const float bg = WRITE_RES ? p.bg : 0.f;

  // state carried from frame to frame: what the staged slices were built for
  int pw_lo[3] = {0x7fffffff, 0, 0}, pw_hi[3] = {0, 0, 0}, prev_L = -1;
  int n_restaged = 0;  // frames whose slices had to be (re)built
  bool prev_fast = false;  // previous list came from the cached candidates with prefetched traces

  for (int fi = 0; fi < nb; ++fi) {
    const int b = b_first + fi;
    const int t = __shfl_sync(0xffffffffu, my_frame, fi);
    if (tid < 30) sBeta[tid] = beta_next;
    if (prefetch_c) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = tid + u * NT;
        if (i < ncand) sCandC[i] = cc_next[u];
      }
    }
    if (!WRITE_YHAT && !bulk) load_tile(fi);
    cta_sync();
    if (fi + 1 < nb) prefetch_frame(fi + 1);

    // ---- conservative window of this tile under beta_t (same code as the binning kernel) ----
    if (tid < 3) {
      const int s = tid == 0 ? p.X : (tid == 1 ? p.Y : p.Z);
      int wlo, whi;
      bool clipped;
      tile_window_axis(sBeta + tid, 3, (float)x0, (float)y0, (float)z0, (float)(x0 + nx - 1),
                       (float)(y0 + ny - 1), (float)(z0 + nz - 1), s, wlo, whi, clipped);
      sInt[tid] = wlo;
      sInt[3 + tid] = whi;
      sInt[24 + tid] = clipped ? 1 : 0;
    }
    cta_sync();
    int wlo[3], whi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      wlo[d] = sInt[d];
      whi[d] = sInt[3 + d];
    }
    const bool window_clipped = (sInt[24] | sInt[25] | sInt[26]) != 0;

    // ---- neuron list: ascending k, ballot compaction (single pass for one warp, two passes else).
    // Candidates come from the tile's static list (in shared memory) when the window stays inside the
    // expanded identity window the list was built for, else from a scan over all K neurons.
    int L = 0;
    bool changed = false;
    bool same_window = prev_fast;
#pragma unroll
    for (int d = 0; d < 3; ++d) same_window = same_window && (wlo[d] == pw_lo[d]) && (whi[d] == pw_hi[d]);
    if (same_window) {
      // the list is a function of the window and the static candidates: unchanged.  Only the traces of the
      // staged slots are new.
      L = prev_L;
      if constexpr (MODE != 3)
        for (int pos = tid; pos < min(L, CAP); pos += NT) sCk[pos] = sCandC[sSlotCand[pos]];
    } else {
      bool inside = false;
      if (p.cand_off != nullptr) {
        const int e = p.cand_expand;
        inside = wlo[0] >= max(x0 - 1, -2) - e && whi[0] <= min(x0 + nx, p.X) + e &&
                 wlo[1] >= max(y0 - 1, -2) - e && whi[1] <= min(y0 + ny, p.Y) + e &&
                 wlo[2] >= max(z0 - 1, -2) - e && whi[2] <= min(z0 + nz, p.Z) + e;
      }
      const bool from_smem = inside && have_cand;
      const int* __restrict__ cand = (inside && !have_cand) ? p.cand_ids + cand_r0 : nullptr;
      const int r1 = (inside) ? ncand : p.K;
      // candidate idx -> (k, in window, trace)
      auto probe = [&](int idx, int& k, float& ck) -> bool {
        k = -1;
        ck = 0.f;
        if (idx >= r1) return false;
        if (from_smem) {
          k = sCand[idx];
          if (!neuron_in_window(sCandRng + idx * 6, wlo, whi)) return false;
          if constexpr (MODE != 3) ck = prefetch_c ? sCandC[idx] : __ldg(p.C + (size_t)k * p.T + t);
          return true;
        }
        k = cand ? cand[idx] : idx;
        if (!neuron_in_window(p.rng + (size_t)k * 6, wlo, whi)) return false;
        if constexpr (MODE != 3) ck = __ldg(p.C + (size_t)k * p.T + t);
        return true;
      };
      auto put = [&](int pos, int k, float ck, int idx) {
        changed |= (pos >= prev_L) || (sList[pos] != (unsigned short)k);
        sList[pos] = (unsigned short)k;
        if (pos < CAP) {
          sCk[pos] = ck;
          sSlotCand[pos] = (unsigned short)idx;
        }
      };
      if (NW == 1) {
        for (int c0i = 0; c0i < r1; c0i += 32) {
          int k;
          float ck;
          const bool ok = probe(c0i + lane, k, ck);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) put(L + __popc(m & ((1u << lane) - 1u)), k, ck, c0i + lane);
          L += __popc(m);
        }
        changed = __any_sync(0xffffffffu, changed);
      } else {
        const int per = ((r1 + NT - 1) / NT) * 32;
        const int kb = warp * per;
        int cnt = 0;
        for (int c0i = kb; c0i < kb + per; c0i += 32) {
          int k;
          float ck;
          cnt += __popc(__ballot_sync(0xffffffffu, probe(c0i + lane, k, ck)));
        }
        if (lane == 0) sInt[8 + warp] = cnt;
        __syncthreads();
        int off = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const int c = sInt[8 + w];
          if (w < warp) off += c;
          L += c;
        }
        for (int c0i = kb; c0i < kb + per; c0i += 32) {
          int k;
          float ck;
          const bool ok = probe(c0i + lane, k, ck);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) put(off + __popc(m & ((1u << lane) - 1u)), k, ck, c0i + lane);
          off += __popc(m);
        }
        changed = __syncthreads_or(changed ? 1 : 0) != 0;
      }
      changed = changed || (L != prev_L);
#pragma unroll
      for (int d = 0; d < 3; ++d) changed = changed || (wlo[d] != pw_lo[d]) || (whi[d] != pw_hi[d]);
      prev_fast = from_smem && track_c;
    }
    cta_sync();

    // ---- table slices of the first nst listed neurons ----
    const int W0 = whi[0] - wlo[0] + 1, W1 = whi[1] - wlo[1] + 1, W2 = whi[2] - wlo[2] + 1;
    const bool fits = (W0 <= p.wmax0) && (W1 <= p.wmax1) && (W2 <= p.wmax2);
    const int nst = fits ? min(L, CAP) : 0;
    const int npair = (nst + 1) >> 1;
    if ((nst & 1) && tid == 0) sCk[nst] = 0.f;  // partner of the last neuron of an odd list: zero footprint
    if (changed) {
      ++n_restaged;
      // One thread owns table entry e of every slot.  Slot pairs (2p, 2p+1) share one float4
      // (G_2p, G_2p+1, D_2p, D_2p+1) = the packed operands of FFMA2; an odd list is completed with a zero
      // footprint.  Loads are issued four pairs at a time ahead of the stores.  The x slice is kept without
      // the traces (they change with the frame, the slices usually do not).
      const int Wt = W0 + W1 + W2;
      for (int e = tid; e < Wt; e += NT) {
        const float2* src;
        int row;
        float4* dst;
        if (e < W0) {
          src = p.tab0 + (wlo[0] + 2 + e);
          row = sX3;
          dst = reinterpret_cast<float4*>(sXraw + (size_t)e * CAP);
        } else if (e < W0 + W1) {
          src = p.tab1 + (wlo[1] + 2 + (e - W0));
          row = sY3;
          dst = reinterpret_cast<float4*>(sTab + (size_t)(p.wmax0 + (e - W0)) * CAP);
        } else {
          src = p.tab2 + (wlo[2] + 2 + (e - W0 - W1));
          row = sZ3;
          dst = reinterpret_cast<float4*>(sTab + (size_t)(p.wmax0 + p.wmax1 + (e - W0 - W1)) * CAP);
        }
        constexpr int kBatch = 4;
        for (int p0 = 0; p0 < npair; p0 += kBatch) {
          float2 va[kBatch], vb[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int j = 2 * (p0 + u);
            va[u] = vb[u] = make_float2(0.f, 0.f);
            if (j < nst) va[u] = __ldg(src + (size_t)sList[j] * row);
            if (j + 1 < nst) vb[u] = __ldg(src + (size_t)sList[j + 1] * row);
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u)
            if (p0 + u < npair) dst[p0 + u] = make_float4(va[u].x, vb[u].x, va[u].y, vb[u].y);
        }
      }
    }
    cta_sync();
    // x slice of this frame: C[k,t] folded in
    for (int e = tid; e < W0; e += NT) {
      const float4* src = reinterpret_cast<const float4*>(sXraw + (size_t)e * CAP);
      float4* dst = reinterpret_cast<float4*>(sTab + (size_t)e * CAP);
      for (int pp = 0; pp < npair; ++pp) {
        const float4 v = src[pp];
        const float2 c = MODE == 3 ? make_float2(1.f, 1.f) : *reinterpret_cast<const float2*>(sCk + 2 * pp);
        dst[pp] = make_float4(v.x * c.x, v.y * c.y, v.z * c.x, v.w * c.y);
      }
    }
    cta_sync();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      pw_lo[d] = wlo[d];
      pw_hi[d] = whi[d];
    }
    prev_L = L;

    if (bulk) {  // wait for the bulk copies of the Y tile (phase fi of the mbarrier)
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(fi & 1)
            : "memory");
      }
    }

    auto fill_march_args = [&](MarchArgs& a) {
      const float yfA = (float)(y0 + ly0), yfB = (float)(y0 + ly0 + kWarpY);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float b0 = sBeta[d], bx_ = sBeta[3 + d], by_ = sBeta[6 + d], bz_ = sBeta[9 + d];
        const float bxx = sBeta[12 + d], byy = sBeta[15 + d], bxy = sBeta[21 + d], bxz = sBeta[24 + d],
                    byz = sBeta[27 + d];
        float vA = fmaf(bx_, xf, b0), vB = vA;
        vA = fmaf(by_, yfA, vA);
        vB = fmaf(by_, yfB, vB);
        vA = fmaf(bxx, xf * xf, vA);
        vB = fmaf(bxx, xf * xf, vB);
        vA = fmaf(byy, yfA * yfA, vA);
        vB = fmaf(byy, yfB * yfB, vB);
        vA = fmaf(bxy, xf * yfA, vA);
        vB = fmaf(bxy, xf * yfB, vB);
        const float wA = fmaf(byz, yfA, fmaf(bxz, xf, bz_)), wB = fmaf(byz, yfB, fmaf(bxz, xf, bz_));
        a.c0[d] = make_float2(vA + vA, vB + vB);  // exact doubling: the march evaluates 2q directly
        a.c1[d] = make_float2(wA + wA, wB + wB);
        a.c2[d] = sBeta[18 + d] + sBeta[18 + d];
      }
      {
        const unsigned ka = smem_u32(sK);
        const float4 k0 = lds128r(ka), k1 = lds128r(ka + 16u), k2 = lds128r(ka + 32u), k3 = lds128r(ka + 48u);
        a.base[0] = __float_as_uint(k3.x), a.base[1] = __float_as_uint(k3.y), a.base[2] = __float_as_uint(k3.z);
        a.rcp[0] = k0.x, a.rcp[1] = k0.y, a.rcp[2] = k0.z;
        a.sm1[0] = k0.w, a.sm1[1] = k1.x, a.sm1[2] = k1.y;
        a.hsm1[0] = k1.z, a.hsm1[1] = k1.w, a.hsm1[2] = k2.x;
        a.strideB = __float_as_uint(k2.y);
        a.yoffB = __float_as_uint(k2.z);
        a.oz = k2.w;
      }
      a.wl[0] = wlo[0], a.wl[1] = wlo[1], a.wl[2] = wlo[2];
      a.wm1[0] = W0 - 1, a.wm1[1] = W1 - 1, a.wm1[2] = W2 - 1;
      a.yaddrA = smem_u32(sY + lx * RS + ly0 * zs);
      a.zf0 = (float)z0;
      a.nz = nz;
      a.validA = (gx < p.X) && (y0 + ly0 < p.Y);
      a.validB = (gx < p.X) && (y0 + ly0 + kWarpY < p.Y);
      a.bg = bg;
      a.zskew = 0;
      if (p.z_skew != 0 && nz >= 4) a.zskew = ((lane >> 3) * p.z_skew) & 3;
    };

    if constexpr (MODE == 3) {
      // ---- trace statistics of this tile-frame ----
      if constexpr (SUB == 2 && FAST_DIV) {
        if (L > nst) {
          if (tid == 0) atomicMax(p.mu_overflow, L);  // not fully staged: the caller reruns the generic kernel
        } else if (npair > 0) {
          MarchArgs a;
          fill_march_args(a);
          double* Gt = p.muG + (size_t)t * p.K * p.K;
          double* bt = p.mub + (size_t)t * p.K;
          const int nblk = (npair + 2) / 3;
          for (int pb = 0; pb < nblk; ++pb) {
            for (int lb = pb; lb < nblk; ++lb) {
              float2 G[3][6], bv[3];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                bv[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 6; ++c) G[i][c] = make_float2(0.f, 0.f);
              }
              const int rowp = min(3, npair - 3 * pb), colp = min(3, npair - 3 * lb);
              const unsigned rowoff = 48u * (unsigned)pb, coloff = 48u * (unsigned)lb;
              if (pb != lb) {
                march_stats<3, true>(a, rowoff, coloff, rowp, colp, G, bv);
                flush_stats<3, true>(G, bv, pb, lb, nst, sList, Gt, bt, p.K, lane);
              } else if (rowp == 1) {
                march_stats<1, false>(a, rowoff, rowoff, 1, 1, G, bv);
                flush_stats<1, false>(G, bv, pb, lb, nst, sList, Gt, bt, p.K, lane);
              } else if (rowp == 2) {
                march_stats<2, false>(a, rowoff, rowoff, 2, 2, G, bv);
                flush_stats<2, false>(G, bv, pb, lb, nst, sList, Gt, bt, p.K, lane);
              } else {
                march_stats<3, false>(a, rowoff, rowoff, 3, 3, G, bv);
                flush_stats<3, false>(G, bv, pb, lb, nst, sList, Gt, bt, p.K, lane);
              }
            }
          }
        }
      }
      cta_sync();
      if (bulk && fi + 1 < nb) load_tile(fi + 1);
      continue;
    }

    // ---- main loop: one (x,y) column per lane (per sub-tile), march along z ----
    float S0[SUB][3], S1[SUB][3], S2[3] = {0.f, 0.f, 0.f};
    float sse = 0.f, sum_r = 0.f;
    const bool has_overflow = L > nst;
    bool marched = false;
    if constexpr (SUB == 2 && FAST_DIV) {
      if (!has_overflow && (npair <= kMaxNP || !DNMF_UNROLLED_MARCH)) {
        MarchArgs a;
        fill_march_args(a);
        MarchOut o;
        const bool safe = DNMF_ALWAYS_SAFE || window_clipped || nx < TX || ny < TY;
#if DNMF_UNROLLED_MARCH
        switch (npair * 2 + (safe ? 1 : 0)) {
#define DNMF_MARCH(n)                   \
  case 2 * n:                           \
    march_pairs<n, false, MODE>(a, o);  \
    break;                              \
  case 2 * n + 1:                       \
    march_pairs<n, true, MODE>(a, o);   \
    break;
          DNMF_MARCH(0)
          DNMF_MARCH(1)
          DNMF_MARCH(2)
          DNMF_MARCH(3)
          DNMF_MARCH(4)
          DNMF_MARCH(5)
          DNMF_MARCH(6)
          DNMF_MARCH(7)
          DNMF_MARCH(8)
#undef DNMF_MARCH
          default:
            break;
        }
#else
        if (npair == 0) {
          if (safe)
            march_pairs<0, true, MODE>(a, o);
          else
            march_pairs<0, false, MODE>(a, o);
        } else {
          const int npf = nst >> 1;  // full slot pairs; an odd list ends with a single slot
          const int tail = (nst & 1) ? (npf == 0 ? 2 : 1) : 0;
          if (p.dyn_tail) {
            switch ((p.z_skew != 0 ? 2 : 0) + (safe ? 1 : 0)) {
              case 0: march_rolled<false, MODE, 3, false>(a, npf, o, tail); break;
              case 1: march_rolled<true, MODE, 3, false>(a, npf, o, tail); break;
              case 2: march_rolled<false, MODE, 3, true>(a, npf, o, tail); break;
              default: march_rolled<true, MODE, 3, true>(a, npf, o, tail); break;
            }
          } else
#if DNMF_MERGE_TAIL01
          switch ((p.z_skew != 0 ? 4 : 0) + (tail == 2 ? 2 : 0) + (safe ? 1 : 0)) {
            case 0: march_rolled<false, MODE, 4, false>(a, npf, o, tail); break;
            case 1: march_rolled<true, MODE, 4, false>(a, npf, o, tail); break;
            case 2: march_rolled<false, MODE, 2, false>(a, npf, o); break;
            case 3: march_rolled<true, MODE, 2, false>(a, npf, o); break;
            case 4: march_rolled<false, MODE, 4, true>(a, npf, o, tail); break;
            case 5: march_rolled<true, MODE, 4, true>(a, npf, o, tail); break;
            case 6: march_rolled<false, MODE, 2, true>(a, npf, o); break;
            default: march_rolled<true, MODE, 2, true>(a, npf, o); break;
          }
#else
          switch ((p.z_skew != 0 ? 6 : 0) + tail * 2 + (safe ? 1 : 0)) {
            case 0: march_rolled<false, MODE, 0, false>(a, npf, o); break;
            case 1: march_rolled<true, MODE, 0, false>(a, npf, o); break;
            case 2: march_rolled<false, MODE, 1, false>(a, npf, o); break;
            case 3: march_rolled<true, MODE, 1, false>(a, npf, o); break;
            case 4: march_rolled<false, MODE, 2, false>(a, npf, o); break;
            case 5: march_rolled<true, MODE, 2, false>(a, npf, o); break;
            case 6: march_rolled<false, MODE, 0, true>(a, npf, o); break;
            case 7: march_rolled<true, MODE, 0, true>(a, npf, o); break;
            case 8: march_rolled<false, MODE, 1, true>(a, npf, o); break;
            case 9: march_rolled<true, MODE, 1, true>(a, npf, o); break;
            case 10: march_rolled<false, MODE, 2, true>(a, npf, o); break;
            default: march_rolled<true, MODE, 2, true>(a, npf, o); break;
          }
#endif
        }
#endif
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          S0[0][d] = o.S0[d].x;
          S0[1][d] = o.S0[d].y;
          S1[0][d] = o.S1[d].x;
          S1[1][d] = o.S1[d].y;
          S2[d] = o.S2[d].x + o.S2[d].y;
        }
        sse = o.sse.x + o.sse.y;
        sum_r = o.sum_r.x + o.sum_r.y;
        marched = true;
      }
    }
    if (!marched) {
      GenericArgs a;
      a.p = &p;
      a.sBeta = sBeta;
      a.sList = sList;
      a.sY = sY;
      a.t = t, a.L = L, a.nst = nst;
      a.x0 = x0, a.y0 = y0, a.z0 = z0, a.nz = nz;
      a.lx = lx, a.ly0 = ly0, a.RS = RS, a.zs = zs;
      a.wl[0] = wlo[0], a.wl[1] = wlo[1], a.wl[2] = wlo[2];
      a.wm1[0] = W0 - 1, a.wm1[1] = W1 - 1, a.wm1[2] = W2 - 1;
      a.base[0] = baseX, a.base[1] = baseY, a.base[2] = baseZ;
      a.strideB = strideB;
      a.bg = bg;
      march_generic<SUB, MODE, FAST_DIV>(a, S0, S1, S2, sse, sum_r);
    }

    // ---- expand z-moments with this lane's (x,y) monomials, transposing warp reduction ----
    const unsigned cta_linear = ((unsigned)(b * p.ntz + bz) * p.nty + by) * p.ntx + bx;
    {
      float v[32];
      const float xx = xf * xf;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float t0 = 0.f, t1 = 0.f, y0s = 0.f, yy0s = 0.f, y1s = 0.f;  // sums over the sub-tiles' y rows
#pragma unroll
        for (int h = 0; h < SUB; ++h) {
          const float yh_ = (float)(y0 + ly0 + h * kWarpY);
          t0 += S0[h][d];
          t1 += S1[h][d];
          y0s = fmaf(yh_, S0[h][d], y0s);
          yy0s = fmaf(yh_ * yh_, S0[h][d], yy0s);
          y1s = fmaf(yh_, S1[h][d], y1s);
        }
        v[0 * 3 + d] = t0;
        v[1 * 3 + d] = xf * t0;
        v[2 * 3 + d] = y0s;
        v[3 * 3 + d] = t1;
        v[4 * 3 + d] = xx * t0;
        v[5 * 3 + d] = yy0s;
        v[6 * 3 + d] = S2[d];
        v[7 * 3 + d] = xf * y0s;
        v[8 * 3 + d] = xf * t1;
        v[9 * 3 + d] = y1s;
      }
      v[30] = sse;
      v[31] = sum_r;  // sum of residuals (gradient of the scalar background), MODE 2 only
      const float tot = warp_transpose_sum(v, lane);
      if (NW == 1) {
        p.partials[(size_t)cta_linear * kNumPartials + lane] = tot;
      } else {
        sRed[warp * kNumPartials + lane] = tot;
      }
    }
    if (NW > 1) {
      __syncthreads();
      if (tid < kNumPartials) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += sRed[w * kNumPartials + tid];
        p.partials[(size_t)cta_linear * kNumPartials + tid] = v;
      }
    }
    if (WRITE_YHAT || WRITE_RES) {
      cta_sync();
      float* out = p.yhat + (size_t)b * Nvox;
      if (p.full_depth) {
        for (int lxx = warp; lxx < nx; lxx += NW) {
          float* dst = out + ((size_t)(x0 + lxx) * p.Y + y0) * p.Z;
          for (int e = lane; e < run; e += 32) dst[e] = sY[lxx * RS + e];
        }
      } else {
        for (int row = warp; row < nx * TY; row += NW) {
          int lxx = row / TY, lyy = row - lxx * TY;
          if (lyy < ny) {
            float* dst = out + ((size_t)(x0 + lxx) * p.Y + (y0 + lyy)) * p.Z + z0;
            for (int e = lane; e < nz; e += 32) dst[e] = sY[lxx * RS + lyy * zs + e];
          }
        }
      }
    }
    // every thread is done with this frame's Y tile, list and slices
    cta_sync();
    if (bulk && fi + 1 < nb) load_tile(fi + 1);
  }
  if (MODE == 0 && p.restage_count != nullptr && tid == 0 && n_restaged > 0)
    atomicAdd(p.restage_count + ((blockIdx.x + blockIdx.y) & 31), (unsigned)n_restaged);
}

// Second stage: per frame, sum the CTA partials in a fixed order (double), scale by 2/(B_global*N),
// write the frame's gradient column and its sum of squared residuals.  grid = B, block = 256.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, const int* __restrict__ frame_ids,
                                       int nt, int T, double grad_scale, float* __restrict__ grad,
                                       double* __restrict__ sse_out, double* __restrict__ sumr_out,
                                       const double* __restrict__ scale_per_frame = nullptr) {
  __shared__ double s[8][kNumPartials];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = partials + (size_t)b * nt * kNumPartials;
  double acc = 0.0;
  for (int i = warp; i < nt; i += 8) acc += (double)src[(size_t)i * kNumPartials + lane];
  s[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s[w][lane];
    const int t = frame_ids[b];
    if (scale_per_frame != nullptr) grad_scale = scale_per_frame[b];
    if (lane < 30) grad[(size_t)lane * T + t] = (float)(v * grad_scale);
    if (lane == 30) sse_out[b] = v;
    if (lane == 31 && sumr_out != nullptr) sumr_out[b] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 3a: dense Adam, torch _single_tensor_adam formula in fp32; block 0 also
// reduces the batch loss.  grad is consumed and zeroed so the dense buffer stays all-zero
// outside the next batch.
// ------------------------------------------------------------------------------------------------
struct AdamParams {
  float w1;         // 1 - beta1
  float b2;         // beta2
  float w2;         // 1 - beta2
  float step_size;  // lr / (1 - beta1^step)
  float bc2_sqrt;   // sqrt(1 - beta2^step)
  float eps;
};

__device__ __forceinline__ void adam_update(float& pi, float& mi, float& vi, float gi, float w1, float b2, float w2,
                                            float step_size, float bc2_sqrt, float eps) {
  mi = fmaf(w1, __fsub_rn(gi, mi), mi);
  vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(w2, gi), gi));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
  pi = __fsub_rn(pi, __fmul_rn(step_size, __fdiv_rn(mi, denom)));
}

// ------------------------------------------------------------------------------------------------
// Frame-parallel epoch (dnmf_motion_epoch).  The reference's model has no parameter shared between frames
// (Demix/dNMF.py:29-33: only beta[:, :, t] is learnable), and Adam is elementwise, so the minibatches of an
// epoch in which every frame occurs once touch disjoint columns: the column of frame t sees the steps of the
// other minibatches as Adam updates with a zero gradient (momentum keeps moving it) and one step
// with its own gradient.  The epoch therefore runs as
//   phase 0: each column replays the zero-gradient steps that precede its minibatch,
//   ONE fused launch over all frames of the epoch (each minibatch with its own 2/(B N) scale),
//   phase 1: each column takes its gradient step and replays the zero-gradient steps that follow,
// with the same fp32 operations in the same order per column as the step-by-step schedule: bit-identical
// (tests/test_gpu_edge.py::test_epoch_call_equals_per_step_calls).
// ------------------------------------------------------------------------------------------------
// batch_of[t] = minibatch index of frame t (-1: not in this epoch); scale[b] = 2/(B_i * global scale * N);
// status: bit 0 = a frame occurs twice, bit 1 = frame id out of range.
__global__ void epoch_index_kernel(const int* __restrict__ frame_ids, const int* __restrict__ offsets, int nbatches,
                                   int T, double n_vox, int global_batch_scale, int* __restrict__ batch_of,
                                   double* __restrict__ scale, int* __restrict__ status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= offsets[nbatches]) return;
  int lo = 0, hi = nbatches - 1;  // last i with offsets[i] <= b
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (offsets[mid] <= b) lo = mid; else hi = mid - 1;
  }
  const int Bi = offsets[lo + 1] - offsets[lo];
  scale[b] = 2.0 / ((double)Bi * (double)global_batch_scale * n_vox);
  const int t = frame_ids[b];
  if (t < 0 || t >= T) {
    atomicOr(status, 2);
    return;
  }
  if (atomicExch(batch_of + t, lo) != -1) atomicOr(status, 1);
}

// step_scalars[s] = (lr / (1 - beta1^step), sqrt(1 - beta2^step)) of step first_step + s, computed in double on
// the host like torch does.
__global__ void epoch_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int n, int T, int affine, float w1, float b2, float w2,
                                  float eps, const float2* __restrict__ step_scalars, int nsteps,
                                  const int* __restrict__ batch_of, int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int it = batch_of[i % T];
  int lo, hi;
  float gi = 0.f;
  if (phase == 0) {
    lo = 0;
    hi = it < 0 ? nsteps : it;
  } else {
    if (it < 0) return;
    lo = it;
    hi = nsteps;
    gi = g[i];
    g[i] = 0.f;
    if (affine && (i / (3 * T)) >= 4) gi = 0.f;
  }
  if (lo >= hi) return;
  float pi = p[i], mi = m[i], vi = v[i];
  for (int s_ = lo; s_ < hi; ++s_) {
    const float2 a = __ldg(step_scalars + s_);
    adam_update(pi, mi, vi, gi, w1, b2, w2, a.x, a.y, eps);
    gi = 0.f;
  }
  p[i] = pi;
  m[i] = mi;
  v[i] = vi;
}

// loss of minibatch i = sum of its frames' SSE / (B_i * global scale * N), summed as adam_kernel's block 0 does
__global__ void epoch_loss_kernel(const double* __restrict__ sse, const int* __restrict__ offsets, int nbatches,
                                  double n_vox, int global_batch_scale, double* __restrict__ loss_out) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= nbatches) return;
  const int b0 = offsets[i], B = offsets[i + 1] - b0;
  double acc = 0.0;
  for (int j = lane; j < B; j += 32) acc += sse[b0 + j];
  acc = warp_sum_d(acc);
  if (lane == 0) loss_out[i] = acc * (1.0 / ((double)B * (double)global_batch_scale * n_vox));
}

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int n, int row_len, int affine, AdamParams a,
                            const double* __restrict__ sse, int B, double loss_scale,
                            double* __restrict__ loss_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float gi = g[i];
    g[i] = 0.f;
    if (affine && (i / row_len) >= 4) gi = 0.f;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, gi, a.w1, a.b2, a.w2, a.step_size, a.bc2_sqrt, a.eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
  if (blockIdx.x == 0 && loss_out != nullptr && threadIdx.x < 32) {
    double acc = 0.0;
    for (int j = threadIdx.x; j < B; j += 32) acc += sse[j];
    acc = warp_sum_d(acc);
    if (threadIdx.x == 0) *loss_out = acc * loss_scale;
  }
}

__global__ void clamp_negative_kernel(float* __restrict__ x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = fmaxf(x[i], 0.f);
}

// ------------------------------------------------------------------------------------------------
// dense forward for the small-problem API outputs of ExponentialFP.forward (A_t, grid):
// one thread per (frame, voxel), all K neurons through the global tables.  Not a hot path.
// ------------------------------------------------------------------------------------------------
__global__ void dense_forward_kernel(Geom g, const int* __restrict__ frame_ids, int B,
                                     const float* __restrict__ beta, const float2* __restrict__ tab0,
                                     const float2* __restrict__ tab1, const float2* __restrict__ tab2,
                                     float* __restrict__ At, float* __restrict__ grid) {
  const size_t N = (size_t)g.X * g.Y * g.Z;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * B) return;
  const int b = (int)(idx / N);
  const size_t v = idx - (size_t)b * N;
  const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((size_t)g.Z * g.Y));
  const int t = frame_ids[b];
  const float xf = (float)x, yf = (float)y, zf = (float)z;
  const float phi[kBasis] = {1.f, xf, yf, zf, xf * xf, yf * yf, zf * zf, xf * yf, xf * zf, yf * zf};
  const int sz[3] = {g.X, g.Y, g.Z};
  int ii[3];
  float ff[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float q = 0.f;
#pragma unroll
    for (int a = 0; a < kBasis; ++a) q = __fadd_rn(q, __fmul_rn(phi[a], beta[((size_t)a * 3 + d) * g.T + t]));
    const float sm1 = (float)(sz[d] - 1);
    if (grid) {
      float u = sm1 == 0.f ? 0.f : __fsub_rn(__fdiv_rn(__fmul_rn(2.f, q), sm1), 1.f);
      grid[(v * 3 + d) * B + b] = u;
    }
    split_coord(sample_coord(q, sm1), sz[d], ii[d], ff[d]);
  }
  if (At) {
    for (int k = 0; k < g.K; ++k) {
      float2 ex = tab0[(size_t)k * (g.X + 3) + ii[0] + 2];
      float2 ey = tab1[(size_t)k * (g.Y + 3) + ii[1] + 2];
      float2 ez = tab2[(size_t)k * (g.Z + 3) + ii[2] + 2];
      float a = fmaf(ff[0], ex.y, ex.x) * (fmaf(ff[1], ey.y, ey.x) * fmaf(ff[2], ez.y, ez.x));
      At[((size_t)b * g.K + k) * N + v] = a;
    }
  }
}

}  // namespace dnmf

// ================================================================================================
// host side: context + C ABI
// ================================================================================================
using namespace dnmf;

struct dnmf_ctx {
  int X = 0, Y = 0, Z = 0, K = 0, T = 0, device = 0;
  size_t N = 0;
  int num_sms = 148;
  int max_smem_optin = 0;
  // footprints
  float *d_pos = nullptr, *d_sigma = nullptr;
  int* d_rng = nullptr;
  float2* d_tab[3] = {nullptr, nullptr, nullptr};
  float2* d_tab_dpos[3] = {nullptr, nullptr, nullptr};  // extension: d/dpos, d/dsigma tables (dnmf_ext_enable)
  float2* d_tab_dsig[3] = {nullptr, nullptr, nullptr};
  float* d_resid = nullptr;
  size_t resid_cap = 0;
  double* d_sumr = nullptr;
  size_t sumr_cap = 0;
  bool have_footprints = false;
  float cutoff = 0.f;
  // tiling
  int nwx = 1, nwy = 1, tz = 0, cap = 0, user_cap = 0;
  int sub = 1;  // y-adjacent sub-tiles per warp (fit kernel only)
  bool auto_tiling = true;  // until dnmf_set_tiling is called: pick the warp layout from the list lengths
  int tx = 8, ty = 4, ntx = 0, nty = 0, ntz = 0;
  int wmax[3] = {0, 0, 0};
  int lmax_identity = 0;
  double mean_list_identity = 0.0;
  size_t fit_smem = 0;
  int fast_div = 0;
  float rcp[3] = {0.f, 0.f, 0.f};
  long long* d_cand_off = nullptr;  // static candidate lists per tile (identity windows +- cand_expand)
  int* d_cand_ids = nullptr;
  int cand_expand = 6;
  int cand_cap = 0;
  int fpc_override = 0;  // DNMF_FPC environment override of the frames-per-CTA heuristic (tuning)
  int y_pitch = 0, z_skew = 0;  // shared-memory layout of the Y tile (bank conflicts, configure_tiling_fixed)
  // tensor map of the frame buffer the fused kernel last ran on (resident slab or caller's batch)
  alignas(64) CUtensorMap tmap;
  const float* tmap_ptr = nullptr;
  long long tmap_frames = -1;
  int tmap_tx = 0, tmap_ty = 0;
  bool tmap_valid = false;
  void* encode_tiled = nullptr;
  // video
  float* d_video = nullptr;
  // scratch
  float* d_partials = nullptr;
  size_t partials_cap = 0;
  float* d_grad = nullptr;  // [10][3][T]
  double* d_sse = nullptr;
  size_t sse_cap = 0;
  float* d_batch = nullptr;
  size_t batch_cap = 0;
  int* d_ids = nullptr;
  size_t ids_cap = 0;
  double* d_loss = nullptr;
  int* d_tmp_counts = nullptr;
  size_t tmp_counts_cap = 0;
  long long* d_tmp_offsets = nullptr;
  int* d_tmp_max = nullptr;
  float* d_identity_beta = nullptr;
  int* d_ids_zero = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  // mu statistics
  double* d_G = nullptr;  // [T][K][K]
  double* d_b = nullptr;  // [T][K]
  double* d_Cd[2] = {nullptr, nullptr};  // traces in fp64 during the sweeps, [T][K]
  int cd_cur = 0;
  // sparse sweeps: static neighbour lists (neurons whose truncated supports overlap) and G compacted to them
  int* d_mu_nbr = nullptr;     // [K][mu_nbrw], ascending, -1 padded
  int mu_nbrw = 0;             // 0: lists not worth it (cutoff off / dense overlap) -> dense sweeps
  bool mu_nbr_built = false;
  double* d_Gc = nullptr;      // [T][K][mu_nbrw]
  size_t gc_cap = 0;
  bool gc_valid = false;       // compacted copy matches d_G
  int mu_dense_sweeps = 0;     // dnmf_mu_path flag bit 1 / DNMF_MU_DENSE_SWEEPS
  int mu_last_sparse = 0;
  int mu_sweep_per_launch = 0; // DNMF_MU_SWEEP_PER_LAUNCH / dnmf_mu_path bit 2: one launch per sweep even without coupling
  int mu_block4 = 0;           // DNMF_MU_BLOCK4: keep the 4x4 register blocks of the panel kernel for every list length
  int mu_capM = 0;
  // adaptive main-loop variant (FitParams::dyn_tail): restage counts of the previous fused launch
  unsigned* d_restage = nullptr;   // [32]
  unsigned* h_restage = nullptr;   // pinned [32], refreshed asynchronously after every fused launch
  long long restage_den_pending = 0;  // tile-frames of the launch the pending copy of the counters describes
  cudaEvent_t ev_restage = nullptr;
  int dyn_tail_mode = 0;           // DNMF_DYN_TAIL: 0 (default) / 1 force a variant, -1 automatic from the counters
  int dyn_tail_cur = 0;
  // frame-parallel epoch (dnmf_motion_epoch)
  int* d_epoch_batch_of = nullptr;
  size_t epoch_batch_of_cap = 0;
  int* d_epoch_offsets = nullptr;
  size_t epoch_offsets_cap = 0;
  float2* d_epoch_scalars = nullptr;
  size_t epoch_scalars_cap = 0;
  double* d_epoch_scale = nullptr;
  size_t epoch_scale_cap = 0;
  int epoch_sequential = 0;     // DNMF_EPOCH_SEQUENTIAL / dnmf_epoch_mode: one launch sequence per minibatch
  int epoch_last_parallel = 0;  // what the last dnmf_motion_epoch did
  int mu_force_panel = 0;  // dnmf_mu_path / DNMF_MU_PANEL: skip the fused-tile statistics kernel
  int mu_last_path = 0;    // 1 = fused tiles, 0 = panel kernel
  int mu_fused_need = 0;   // longest list seen by an overflowing fused-tile statistics launch (capacity hint)
  int mu_fused_off = 0;    // that capacity does not fit in shared memory: go straight to the panel kernel
  unsigned long long* d_keys = nullptr;
  size_t keys_cap = 0;
  int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static Geom geom_of(const dnmf_ctx* c) {
  Geom g;
  g.X = c->X;
  g.Y = c->Y;
  g.Z = c->Z;
  g.K = c->K;
  g.T = c->T;
  g.tx = c->tx;
  g.ty = c->ty;
  g.tz = c->tz;
  g.ntx = c->ntx;
  g.nty = c->nty;
  g.ntz = c->ntz;
  return g;
}

template <typename T>
static int ensure(T** ptr, size_t* cap, size_t need) {
  if (*cap >= need && *ptr) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  size_t want = need + need / 4;
  CU(cudaMalloc((void**)ptr, want * sizeof(T)));
  *cap = want;
  return 0;
}

extern "C" int dnmf_abi_version(void) { return DNMF_ABI_VERSION; }
extern "C" const char* dnmf_last_error(void) { return g_err.c_str(); }

extern "C" int dnmf_create(dnmf_ctx** out, int X, int Y, int Z, int K, int T, int device) {
  if (!out) return fail("dnmf_create: out is NULL");
  if (X < 1 || Y < 1 || Z < 1 || K < 1 || T < 1) return fail("dnmf_create: sizes must be positive");
  if (K > 65535) return fail("dnmf_create: K > 65535 not supported (uint16 neuron lists)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(std::string("dnmf_create: no CUDA device (") + cudaGetErrorString(e) +
                "); this library has no CPU fallback");
  CU(cudaSetDevice(device));
  dnmf_ctx* c = new dnmf_ctx();
  c->X = X;
  c->Y = Y;
  c->Z = Z;
  c->K = K;
  c->T = T;
  c->device = device;
  c->N = (size_t)X * Y * Z;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (const char* ev = getenv("DNMF_FPC")) c->fpc_override = atoi(ev);
  if (const char* ev = getenv("DNMF_DYN_TAIL")) c->dyn_tail_mode = atoi(ev);
  if (const char* ev = getenv("DNMF_MU_PANEL")) c->mu_force_panel = atoi(ev) != 0;
  if (const char* ev = getenv("DNMF_MU_SWEEP_PER_LAUNCH")) c->mu_sweep_per_launch = atoi(ev) != 0;
  if (const char* ev = getenv("DNMF_MU_BLOCK4")) c->mu_block4 = atoi(ev) != 0;
  if (const char* ev = getenv("DNMF_MU_DENSE_SWEEPS")) c->mu_dense_sweeps = atoi(ev) != 0;
  if (const char* ev = getenv("DNMF_EPOCH_SEQUENTIAL")) c->epoch_sequential = atoi(ev) != 0;
  CU(cudaMalloc((void**)&c->d_pos, (size_t)K * 3 * sizeof(float)));
  CU(cudaMalloc((void**)&c->d_sigma, (size_t)K * sizeof(float)));
  CU(cudaMalloc((void**)&c->d_rng, (size_t)K * 6 * sizeof(int)));
  const int s[3] = {X, Y, Z};
  for (int d = 0; d < 3; ++d) CU(cudaMalloc((void**)&c->d_tab[d], (size_t)K * (s[d] + 3) * sizeof(float2)));
  CU(cudaMalloc((void**)&c->d_grad, (size_t)30 * T * sizeof(float)));
  CU(cudaMemset(c->d_grad, 0, (size_t)30 * T * sizeof(float)));
  CU(cudaMalloc((void**)&c->d_loss, sizeof(double)));
  CU(cudaMalloc((void**)&c->d_tmp_max, sizeof(int)));
  CU(cudaMalloc((void**)&c->d_identity_beta, 30 * sizeof(float)));
  float idb[30];
  memset(idb, 0, sizeof(idb));
  idb[1 * 3 + 0] = idb[2 * 3 + 1] = idb[3 * 3 + 2] = 1.f;
  CU(cudaMemcpy(c->d_identity_beta, idb, sizeof(idb), cudaMemcpyHostToDevice));
  CU(cudaMalloc((void**)&c->d_ids_zero, sizeof(int)));
  CU(cudaMemset(c->d_ids_zero, 0, sizeof(int)));
  // enable the 3-instruction exact division only after an exhaustive device-side proof per axis size
  {
    static std::mutex mu;
    static std::vector<std::pair<int, int>> proven;  // (s-1, ok)
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long* d_bad = nullptr;
    CU(cudaMalloc((void**)&d_bad, sizeof(unsigned long long)));
    int all_ok = 1;
    for (int d = 0; d < 3; ++d) {
      const int sm1 = s[d] - 1;
      c->rcp[d] = sm1 > 0 ? (float)(1.0 / (double)sm1) : 0.f;
      int ok = -1;
      for (auto& pr : proven)
        if (pr.first == sm1) ok = pr.second;
      if (ok < 0) {
        if (sm1 <= 0) {
          ok = 0;
        } else {
          CU(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
          verify_coord_kernel<<<65536, 256>>>((float)sm1, c->rcp[d], d_bad);
          CU(cudaGetLastError());
          unsigned long long bad = 1;
          CU(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
          ok = bad == 0 ? 1 : 0;
        }
        proven.push_back({sm1, ok});
      }
      all_ok = all_ok && ok;
    }
    cudaFree(d_bad);
    c->fast_div = all_ok;
  }
  *out = c;
  return 0;
}

extern "C" void dnmf_destroy(dnmf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  void* ptrs[] = {c->d_pos,     c->d_sigma, c->d_rng,        c->d_tab[0],      c->d_tab[1],
                  c->d_tab[2],  c->d_video, c->d_partials,   c->d_grad,        c->d_sse,
                  c->d_batch,   c->d_ids,   c->d_loss,       c->d_tmp_counts,  c->d_tmp_offsets,
                  c->d_tmp_max, c->d_G,     c->d_b,          c->d_identity_beta,
                  c->d_Cd[0],   c->d_Cd[1], c->d_keys,       c->d_cand_off,    c->d_cand_ids,
                  c->d_tab_dpos[0], c->d_tab_dpos[1], c->d_tab_dpos[2], c->d_tab_dsig[0], c->d_tab_dsig[1],
                  c->d_tab_dsig[2], c->d_resid, c->d_sumr, c->d_ids_zero,
                  c->d_epoch_batch_of, c->d_epoch_offsets, c->d_epoch_scalars, c->d_epoch_scale,
                  c->d_mu_nbr, c->d_Gc, c->d_restage};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(c->ev_copied[i]);
      cudaEventDestroy(c->ev_done[i]);
    }
  }
  if (c->ev_restage) {
    cudaEventSynchronize(c->ev_restage);  // a copy into the pinned counters may still be in flight
    cudaEventDestroy(c->ev_restage);
  }
  if (c->h_restage) cudaFreeHost(c->h_restage);
  delete c;
}

// ---- tiling -----------------------------------------------------------------------------------
static int configure_tiling(dnmf_ctx* c, cudaStream_t st);

extern "C" int dnmf_set_tiling(dnmf_ctx* c, int warps_x, int warps_y, int tz, int slot_capacity, int subtiles_y) {
  if (!c) return fail("dnmf_set_tiling: ctx is NULL");
  if (subtiles_y < 1) subtiles_y = 1;
  if (subtiles_y > 2 || (subtiles_y == 2 && warps_y > 2))
    return fail("dnmf_set_tiling: subtiles_y = 2 is supported for the 1x1, 2x1 and 2x2 warp layouts");
  const bool ok = (warps_x == 1 && warps_y == 1) || (warps_x == 2 && warps_y == 1) ||
                  (warps_x == 2 && warps_y == 2) || (warps_x == 2 && warps_y == 4);
  if (!ok) return fail("dnmf_set_tiling: supported warp layouts are 1x1, 2x1, 2x2, 2x4");
  if (tz < 0) return fail("dnmf_set_tiling: tz must be >= 0");
  c->nwx = warps_x;
  c->nwy = warps_y;
  c->tz = tz;
  c->user_cap = slot_capacity;
  c->sub = subtiles_y;
  c->auto_tiling = false;
  CU(cudaSetDevice(c->device));
  if (c->have_footprints) return configure_tiling(c, 0);
  return 0;
}

extern "C" int dnmf_get_tiling(dnmf_ctx* c, int32_t* out) {
  if (!c || !out) return fail("dnmf_get_tiling: NULL argument");
  int32_t v[11] = {c->tx, c->ty, c->tz, c->ntx, c->nty, c->ntz, c->nwx, c->nwy, c->cap, c->sub, c->fast_div};
  memcpy(out, v, sizeof(v));
  return 0;
}

static int run_bin_count(dnmf_ctx* c, const float* beta, int beta_T, const int* ids, int B, int* counts,
                         int* windows, cudaStream_t st, int expand = 0) {
  Geom g = geom_of(c);
  g.T = beta_T;
  const long long items = (long long)B * g.ntx * g.nty * g.ntz;
  const int wpb = 8;
  bin_tiles_kernel<false><<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, 0, st>>>(
      g, beta, ids, B, c->d_rng, counts, nullptr, windows, nullptr, 0, expand);
  CU(cudaGetLastError());
  return 0;
}

static int configure_tiling_fixed(dnmf_ctx* c, cudaStream_t st);

// Short lists (cfg1-3: a handful of neurons per tile) want the smallest CTA footprint (one warp, tightest
// lists); dense configurations (cfg4: ~100 neurons per tile) want several warps sharing one staged copy
// of the table slices, or shared memory caps occupancy at a few warps per SM.
static int configure_tiling(dnmf_ctx* c, cudaStream_t st) {
  if (!c->auto_tiling) return configure_tiling_fixed(c, st);
  // Instruction-count model of the fused kernel per 32-voxel row (from the ncu source pages, profiles/):
  // ~85 fixed + ~9.5 per listed neuron (one sub-tile per warp; the packed two-sub-tile march: 42 + 7.3) + the tile
  // prologue/epilogue amortised over the rows of the tile, inflated when shared memory leaves too few warps per SM.
  static const int layouts[7][3] = {{1, 1, 2}, {1, 1, 1}, {2, 1, 2}, {2, 2, 2}, {2, 1, 1}, {2, 2, 1}, {2, 4, 1}};
  int best = 0;
  double best_cost = 1e300;
  for (int i = 0; i < 7; ++i) {
    c->nwx = layouts[i][0];
    c->nwy = layouts[i][1];
    c->sub = layouts[i][2];
    if (configure_tiling_fixed(c, st)) return 1;
    const int nw = c->nwx * c->nwy;
    const int ctas = std::min(32, (int)((size_t)227 * 1024 / (c->fit_smem + 1024)));
    const double warps = std::min(64, ctas * nw);
    const double rows = (double)c->sub * c->tz;
    // two sub-tiles per warp run the packed (A, B) march: ~42 fixed + ~7.3 per listed neuron per row
    const double fixed = c->sub == 2 ? 42.0 : 85.0, per = c->sub == 2 ? 7.3 : 9.5;
    // fewer than ~8 resident warps per SM cannot keep the FP32 pipe fed (cfg4, measured: 16 warps 1.0, 8 warps
    // 1.04, 6 warps 1.6, 4 warps 2.0 relative cost)
    // lanes past the volume edge still cost: padded volume / volume
    const double edge = ((double)c->ntx * c->tx / c->X) * ((double)c->nty * c->ty / c->Y);
    double cost = (fixed + per * c->mean_list_identity + (300.0 + 600.0 / nw) / rows) *
                  std::pow(std::max(1.0, 8.0 / warps), 1.5) * edge;
    if (nw > 1 && c->mean_list_identity < 16.0) cost *= 1.25;  // sharing the staged slices only pays for long lists
    if (cost < best_cost) {
      best_cost = cost;
      best = i;
    }
  }
  c->nwx = layouts[best][0];
  c->nwy = layouts[best][1];
  c->sub = layouts[best][2];
  return configure_tiling_fixed(c, st);
}

static int configure_tiling_fixed(dnmf_ctx* c, cudaStream_t st) {
  c->tx = kWarpX * c->nwx;
  c->ty = kWarpY * c->nwy * c->sub;
  if (c->tz <= 0 || c->tz > c->Z) c->tz = std::min(c->Z, 32);
  c->ntx = (c->X + c->tx - 1) / c->tx;
  c->nty = (c->Y + c->ty - 1) / c->ty;
  c->ntz = (c->Z + c->tz - 1) / c->tz;
  const int margin = 2;
  c->wmax[0] = std::min(c->tx + 2 + margin, c->X + 3);
  c->wmax[1] = std::min(c->ty + 2 + margin, c->Y + 3);
  c->wmax[2] = std::min(c->tz + 2 + margin, c->Z + 3);
  {
    // Y tile in shared memory: lane (lx, ly) of a warp reads float lx*pitch + ly*zs + z.  With the dense pitch
    // ty*zs the 32 lanes can fall on very few banks (Z = 32: all on one).  When the dense layout is worse than
    // 2-way, pad the x pitch to 4 (mod 32) and rotate the z order of the lanes by ly*(1-zs) (mod 4): bank =
    // 4*lx + (ly mod 4) + const.  The dense layout is kept otherwise (it allows the single tensor-TMA copy).
    const int zs = c->tz;
    const int dense = c->ty * zs;
    int worst = 0, hist[32] = {0};
    for (int l = 0; l < 32; ++l) worst = std::max(worst, ++hist[((l & 7) * dense + (l >> 3) * zs) & 31]);
    c->y_pitch = dense;
    c->z_skew = 0;
    if (worst > 2) {
      int pitch = (dense + 3) & ~3;
      while ((pitch & 31) != 4) pitch += 4;
      c->y_pitch = pitch;
      c->z_skew = ((1 - zs) % 4 + 4) % 4;
    }
  }
  // longest list at identity deformation -> staged-slot capacity
  const int nt = c->ntx * c->nty * c->ntz;
  if (ensure(&c->d_tmp_counts, &c->tmp_counts_cap, (size_t)nt)) return 1;
  if (c->d_tmp_offsets) cudaFree(c->d_tmp_offsets);
  CU(cudaMalloc((void**)&c->d_tmp_offsets, ((size_t)nt + 1) * sizeof(long long)));
  int zero = 0;
  int* d_zero = nullptr;
  CU(cudaMalloc((void**)&d_zero, sizeof(int)));
  CU(cudaMemcpyAsync(d_zero, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  if (run_bin_count(c, c->d_identity_beta, 1, d_zero, 1, c->d_tmp_counts, nullptr, st)) return 1;
  scan_counts_kernel<<<1, 1024, 0, st>>>(c->d_tmp_counts, nt, c->d_tmp_offsets, c->d_tmp_max);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(&c->lmax_identity, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // static candidate lists: identity windows expanded by cand_expand nodes on every side
  {
    if (c->d_cand_off) cudaFree(c->d_cand_off);
    if (c->d_cand_ids) cudaFree(c->d_cand_ids);
    c->d_cand_off = nullptr;
    c->d_cand_ids = nullptr;
    CU(cudaMalloc((void**)&c->d_cand_off, ((size_t)nt + 1) * sizeof(long long)));
    if (run_bin_count(c, c->d_identity_beta, 1, d_zero, 1, c->d_tmp_counts, nullptr, st, c->cand_expand)) return 1;
    scan_counts_kernel<<<1, 1024, 0, st>>>(c->d_tmp_counts, nt, c->d_cand_off, nullptr);
    CU(cudaGetLastError());
    long long total = 0;
    CU(cudaMemcpyAsync(&total, c->d_cand_off + nt, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaMalloc((void**)&c->d_cand_ids, (size_t)std::max<long long>(total, 1) * sizeof(int)));
    {  // longest static candidate list -> shared-memory capacity of the fused kernel's candidate cache
      std::vector<long long> h_off((size_t)nt + 1);
      CU(cudaMemcpy(h_off.data(), c->d_cand_off, ((size_t)nt + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
      long long longest = 0;
      for (int i = 0; i < nt; ++i) longest = std::max(longest, h_off[(size_t)i + 1] - h_off[(size_t)i]);
      c->cand_cap = (int)std::min<long long>((longest + 1) & ~1LL, 1024);
    }
    Geom g1 = geom_of(c);
    g1.T = 1;
    const int wpb = 8;
    bin_tiles_kernel<true><<<(unsigned)((nt + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        g1, c->d_identity_beta, d_zero, 1, c->d_rng, c->d_tmp_counts, c->d_cand_off, nullptr, c->d_cand_ids, total,
        c->cand_expand);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
  }
  cudaFree(d_zero);
  // Staged-slot capacity: the smallest capacity that keeps all but ~2 % of the (tile, neuron) pairs of the
  // identity-deformation lists in shared memory (+1 slot of slack).  The few longest lists send their tail
  // through the L2-resident tables instead of forcing every CTA to reserve shared memory for the maximum.
  int cap = c->user_cap;
  std::vector<int> h_counts((size_t)nt);
  if (run_bin_count(c, c->d_identity_beta, 1, c->d_ids_zero, 1, c->d_tmp_counts, nullptr, st)) return 1;
  CU(cudaMemcpyAsync(h_counts.data(), c->d_tmp_counts, (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  long long total = 0;
  for (int v : h_counts) total += v;
  c->mean_list_identity = nt > 0 ? (double)total / nt : 0.0;
  if (cap <= 0) {
    cap = 2;
    for (;; cap += 1) {
      long long over = 0;
      for (int v : h_counts) over += std::max(0, v - cap);
      if (over * 50 <= total || cap >= c->lmax_identity) break;
    }
    cap += 1;
  }
  cap = std::max(2, std::min(cap, c->K + 1));
  cap = (cap + 1) & ~1;               // slots are consumed in pairs (LDS.128)
  if ((cap & 3) == 0) cap += 2;       // slot-row stride = 2 (mod 4) float2: spreads entries over banks
  const int wsum = c->wmax[0] + c->wmax[1] + c->wmax[2];
  const int nw = c->nwx * c->nwy;
  // keep at least ~2 CTAs per SM worth of shared memory when possible
  const size_t budget = std::min<size_t>((size_t)c->max_smem_optin, (size_t)113 * 1024);
  while (cap > 2 && fit_smem_layout(nw, c->tx, c->ty, c->tz, cap, wsum, c->K, c->wmax[0], c->cand_cap, c->y_pitch).bytes > budget)
    cap -= 4;
  if (cap < 2) cap = 2;
  c->cap = cap;
  c->fit_smem = fit_smem_layout(nw, c->tx, c->ty, c->tz, cap, wsum, c->K, c->wmax[0], c->cand_cap, c->y_pitch).bytes;
  if (c->fit_smem > (size_t)c->max_smem_optin)
    return fail("configure_tiling: tile does not fit in shared memory; use a smaller tz");
  return 0;
}

// ---- footprints ---------------------------------------------------------------------------------
extern "C" int dnmf_set_footprints(dnmf_ctx* c, const float* pos_host, const float* sigma_host,
                                   float cutoff, void* stream) {
  if (!c || !pos_host || !sigma_host) return fail("dnmf_set_footprints: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->d_pos, pos_host, (size_t)c->K * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(c->d_sigma, sigma_host, (size_t)c->K * sizeof(float), cudaMemcpyHostToDevice, st));
  c->cutoff = cutoff;
  build_ranges_kernel<<<(c->K + 127) / 128, 128, 0, st>>>(c->d_pos, c->d_sigma, c->K, c->X, c->Y, c->Z,
                                                          cutoff, c->d_rng);
  CU(cudaGetLastError());
  const int s[3] = {c->X, c->Y, c->Z};
  for (int d = 0; d < 3; ++d) {
    dim3 grid((s[d] + 3 + 127) / 128, c->K);
    build_tables_kernel<<<grid, 128, 0, st>>>(c->d_pos, c->d_sigma, c->d_rng, c->K, s[d], d, c->d_tab[d],
                                              c->d_tab_dpos[d], c->d_tab_dsig[d]);
    CU(cudaGetLastError());
  }
  c->have_footprints = true;
  c->mu_capM = 0;
  c->mu_nbr_built = false;
  c->gc_valid = false;
  c->mu_fused_need = 0;
  c->mu_fused_off = 0;
  c->counters[3]++;
  return configure_tiling(c, st);
}

extern "C" int dnmf_get_ranges(dnmf_ctx* c, int32_t* out) {
  if (!c || !out) return fail("dnmf_get_ranges: NULL argument");
  if (!c->have_footprints) return fail("dnmf_get_ranges: call dnmf_set_footprints first");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(out, c->d_rng, (size_t)c->K * 6 * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int dnmf_get_table(dnmf_ctx* c, int axis, float* out) {
  if (!c || !out || axis < 0 || axis > 2) return fail("dnmf_get_table: bad argument");
  if (!c->have_footprints) return fail("dnmf_get_table: call dnmf_set_footprints first");
  const int s[3] = {c->X, c->Y, c->Z};
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(out, c->d_tab[axis], (size_t)c->K * (s[axis] + 3) * sizeof(float2), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- video --------------------------------------------------------------------------------------
extern "C" int dnmf_upload_frames(dnmf_ctx* c, const float* frames_host, int t0, int n, int clamp_negative,
                                  void* stream) {
  if (!c || !frames_host) return fail("dnmf_upload_frames: NULL argument");
  if (t0 < 0 || n < 0 || t0 + n > c->T) return fail("dnmf_upload_frames: frame range outside [0,T)");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (!c->d_video) CU(cudaMalloc((void**)&c->d_video, c->N * (size_t)c->T * sizeof(float)));
  float* dst = c->d_video + (size_t)t0 * c->N;
  CU(cudaMemcpyAsync(dst, frames_host, c->N * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (clamp_negative && n > 0) {
    clamp_negative_kernel<<<c->num_sms * 8, 256, 0, st>>>(dst, c->N * (size_t)n);
    CU(cudaGetLastError());
  }
  return 0;
}

extern "C" int dnmf_video_devptr(dnmf_ctx* c, float** out) {
  if (!c || !out) return fail("dnmf_video_devptr: NULL argument");
  CU(cudaSetDevice(c->device));
  if (!c->d_video) CU(cudaMalloc((void**)&c->d_video, c->N * (size_t)c->T * sizeof(float)));
  *out = c->d_video;
  return 0;
}

// ---- stand-alone binning ---------------------------------------------------------------------------
extern "C" int dnmf_bin_tiles(dnmf_ctx* c, const float* beta_dev, const int32_t* frame_ids_dev, int B,
                              int32_t* counts_dev, int64_t* offsets_dev, int32_t* windows_dev,
                              int32_t* ids_dev, int64_t ids_capacity, int64_t* total_host, void* stream) {
  if (!c || !beta_dev || !frame_ids_dev || !counts_dev || !offsets_dev)
    return fail("dnmf_bin_tiles: NULL argument");
  if (!c->have_footprints) return fail("dnmf_bin_tiles: call dnmf_set_footprints first");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  const long long items = (long long)B * c->ntx * c->nty * c->ntz;
  if (run_bin_count(c, beta_dev, c->T, frame_ids_dev, B, counts_dev, windows_dev, st)) return 1;
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts_dev, items, (long long*)offsets_dev, nullptr);
  CU(cudaGetLastError());
  if (ids_dev && ids_capacity > 0) {
    Geom g = geom_of(c);
    const int wpb = 8;
    bin_tiles_kernel<true><<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        g, beta_dev, frame_ids_dev, B, c->d_rng, counts_dev, (const long long*)offsets_dev, nullptr, ids_dev,
        ids_capacity, 0);
    CU(cudaGetLastError());
  }
  if (total_host) {
    long long tot = 0;
    CU(cudaMemcpyAsync(&tot, offsets_dev + items, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *total_host = tot;
  }
  c->counters[2]++;
  return 0;
}

// ---- fused step -----------------------------------------------------------------------------------
template <int NWX, int NWY, int SUB, int MD_, bool FD_>
static int launch_fit(const FitParams& p0, int B, size_t smem, cudaStream_t st) {
  auto kern = fit_tile_kernel<NWX, NWY, SUB, MD_, FD_>;
  static size_t configured[64] = {0};  // per device: the attribute is a per-device property of the function
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  // grid = (ntx, nty, chunks*ntz), one chunk = fpc consecutive frames; gridDim.z <= 65535, so very large
  // batches go out in several launches
  const int fpc = std::max(1, std::min(p0.fpc, 32));
  const int maxB = std::max(1, 65535 / p0.ntz) * fpc;
  const size_t N = (size_t)p0.X * p0.Y * p0.Z;
  const size_t nt = (size_t)p0.ntx * p0.nty * p0.ntz;
  for (int b0 = 0; b0 < B; b0 += maxB) {
    const int nb = std::min(maxB, B - b0);
    FitParams p = p0;
    p.fpc = fpc;
    p.B = nb;
    p.b_base = b0;
    p.frame_ids = p0.frame_ids + b0;
    p.partials = p0.partials + (size_t)b0 * nt * kNumPartials;
    if (p0.frames_are_batch) p.frames = p0.frames + (size_t)b0 * N;
    if (p0.yhat) p.yhat = p0.yhat + (size_t)b0 * N;
    const int chunks = (nb + fpc - 1) / fpc;
    dim3 grid((unsigned)p0.ntx, (unsigned)p0.nty, (unsigned)(chunks * p0.ntz));
    kern<<<grid, 32 * NWX * NWY, smem, st>>>(p);
    CU(cudaGetLastError());
  }
  return 0;
}

template <int MD_>
static int dispatch_fit(dnmf_ctx* c, const FitParams& p, int B, cudaStream_t st) {
  const size_t smem = c->fit_smem;
  if (c->nty > 65535) return fail("dispatch_fit: more than 65535 tiles along y");
  const bool fd = c->fast_div != 0;
#define DNMF_DISPATCH(a, b, sb)                                                       \
  if (c->nwx == a && c->nwy == b && c->sub == sb)                                      \
    return fd ? launch_fit<a, b, sb, MD_, true>(p, B, smem, st) : launch_fit<a, b, sb, MD_, false>(p, B, smem, st);
  DNMF_DISPATCH(1, 1, 1)
  DNMF_DISPATCH(1, 1, 2)
  DNMF_DISPATCH(2, 1, 1)
  DNMF_DISPATCH(2, 1, 2)
  DNMF_DISPATCH(2, 2, 1)
  DNMF_DISPATCH(2, 2, 2)
  DNMF_DISPATCH(2, 4, 1)
#undef DNMF_DISPATCH
  return fail("dispatch_fit: unsupported warp layout");
}

// MODE 3 (trace statistics) exists for the two-sub-tile layouts with the verified fast division only
static bool fused_stats_available(const dnmf_ctx* c) {
  return c->sub == 2 && c->fast_div != 0 && ((c->nwx == 1 && c->nwy == 1) || (c->nwx == 2 && c->nwy <= 2));
}
static int dispatch_stats(dnmf_ctx* c, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  if (c->nty > 65535) return fail("dispatch_stats: more than 65535 tiles along y");
  if (!fused_stats_available(c)) return fail("dispatch_stats: layout without a fused statistics kernel");
  if (c->nwx == 1) return launch_fit<1, 1, 2, 3, true>(p, B, smem, st);
  if (c->nwy == 1) return launch_fit<2, 1, 2, 3, true>(p, B, smem, st);
  return launch_fit<2, 2, 2, 3, true>(p, B, smem, st);
}

static int fill_fit_params(dnmf_ctx* c, FitParams& p, const float* frames_dev, const int32_t* ids, int B,
                           const float* beta, const float* C) {
  if (!c->have_footprints) return fail("fit: call dnmf_set_footprints first");
  if (!frames_dev && !c->d_video) return fail("fit: no resident video (dnmf_upload_frames) and frames_dev is NULL");
  if ((long long)B * c->ntx * c->nty * c->ntz > 2147483647LL) return fail("fit: too many tiles in one launch");
  p.frames = frames_dev ? frames_dev : c->d_video;
  p.frames_are_batch = frames_dev ? 1 : 0;
  p.frame_ids = ids;
  p.beta = beta;
  p.C = C;
  p.tab0 = c->d_tab[0];
  p.tab1 = c->d_tab[1];
  p.tab2 = c->d_tab[2];
  p.rng = c->d_rng;
  p.X = c->X;
  p.Y = c->Y;
  p.Z = c->Z;
  p.K = c->K;
  p.T = c->T;
  p.tz = c->tz;
  p.ntx = c->ntx;
  p.nty = c->nty;
  p.ntz = c->ntz;
  p.cap = c->cap;
  p.wmax0 = c->wmax[0];
  p.wmax1 = c->wmax[1];
  p.wmax2 = c->wmax[2];
  p.full_depth = (c->tz == c->Z) ? 1 : 0;
  p.bulk_ok = (((uintptr_t)p.frames & 15) == 0) && (((size_t)c->Y * c->Z) % 4 == 0) &&
              (((size_t)c->ty * c->Z) % 4 == 0);
  p.muG = nullptr;
  p.mub = nullptr;
  p.mu_overflow = nullptr;
  p.dyn_tail = 0;
  p.restage_count = nullptr;
  p.y_pitch = c->y_pitch;
  p.z_skew = c->z_skew;
  p.b_base = 0;
  p.tmap_ok = 0;
  memset(&p.tmap, 0, sizeof(p.tmap));
  if (p.bulk_ok && p.full_depth && (size_t)c->ty * c->Z <= 256 && c->tx <= 256 && c->y_pitch == c->ty * c->Z) {
    // 3-D tensor map over the frame buffer: [frames][X][Y*Z] floats, box = one tile (tx rows of ty*Z floats)
    const long long nframes = frames_dev ? (long long)B : (long long)c->T;
    if (c->tmap_ptr != p.frames || c->tmap_frames != nframes || c->tmap_tx != c->tx || c->tmap_ty != c->ty) {
      c->tmap_valid = false;
      if (!c->encode_tiled) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
          c->encode_tiled = fn;
        else
          (void)cudaGetLastError();
      }
      if (c->encode_tiled) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        const cuuint64_t gdim[3] = {(cuuint64_t)c->Y * c->Z, (cuuint64_t)c->X, (cuuint64_t)nframes};
        const cuuint64_t gstride[2] = {(cuuint64_t)c->Y * c->Z * 4, (cuuint64_t)c->N * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(c->ty * c->Z), (cuuint32_t)c->tx, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = ((EncodeFn)c->encode_tiled)(
            &c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.frames), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->tmap_valid = (r == CUDA_SUCCESS);
      }
      c->tmap_ptr = p.frames;
      c->tmap_frames = nframes;
      c->tmap_tx = c->tx;
      c->tmap_ty = c->ty;
    }
    if (c->tmap_valid) {
      p.tmap = c->tmap;
      p.tmap_ok = 1;
    }
  }
  p.fast_div = c->fast_div;
  p.rcp0 = c->rcp[0];
  p.rcp1 = c->rcp[1];
  p.rcp2 = c->rcp[2];
  p.cand_off = c->d_cand_off;
  p.cand_ids = c->d_cand_ids;
  p.cand_expand = c->cand_expand;
  p.cand_cap = c->cand_cap;
  p.B = B;
  {  // frames per CTA: as many as keep ~8 waves of CTAs in flight, at most 8 (measured: 8 = 341k, 16 = 338k, 1 = 320k frame-iters/s at cfg2)
    const long long tiles = (long long)c->ntx * c->nty * c->ntz;
    const long long slots = (long long)c->num_sms * 16 * 8;
    p.fpc = (int)std::max<long long>(1, std::min<long long>(8, (long long)B * tiles / std::max<long long>(slots, 1)));
    if (c->fpc_override > 0) p.fpc = std::min(c->fpc_override, 32);
  }
  p.yhat = nullptr;
  p.bg = 0.f;
  const size_t need = (size_t)B * c->ntx * c->nty * c->ntz * kNumPartials;
  if (ensure(&c->d_partials, &c->partials_cap, need)) return 1;
  p.partials = c->d_partials;
  return 0;
}

// Fused fit launch (MODE 0) with the main-loop variant picked from what the previous launches did: the kernel
// counts the tile-frames whose slices were (re)built, the counters come back through pinned memory without a
// synchronisation (read one or two launches late), and above half of all tile-frames the single-body main loop
// is used (march_rolled TAIL 3).  Both variants execute the same arithmetic: results do not depend on the choice.
static int launch_fused_fit(dnmf_ctx* c, FitParams& p, int B, cudaStream_t st) {
  if (!c->d_restage) {
    CU(cudaMalloc((void**)&c->d_restage, 32 * sizeof(unsigned)));
    CU(cudaMemset(c->d_restage, 0, 32 * sizeof(unsigned)));
    CU(cudaMallocHost((void**)&c->h_restage, 32 * sizeof(unsigned)));
    CU(cudaEventCreateWithFlags(&c->ev_restage, cudaEventDisableTiming));
  }
  bool can_enqueue = true;
  if (c->restage_den_pending > 0) {
    const cudaError_t q = cudaEventQuery(c->ev_restage);
    if (q == cudaSuccess) {
      unsigned long long sum = 0;
      for (int i = 0; i < 32; ++i) sum += c->h_restage[i];
      const double frac = (double)sum / (double)c->restage_den_pending;
      if (frac > 0.5) c->dyn_tail_cur = 1;
      else if (frac < 0.35) c->dyn_tail_cur = 0;
      c->restage_den_pending = 0;
    } else if (q == cudaErrorNotReady) {
      can_enqueue = false;  // the pinned buffer is still owed a copy
    } else {
      CU(q);
    }
  }
  p.dyn_tail = c->dyn_tail_mode >= 0 ? (c->dyn_tail_mode != 0) : c->dyn_tail_cur;
  p.restage_count = c->d_restage;
  if (dispatch_fit<0>(c, p, B, st)) return 1;
  if (can_enqueue) {
    CU(cudaMemcpyAsync(c->h_restage, c->d_restage, 32 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(c->d_restage, 0, 32 * sizeof(unsigned), st));
    CU(cudaEventRecord(c->ev_restage, st));
    c->restage_den_pending = (long long)B * c->ntx * c->nty * c->ntz;
  }
  return 0;
}

extern "C" int dnmf_loss_grad(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                              int B_global, const float* beta_dev, const float* C_dev, float* grad_dev,
                              double* sse_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !C_dev || !grad_dev || !sse_dev)
    return fail("dnmf_loss_grad: NULL argument");
  if (B < 1 || B_global < B) return fail("dnmf_loss_grad: need 1 <= B <= B_global");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  FitParams p;
  if (fill_fit_params(c, p, frames_dev, frame_ids_dev, B, beta_dev, C_dev)) return 1;
  const int nt = c->ntx * c->nty * c->ntz;
  if (launch_fused_fit(c, p, B, st)) return 1;
  const double scale = 2.0 / ((double)B_global * (double)c->N);
  reduce_partials_kernel<<<B, 256, 0, st>>>(c->d_partials, frame_ids_dev, nt, c->T, scale, grad_dev, sse_dev, nullptr);
  CU(cudaGetLastError());
  c->counters[0] += 1;  // fused launches
  c->counters[1] += 1;  // reduce launches
  return 0;
}

extern "C" int dnmf_adam_step(dnmf_ctx* c, float* beta_dev, float* grad_dev, float* m_dev, float* v_dev,
                              double lr, double beta1, double beta2, double eps, int64_t step, int affine,
                              const double* sse_dev, int B, int B_global, double* loss_dev, void* stream) {
  if (!c || !beta_dev || !grad_dev || !m_dev || !v_dev) return fail("dnmf_adam_step: NULL argument");
  if (step < 1) return fail("dnmf_adam_step: step is 1-based");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  AdamParams a;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  a.w1 = (float)(1.0 - beta1);
  a.b2 = (float)beta2;
  a.w2 = (float)(1.0 - beta2);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.eps = (float)eps;
  const int n = 30 * c->T;
  const double loss_scale = 1.0 / ((double)(B_global > 0 ? B_global : 1) * (double)c->N);
  adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(beta_dev, grad_dev, m_dev, v_dev, n, 3 * c->T, affine, a,
                                               sse_dev, B, loss_scale, (sse_dev && loss_dev) ? loss_dev : nullptr);
  CU(cudaGetLastError());
  c->counters[4] += 1;
  return 0;
}

extern "C" int dnmf_motion_step(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                                int B_global, float* beta_dev, float* m_dev, float* v_dev, const float* C_dev,
                                double lr, double beta1, double beta2, double eps, int64_t step, int affine,
                                double* loss_dev, void* stream) {
  if (!c) return fail("dnmf_motion_step: ctx is NULL");
  CU(cudaSetDevice(c->device));
  if (ensure(&c->d_sse, &c->sse_cap, (size_t)B)) return 1;
  if (dnmf_loss_grad(c, frames_dev, frame_ids_dev, B, B_global, beta_dev, C_dev, c->d_grad, c->d_sse, stream))
    return 1;
  return dnmf_adam_step(c, beta_dev, c->d_grad, m_dev, v_dev, lr, beta1, beta2, eps, step, affine, c->d_sse, B,
                        B_global, loss_dev, stream);
}

extern "C" int dnmf_motion_epoch(dnmf_ctx* c, const int32_t* frame_ids_dev, const int32_t* batch_offsets_host,
                                 int nbatches, int global_batch_scale, float* beta_dev, float* m_dev, float* v_dev,
                                 const float* C_dev, double lr, double beta1, double beta2, double eps,
                                 int64_t first_step, int affine, double* loss_dev, void* stream) {
  if (!c || !frame_ids_dev || !batch_offsets_host) return fail("dnmf_motion_epoch: NULL argument");
  if (nbatches < 0 || global_batch_scale < 1) return fail("dnmf_motion_epoch: need nbatches >= 0, global_batch_scale >= 1");
  if (!c->d_video) return fail("dnmf_motion_epoch: no resident video (dnmf_upload_frames)");
  if (!beta_dev || !m_dev || !v_dev || !C_dev) return fail("dnmf_motion_epoch: NULL argument");
  if (first_step < 1) return fail("dnmf_motion_epoch: step is 1-based");
  if (nbatches == 0) return 0;
  for (int i = 0; i < nbatches; ++i)
    if (batch_offsets_host[i + 1] - batch_offsets_host[i] < 1) return fail("dnmf_motion_epoch: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  const int b_first = batch_offsets_host[0];
  const long long Btot = (long long)batch_offsets_host[nbatches] - b_first;
  // ---- frame-parallel epoch (see epoch_adam_kernel): every frame at most once, one fused launch ----
  bool parallel = nbatches > 1 && c->epoch_sequential == 0 && Btot <= c->T &&
                  Btot * c->ntx * c->nty * c->ntz <= 2147483647LL;
  if (parallel) {
    if (ensure(&c->d_epoch_batch_of, &c->epoch_batch_of_cap, (size_t)c->T + 1)) return 1;  // [T] + status
    if (ensure(&c->d_epoch_offsets, &c->epoch_offsets_cap, (size_t)nbatches + 1)) return 1;
    if (ensure(&c->d_epoch_scalars, &c->epoch_scalars_cap, (size_t)nbatches)) return 1;
    if (ensure(&c->d_epoch_scale, &c->epoch_scale_cap, (size_t)Btot)) return 1;
    if (ensure(&c->d_sse, &c->sse_cap, (size_t)Btot)) return 1;
    std::vector<int> off((size_t)nbatches + 1);
    for (int i = 0; i <= nbatches; ++i) off[(size_t)i] = batch_offsets_host[i] - b_first;
    std::vector<float2> sc((size_t)nbatches);
    for (int i = 0; i < nbatches; ++i) {
      const double step = (double)(first_step + i);
      sc[(size_t)i] = make_float2((float)(lr / (1.0 - pow(beta1, step))), (float)sqrt(1.0 - pow(beta2, step)));
    }
    int* status = c->d_epoch_batch_of + c->T;
    CU(cudaMemsetAsync(c->d_epoch_batch_of, 0xff, (size_t)c->T * sizeof(int), st));
    CU(cudaMemsetAsync(status, 0, sizeof(int), st));
    CU(cudaMemcpyAsync(c->d_epoch_offsets, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_epoch_scalars, sc.data(), sc.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    const int32_t* ids = frame_ids_dev + b_first;
    epoch_index_kernel<<<(unsigned)((Btot + 255) / 256), 256, 0, st>>>(ids, c->d_epoch_offsets, nbatches, c->T,
                                                                      (double)c->N, global_batch_scale,
                                                                      c->d_epoch_batch_of, c->d_epoch_scale, status);
    CU(cudaGetLastError());
    int h_status = 0;
    CU(cudaMemcpyAsync(&h_status, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));  // also: off / sc may go out of scope
    if (h_status & 2) return fail("dnmf_motion_epoch: frame id out of range");
    parallel = h_status == 0;  // a frame drawn twice in the epoch: its steps depend on each other
  }
  if (parallel) {
    const int n = 30 * c->T;
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    epoch_adam_kernel<<<(n + 127) / 128, 128, 0, st>>>(beta_dev, c->d_grad, m_dev, v_dev, n, c->T, affine, w1, b2, w2,
                                                       epsf, c->d_epoch_scalars, nbatches, c->d_epoch_batch_of, 0);
    CU(cudaGetLastError());
    FitParams p;
    if (fill_fit_params(c, p, nullptr, frame_ids_dev + b_first, (int)Btot, beta_dev, C_dev)) return 1;
    const int nt = c->ntx * c->nty * c->ntz;
    if (launch_fused_fit(c, p, (int)Btot, st)) return 1;
    reduce_partials_kernel<<<(unsigned)Btot, 256, 0, st>>>(c->d_partials, frame_ids_dev + b_first, nt, c->T, 0.0,
                                                          c->d_grad, c->d_sse, nullptr, c->d_epoch_scale);
    CU(cudaGetLastError());
    epoch_adam_kernel<<<(n + 127) / 128, 128, 0, st>>>(beta_dev, c->d_grad, m_dev, v_dev, n, c->T, affine, w1, b2, w2,
                                                       epsf, c->d_epoch_scalars, nbatches, c->d_epoch_batch_of, 1);
    CU(cudaGetLastError());
    if (loss_dev) {
      epoch_loss_kernel<<<(nbatches + 3) / 4, 128, 0, st>>>(c->d_sse, c->d_epoch_offsets, nbatches, (double)c->N,
                                                            global_batch_scale, loss_dev);
      CU(cudaGetLastError());
    }
    c->counters[0] += 1;
    c->counters[1] += 1;
    c->counters[4] += 2;
    c->epoch_last_parallel = 1;
    return 0;
  }
  c->epoch_last_parallel = 0;
  for (int i = 0; i < nbatches; ++i) {
    const int b0 = batch_offsets_host[i], B = batch_offsets_host[i + 1] - b0;
    if (dnmf_motion_step(c, nullptr, frame_ids_dev + b0, B, B * global_batch_scale, beta_dev, m_dev, v_dev, C_dev, lr,
                         beta1, beta2, eps, first_step + i, affine, loss_dev ? loss_dev + i : nullptr, stream))
      return 1;
  }
  return 0;
}

extern "C" int dnmf_epoch_mode(dnmf_ctx* c, int sequential, int* last_parallel_out) {
  if (!c) return fail("dnmf_epoch_mode: NULL context");
  if (sequential >= 0) c->epoch_sequential = sequential != 0;
  if (last_parallel_out) *last_parallel_out = c->epoch_last_parallel;
  return 0;
}

extern "C" int dnmf_motion_step_host(dnmf_ctx* c, const float* frames_host, const int32_t* frame_ids_host,
                                     int B, int B_global, float* beta_dev, float* m_dev, float* v_dev,
                                     const float* C_dev, double lr, double beta1, double beta2, double eps,
                                     int64_t step, int affine, double* loss_host, void* stream) {
  if (!c || !frames_host || !frame_ids_host) return fail("dnmf_motion_step_host: NULL argument");
  if (B < 1) return fail("dnmf_motion_step_host: B must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  // Frames cross PCIe in chunks on a copy stream, double-buffered, while the fused kernel works on the
  // previous chunk; the device only ever holds two chunks of the batch.
  const int chunk = std::min(B, std::max(1, (int)(((size_t)96 << 20) / (c->N * sizeof(float)))));
  if (!c->copy_stream) {
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CU(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
  }
  if (ensure(&c->d_batch, &c->batch_cap, (size_t)2 * chunk * c->N)) return 1;
  if (ensure(&c->d_ids, &c->ids_cap, (size_t)B)) return 1;
  if (ensure(&c->d_sse, &c->sse_cap, (size_t)B)) return 1;
  CU(cudaMemcpyAsync(c->d_ids, frame_ids_host, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(c->ev_done[0], st));  // the copy stream must not overtake earlier work on `st`
  CU(cudaEventRecord(c->ev_done[1], st));
  int i = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++i) {
    const int nb = std::min(chunk, B - b0);
    const int slot = i & 1;
    float* buf = c->d_batch + (size_t)slot * chunk * c->N;
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_done[slot], 0));
    CU(cudaMemcpyAsync(buf, frames_host + (size_t)b0 * c->N, (size_t)nb * c->N * sizeof(float),
                       cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    CU(cudaStreamWaitEvent(st, c->ev_copied[slot], 0));
    if (dnmf_loss_grad(c, buf, c->d_ids + b0, nb, B_global, beta_dev, C_dev, c->d_grad, c->d_sse + b0, stream))
      return 1;
    CU(cudaEventRecord(c->ev_done[slot], st));
  }
  if (dnmf_adam_step(c, beta_dev, c->d_grad, m_dev, v_dev, lr, beta1, beta2, eps, step, affine, c->d_sse, B,
                     B_global, c->d_loss, stream))
    return 1;
  if (loss_host) {
    CU(cudaMemcpyAsync(loss_host, c->d_loss, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int dnmf_forward(dnmf_ctx* c, const int32_t* frame_ids_dev, int B, const float* beta_dev,
                            const float* C_dev, float* AtC_dev, float* At_dev, float* grid_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !C_dev) return fail("dnmf_forward: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (AtC_dev) {
    // Yhat does not depend on the video: the WRITE_YHAT instantiation never reads `frames`.
    FitParams p;
    if (fill_fit_params(c, p, AtC_dev, frame_ids_dev, B, beta_dev, C_dev)) return 1;
    p.yhat = AtC_dev;
    if (dispatch_fit<1>(c, p, B, st)) return 1;
    c->counters[0] += 1;
  }
  if (At_dev || grid_dev) {
    if (!c->have_footprints) return fail("dnmf_forward: call dnmf_set_footprints first");
    const size_t total = c->N * (size_t)B;
    dense_forward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        geom_of(c), frame_ids_dev, B, beta_dev, c->d_tab[0], c->d_tab[1], c->d_tab[2], At_dev, grid_dev);
    CU(cudaGetLastError());
    c->counters[5] += 1;
  }
  return 0;
}

extern "C" int dnmf_get_counters(dnmf_ctx* c, int64_t* out) {
  if (!c || !out) return fail("dnmf_get_counters: NULL argument");
  memcpy(out, c->counters, sizeof(c->counters));
  return 0;
}

// ---- trace update + registered video --------------------------------------------------------------
#include "dnmf_mu.inc.cu"
#include "dnmf_ext.inc.cu"
