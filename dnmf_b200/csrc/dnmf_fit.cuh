// kernel 2 of the dNMF fit hot path: fit_tile_kernel (fused forward / residual / loss / analytic beta-gradient,
// forward-only, background + residual, trace statistics) and its launcher.  Included by fit_mode<N>.cu, one
// translation unit per MODE.
#pragma once
#include "dnmf_common.h"

namespace dnmf {

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

// the same, producing the accumulators (the first slot pair of a list): exactly slot_pair<0, true>'s arithmetic
__device__ __forceinline__ void pair2_first(float2 exG, float2 exD, float2 eyG, float2 eyD, float2 ezG, float2 ezD,
                                            float2 f0, float2 f1, float2 f2, float2& yh, float2& g0, float2& g1,
                                            float2& g2) {
  const float2 ca0 = __ffma2_rn(f0, exD, exG);
  const float2 a1 = __ffma2_rn(f1, eyD, eyG);
  const float2 a2 = __ffma2_rn(f2, ezD, ezG);
  const float2 t12 = __fmul2_rn(a1, a2);
  yh = __fmul2_rn(ca0, t12);
  g0 = __fmul2_rn(exD, t12);
  g1 = __fmul2_rn(__fmul2_rn(ca0, a2), eyD);
  g2 = __fmul2_rn(__fmul2_rn(ca0, a1), ezD);
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

// The same for 16 values: afterwards lanes l and l ^ 16 hold the total of v[l & 15] (16 shuffles).
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
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
// Empty list: Yhat = 0, no gradient; only the loss term.
template <bool SAFE, int MODE>
__device__ __forceinline__ void march_empty(const MarchArgs& a, MarchOut& o) {
  // accumulators start from an opaque zero (read back from shared memory): with a literal 0 ptxas peels the first z step
  // into a straight-line copy of the whole loop body (fold of 0 + x), which costs instruction-cache footprint
  const float oz = a.oz;
  const float2 zero2 = make_float2(oz, oz);
#pragma unroll
  for (int d = 0; d < 3; ++d) o.S0[d] = o.S1[d] = o.S2[d] = zero2;
  float2 sse = zero2, sum_r = zero2;
  unsigned yaddr = a.yaddrA;
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
  o.sse = sse;
  o.sum_r = sum_r;
}

// Main loop of the two-sub-tile layouts: the slot-pair count is a run-time (tile-uniform) value and the pair loop is a
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

// TAIL 4: np >= 1 full slot pairs, and a run-time (warp-uniform) test per z step for one last slot of an odd list.
// TAIL 2: a single slot (np == 0).  Two bodies per (SAFE, SKEW) and not more: with even / odd lists as separate bodies
// (three per SAFE) the per-frame list / restage code of a late fit pushes the hot set past the 32 KB instruction cache
// (stall_no_instruction 1.7 per issue, profiles/README.md); one body with a fully run-time tail kind is slower at
// identity.  Measured at cfg2, ms per 1000 frames, identity beta / a deformation per frame: three bodies 2.65 / 3.40,
// one body 2.77 / 3.05, these two 2.65 / 2.94 (round 1).
// SKEW: the lanes of a warp walk z in rotated order (lane-dependent start, wrap-around), which spreads their
// reads of the Y tile over the banks when the tile's y/x pitches are multiples of 32 floats (Z = 32).
// AFF: the frame's quadratic coefficients (rows 4..9 of beta_t) are all zero and their gradient rows are not wanted
// (FitParams::skip_quad): 2q = c1 z + c0 exactly as Horner with c2 = 0 would give it, and the z^2 moments are dropped.
// SHARE (dense lists, NWZ kernels): when every lane's voxels A and B fall on the same x and z table entries -- B is
// A four rows further in y, so they do unless the deformation's y-coupling carries a lane across a node -- the slot
// pairs' x and z slice entries are loaded once for both voxels: four LDS.128 per slot pair and voxel pair instead of
// six.  Dense lists are bound by the shared-memory pipe (79 % of its wavefront peak at cfg4), not by the FP32 pipe.
// The test is one warp vote per z step; the arithmetic is unchanged (same operands, same order).
template <bool SAFE, int MODE, int TAIL, bool SKEW, bool AFF = false, bool SHARE = false>
__device__ __forceinline__ void march_rolled(const MarchArgs& a, int np, MarchOut& o, int tail = 0) {
  const float oz = a.oz;
  const float2 zero2 = make_float2(oz, oz);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    o.S0[d] = o.S1[d] = zero2;
    if (!AFF) o.S2[d] = zero2;  // AFF: set after the loop (kept live across it they were re-materialised every z step)
  }
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
      const float2 q = AFF ? __ffma2_rn(z2, a.c1[d], a.c0[d])
                           : __ffma2_rn(z2, __ffma2_rn(z2, make_float2(a.c2[d], a.c2[d]), a.c1[d]), a.c0[d]);  // = 2q
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
      // the unclamped loop is chosen only when the conservative window covers every sample of the tile
      DNMF_DASSERT(SAFE || (iA >= a.wl[d] && iA - a.wl[d] <= a.wm1[d] && iB >= a.wl[d] && iB - a.wl[d] <= a.wm1[d]));
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
    if (TAIL != 2) {
      // first slot pair produces the accumulators, the rest of the list updates them
      float2 yA, gA0, gA1, gA2, yB, gB0, gB1, gB2;
      bool shared_xz = false;
      if constexpr (SHARE) shared_xz = __all_sync(0xffffffffu, adA[0] == adB[0] && adA[2] == adB[2]);
      if (SHARE && shared_xz) {
        {
          const float4 ex = lds128r(adA[0]), ez = lds128r(adA[2]), eyA = lds128r(adA[1]), eyB = lds128r(adB[1]);
          pair2_first(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(eyA.x, eyA.y),
                      make_float2(eyA.z, eyA.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fA0, fA1, fA2, yA, gA0,
                      gA1, gA2);
          pair2_first(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(eyB.x, eyB.y),
                      make_float2(eyB.z, eyB.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fB0, fB1, fB2, yB, gB0,
                      gB1, gB2);
        }
#pragma unroll 1
        for (unsigned off = 16u; off < pair_bytes; off += 16u) {
          const float4 ex = lds128r(adA[0] + off), ez = lds128r(adA[2] + off);
          const float4 eyA = lds128r(adA[1] + off), eyB = lds128r(adB[1] + off);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(eyA.x, eyA.y),
                           make_float2(eyA.z, eyA.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fA0, fA1, fA2,
                           yA, gA0, gA1, gA2);
          pair2_accumulate(make_float2(ex.x, ex.y), make_float2(ex.z, ex.w), make_float2(eyB.x, eyB.y),
                           make_float2(eyB.z, eyB.w), make_float2(ez.x, ez.y), make_float2(ez.z, ez.w), fB0, fB1, fB2,
                           yB, gB0, gB1, gB2);
        }
      } else {
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
      }  // !shared_xz
      yh = make_float2(yA.x + yA.y, yB.x + yB.y);
      g[0] = make_float2(gA0.x + gA0.y, gB0.x + gB0.y);
      g[1] = make_float2(gA1.x + gA1.y, gB1.x + gB1.y);
      g[2] = make_float2(gA2.x + gA2.y, gB2.x + gB2.y);
      if (tail == 1) tail_slot<false>(adA, adB, pair_bytes, f, oz, yh, g);
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
    // z-moments of r * dYhat/dix_d accumulated as fma(z^m r, g_d, S): one packed op per moment and axis
    // (the product-first form S + fl(r g) z^m needs an extra multiply per axis)
    const float2 zr = __fmul2_rn(z2, r);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      o.S0[d] = __ffma2_rn(r, g[d], o.S0[d]);
      o.S1[d] = __ffma2_rn(zr, g[d], o.S1[d]);
    }
    if constexpr (!AFF) {
      const float2 zzr = __fmul2_rn(z2, zr);
#pragma unroll
      for (int d = 0; d < 3; ++d) o.S2[d] = __ffma2_rn(zzr, g[d], o.S2[d]);
    }
  }
  o.sse = sse;
  o.sum_r = sum_r;
  if (AFF) {
#pragma unroll
    for (int d = 0; d < 3; ++d) o.S2[d] = make_float2(0.f, 0.f);
  }
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

// Warp-reduce the accumulators of one block pair and store them in the tile-frame's partial block
// (StatsPartials: blk[j * ld + l] = sum A_j A_l, blk[j * ld + capL] = sum A_j Y).  Row slots start at slot 6*pb,
// column slots at 6*lb; off-diagonal blocks are mirrored, b_t comes from the diagonal blocks only.  With several
// warps per CTA the warps' totals are added in warp order through shared memory (sRed: kWarpScratch floats per
// warp): every entry of the block is written exactly once, by plain stores -- the per-frame sum over tiles
// happens in stats_reduce_kernel in a fixed order.
template <int NP, bool TWO, int NW>
__device__ __forceinline__ void flush_stats(const float2 (&G)[3][6], const float2 (&bv)[3], int pb, int lb, int nst,
                                            float* blk, int ld, int capL, float* sRed, int lane, int warp) {
  constexpr int NCS = TWO ? 6 : 2 * NP;        // column slots
  constexpr int NG = 2 * NP * NCS;              // G outputs, index = (p*NCS + c)*2 + half
  constexpr int NOUT = NG + (TWO ? 0 : 2 * NP);  // + b outputs
  static_assert(NOUT <= kWarpScratch, "per-warp reduction area too small");
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
    float tot = warp_transpose_sum(v, lane);
    const int idx = r0 + lane;
    if (NW > 1) {
      if (idx < NOUT) sRed[warp * kWarpScratch + idx] = tot;
      __syncthreads();
      if (warp == 0 && idx < NOUT) {
        tot = sRed[idx];
#pragma unroll
        for (int w = 1; w < NW; ++w) tot += sRed[w * kWarpScratch + idx];
      }
    }
    if (NW == 1 || warp == 0) {
      if (idx < NG) {
        const int pc = idx >> 1, pp = pc / NCS, cc = pc % NCS;
        const int j = 6 * pb + 2 * pp + (idx & 1), l = 6 * lb + cc;
        if (j < nst && l < nst) {
          DNMF_DASSERT(j < capL && l < capL && capL < ld);
          blk[(size_t)j * ld + l] = tot;
          if (TWO) blk[(size_t)l * ld + j] = tot;
        }
      } else if (idx < NOUT) {
        const int q = idx - NG, j = 6 * pb + q;
        if (j < nst) blk[(size_t)j * ld + capL] = tot;
      }
    }
    if (NW > 1) __syncthreads();  // sRed is reused by the next round / block pair
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
// AFFK (MODE 0, launched only with FitParams::skip_quad): the instantiation for affine fits.  Its specialised main
// loops are the AFF ones only (march_rolled); a frame that does carry quadratic coefficients takes the generic loop.
// A separate instantiation rather than more bodies in one kernel: the hot code of either kernel stays as small as
// before (instruction cache; with both families in one kernel cfg3 lost 6 % to code placement alone).
// NWZ > 1 (dense configurations): NWZ groups of NWX*NWY warps share one tile and its staged slices, each group marching
// its own part of the tile's z range.  The tile stays 8*NWX x 4*NWY*SUB voxels wide -- the list length follows the
// tile's x, y extent, (8+s)(8+s) against (16+s)(16+s) for a footprint s nodes wide -- while the CTA still brings enough
// warps per SM for its shared memory (the staged slices of ~85 neurons take 40 KB).
template <int NWX, int NWY, int SUB, int MODE, bool FAST_DIV, bool AFFK = false, int NWZ = 1>
__global__ void __launch_bounds__(32 * NWX * NWY * NWZ, (NWX * NWY * NWZ == 1) ? (MODE == 3 ? DNMF_MU_MINB : DNMF_MINB) : (NWZ > 1 ? 512 / (32 * NWX * NWY * NWZ) : 1)) fit_tile_kernel(const __grid_constant__ FitParams p) {
  static_assert(!AFFK || (MODE == 0 && SUB == 2 && FAST_DIV), "affine instantiation: fit, two sub-tiles, fast division");
  static_assert(NWZ == 1 || MODE != 3, "the fused trace statistics do not split z");
  constexpr bool WRITE_YHAT = MODE == 1;
  constexpr bool WRITE_RES = MODE == 2;
  constexpr int NWXY = NWX * NWY;
  constexpr int NW = NWXY * NWZ;
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
  float* sBeta = sRed + NW * kWarpScratch;          // 32 floats
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
  // the last chunks of a launch are short (launch_fit): the tail of the grid then ends within a fraction of a CTA's time
  const bool main_chunk = chunk < p.chunks_main;
  const int b_first = main_chunk ? chunk * p.fpc : p.chunks_main * p.fpc + (chunk - p.chunks_main) * p.fpc_tail;
  const int nb = main_chunk ? p.fpc : min(p.fpc_tail, p.B - b_first);  // frames this CTA walks (<= 32)
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
  // ... and the tile's sample window under beta_t (tile_windows_kernel: eight ints per tile-frame)
  int win_next = 0;
  const size_t win_row0 = ((size_t)(p.b_base + b_first) * (p.ntx * p.nty * p.ntz) + (size_t)((bz * p.nty + by) * p.ntx + bx)) * 8;
  const size_t win_step = (size_t)(p.ntx * p.nty * p.ntz) * 8;
  auto prefetch_frame = [&](int fi) {
    const int t = __shfl_sync(0xffffffffu, my_frame, fi);
    if (tid < 30) beta_next = p.beta[(size_t)tid * p.T + t];
    if (tid < 8) win_next = reinterpret_cast<const int*>(p.windows)[win_row0 + (size_t)fi * win_step + tid];
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
  const int wxy = NWZ == 1 ? warp : warp % NWXY, wz = NWZ == 1 ? 0 : warp / NWXY;
  const int lane_x = lane & 7, lane_y = lane >> 3;  // lane -> (x, y) of the warp's 8 x 4 footprint
  const int lx = (wxy % NWX) * kWarpX + lane_x;
  const int ly0 = (wxy / NWX) * (kWarpY * SUB) + lane_y;
  // this warp's part of the tile's z range
  const int zb = NWZ == 1 ? 0 : (nz * wz) / NWZ;
  const int nzw = NWZ == 1 ? nz : (nz * (wz + 1)) / NWZ - zb;
  const int gx = x0 + lx;
  const float xf = (float)gx;
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
  const unsigned strideB = (unsigned)CAP * 8u;
  const unsigned baseX = smem_u32(sTab);
  const unsigned baseY = baseX + (unsigned)p.wmax0 * strideB;
  const unsigned baseZ = baseY + (unsigned)p.wmax1 * strideB;
  const float bg = WRITE_RES ? (p.bg_dev != nullptr ? *p.bg_dev : p.bg) : 0.f;

  // state carried from frame to frame: what the staged slices were built for
  int pw_lo[3] = {0x7fffffff, 0, 0}, pw_hi[3] = {0, 0, 0}, prev_L = -1;
  bool prev_fast = false;  // previous list came from the cached candidates with prefetched traces

  for (int fi = 0; fi < nb; ++fi) {
    const int b = b_first + fi;
    const int t = __shfl_sync(0xffffffffu, my_frame, fi);
    DNMF_DASSERT(t >= 0 && t < p.T);
    if (tid < 30) sBeta[tid] = beta_next;
    if (tid < 8) sInt[tid] = win_next;
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

    // ---- conservative window of this tile under beta_t: fetched one frame ahead, broadcast through shared memory ----
    int wlo[3], whi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      wlo[d] = sInt[d];
      whi[d] = sInt[3 + d];
    }
    const bool window_clipped = sInt[6] != 0;
    // affine frame (rows 4..9 of beta_t all zero) whose quadratic gradient rows the caller does not want
    bool quad_zero = false;
    if constexpr (AFFK) quad_zero = __all_sync(0xffffffffu, lane >= 18 || sBeta[12 + lane] == 0.f);

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
    DNMF_DASSERT(L >= 0 && L <= p.K && W0 >= 1 && W1 >= 1 && W2 >= 1);
    DNMF_DASSERT(wlo[0] >= -2 && whi[0] <= p.X && wlo[1] >= -2 && whi[1] <= p.Y && wlo[2] >= -2 && whi[2] <= p.Z);  // table rows
    const int npair = (nst + 1) >> 1;
    if ((nst & 1) && tid == 0) sCk[nst] = 0.f;  // partner of the last neuron of an odd list: zero footprint
    if (changed) {
      // Slot pairs (2p, 2p+1) share one float4 (G_2p, G_2p+1, D_2p, D_2p+1) = the packed operands of FFMA2; an odd
      // list is completed with a zero footprint.  The x slice is kept without the traces (they change with the
      // frame, the slices usually do not).  One thread owns table entry e of every slot; loads are issued four
      // pairs at a time ahead of the stores.  (Dealing (pair, entry) items to all lanes instead was measured 7 %
      // slower with a deformation per frame: more address arithmetic per gather than it saves in idle lanes.)
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
        constexpr int kBatch = DNMF_RESTAGE_BATCH;
        for (int p0 = 0; p0 < npair; p0 += kBatch) {
          float2 va[kBatch], vb[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int j = 2 * (p0 + u);
            va[u] = vb[u] = make_float2(0.f, 0.f);
            DNMF_DASSERT(j >= nst || sList[j] < p.K);
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
      a.yaddrA = smem_u32(sY + lx * RS + ly0 * zs + zb);
      a.zf0 = (float)(z0 + zb);
      a.nz = nzw;
      a.validA = (gx < p.X) && (y0 + ly0 < p.Y);
      a.validB = (gx < p.X) && (y0 + ly0 + kWarpY < p.Y);
      a.bg = bg;
      a.zskew = 0;
      if (p.z_skew != 0 && nzw >= 4) a.zskew = (lane_y * p.z_skew) & 3;
    };

    if constexpr (MODE == 3) {
      // ---- trace statistics of this tile-frame ----
      if constexpr (SUB == 2 && FAST_DIV) {
        const int tile = (bz * p.nty + by) * p.ntx + bx;
        const int ntl = p.ntx * p.nty * p.ntz;
        const size_t tf = (size_t)(p.b_base + b) * ntl + tile;
        if (L > nst || L > p.stats.capL) {
          if (tid == 0) atomicMax(p.mu_overflow, L);  // not fully staged: the caller reruns a panel kernel
        } else {
          DNMF_DASSERT(b >= 0 && b < p.B && L <= p.stats.capL);
          if (tid == 0) p.stats.count[tf] = L;
          for (int pos = tid; pos < L; pos += NT) {
            const int k = sList[pos];
            p.stats.ids[tf * p.stats.capL + pos] = (unsigned short)k;
            p.stats.slot_of[((size_t)(p.b_base + b) * p.K + k) * ntl + tile] = (unsigned short)pos;
          }
        }
        if (L <= nst && L <= p.stats.capL && npair > 0) {
          MarchArgs a;
          fill_march_args(a);
          float* blk = p.stats.vals + tf * (size_t)p.stats.capL * p.stats.ld;
          const int ld = p.stats.ld, capL = p.stats.capL;
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
                flush_stats<3, true, NW>(G, bv, pb, lb, nst, blk, ld, capL, sRed, lane, warp);
              } else if (rowp == 1) {
                march_stats<1, false>(a, rowoff, rowoff, 1, 1, G, bv);
                flush_stats<1, false, NW>(G, bv, pb, lb, nst, blk, ld, capL, sRed, lane, warp);
              } else if (rowp == 2) {
                march_stats<2, false>(a, rowoff, rowoff, 2, 2, G, bv);
                flush_stats<2, false, NW>(G, bv, pb, lb, nst, blk, ld, capL, sRed, lane, warp);
              } else {
                march_stats<3, false>(a, rowoff, rowoff, 3, 3, G, bv);
                flush_stats<3, false, NW>(G, bv, pb, lb, nst, blk, ld, capL, sRed, lane, warp);
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
      if (!has_overflow && (!AFFK || quad_zero)) {
        MarchArgs a;
        fill_march_args(a);
        MarchOut o;
        const bool safe = window_clipped || nx < TX || ny < TY;
        if (npair == 0) {
          if (safe)
            march_empty<true, MODE>(a, o);
          else
            march_empty<false, MODE>(a, o);
        } else {
          const int npf = nst >> 1;  // full slot pairs; an odd list ends with a single slot
          const int tail = (nst & 1) ? (npf == 0 ? 2 : 1) : 0;
          constexpr bool kShare = DNMF_SHARE_XZ == 2 || (DNMF_SHARE_XZ && NWZ > 1);  // dense lists: x / z slice entries shared by voxels A and B
          switch ((p.z_skew != 0 ? 4 : 0) + (tail == 2 ? 2 : 0) + (safe ? 1 : 0)) {
            case 0: march_rolled<false, MODE, 4, false, AFFK, kShare>(a, npf, o, tail); break;
            case 1: march_rolled<true, MODE, 4, false, AFFK, kShare>(a, npf, o, tail); break;
            case 2: march_rolled<false, MODE, 2, false, AFFK>(a, npf, o); break;
            case 3: march_rolled<true, MODE, 2, false, AFFK>(a, npf, o); break;
            case 4: march_rolled<false, MODE, 4, true, AFFK, kShare>(a, npf, o, tail); break;
            case 5: march_rolled<true, MODE, 4, true, AFFK, kShare>(a, npf, o, tail); break;
            case 6: march_rolled<false, MODE, 2, true, AFFK>(a, npf, o); break;
            default: march_rolled<true, MODE, 2, true, AFFK>(a, npf, o); break;
          }
        }
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
      a.sY = sY + zb;
      a.t = t, a.L = L, a.nst = nst;
      a.x0 = x0, a.y0 = y0, a.z0 = z0 + zb, a.nz = nzw;
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
    if constexpr (AFFK && NW == 1) {
      // affine fit: rows 0..3 only (12 moments) + SSE -> a 16-value reduction; rows 4..9 are stored as zero
      float v[16];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float t0 = 0.f, t1 = 0.f, y0s = 0.f;
#pragma unroll
        for (int h = 0; h < SUB; ++h) {
          const float yh_ = (float)(y0 + ly0 + h * kWarpY);
          t0 += S0[h][d];
          t1 += S1[h][d];
          y0s = fmaf(yh_, S0[h][d], y0s);
        }
        v[0 * 3 + d] = t0;
        v[1 * 3 + d] = xf * t0;
        v[2 * 3 + d] = y0s;
        v[3 * 3 + d] = t1;
      }
      v[12] = sse;
      v[13] = v[14] = v[15] = 0.f;
      const float tot = warp_transpose_sum16(v, lane);
      const float sse_tot = __shfl_sync(0xffffffffu, tot, 12);
      p.partials[(size_t)cta_linear * kNumPartials + lane] = lane < 12 ? tot : (lane == 30 ? sse_tot : 0.f);
    } else {
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
      if (MODE == 0 && p.skip_quad) {  // frozen quadratic rows: returned as zero whichever main loop ran
#pragma unroll
        for (int i = 12; i < 30; ++i) v[i] = 0.f;
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
}

template <int NWX, int NWY, int SUB, int MD_, bool FD_, bool AFFK_ = false, int NWZ_ = 1>
static int launch_fit(const FitParams& p0, int B, size_t smem, cudaStream_t st) {
  auto kern = fit_tile_kernel<NWX, NWY, SUB, MD_, FD_, AFFK_, NWZ_>;
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
  const int maxB = std::max(1, 65535 / p0.ntz / 2) * fpc;  // half: the short chunks at the end of a launch add to the count
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
    // chunk sizes: fpc frames per CTA, except for the frames of (about) the last wave of CTAs, which go out in chunks
    // a quarter as long -- the grid is dispatched in chunk order, so the short CTAs run last and the SMs run dry
    // within a quarter of a long CTA's time of each other instead of a whole one.
    int fpc_tail = std::max(1, fpc / 4);
    int tail_frames = 0;
    if (p0.cta_slots > 0 && fpc_tail < fpc) {
      const long long tiles = (long long)nt;
      const long long wave_chunks = (p0.cta_slots + tiles - 1) / tiles;  // chunks resident at once
      tail_frames = (int)std::min<long long>(nb, wave_chunks * fpc);
    }
    const int n_main = (nb - tail_frames) / fpc;
    const int rest = nb - n_main * fpc;
    if (rest == 0) fpc_tail = fpc;
    p.chunks_main = n_main;
    p.fpc_tail = fpc_tail;
    const int chunks = n_main + (rest + fpc_tail - 1) / fpc_tail;
    dim3 grid((unsigned)p0.ntx, (unsigned)p0.nty, (unsigned)(chunks * p0.ntz));
    kern<<<grid, 32 * NWX * NWY * NWZ_, smem, st>>>(p);
    CU(cudaGetLastError());
  }
  return 0;
}

// layout dispatch shared by the per-MODE translation units
#define DNMF_FIT_DISPATCH(MD_)                                                                                 \
  do {                                                                                                         \
    if (p.nwz > 1) {  /* z-split CTAs: the 8 x 8 tile of the one-warp layout, two sub-tiles, fast division */    \
      if (!(nwx == 1 && nwy == 1 && sub == 2 && fd)) return fail("dispatch_fit: warps_z > 1 needs the 1x1 layout with two sub-tiles"); \
      if (p.nwz == 2) return launch_fit<1, 1, 2, MD_, true, false, 2>(p, B, smem, st);                          \
      if (p.nwz == 4) return launch_fit<1, 1, 2, MD_, true, false, 4>(p, B, smem, st);                          \
      return fail("dispatch_fit: warps_z must be 1, 2 or 4");                                                  \
    }                                                                                                          \
    if (nwx == 1 && nwy == 1 && sub == 1) return fd ? launch_fit<1, 1, 1, MD_, true>(p, B, smem, st) : launch_fit<1, 1, 1, MD_, false>(p, B, smem, st); \
    if (nwx == 1 && nwy == 1 && sub == 2) return fd ? launch_fit<1, 1, 2, MD_, true>(p, B, smem, st) : launch_fit<1, 1, 2, MD_, false>(p, B, smem, st); \
    if (nwx == 2 && nwy == 1 && sub == 1) return fd ? launch_fit<2, 1, 1, MD_, true>(p, B, smem, st) : launch_fit<2, 1, 1, MD_, false>(p, B, smem, st); \
    if (nwx == 2 && nwy == 1 && sub == 2) return fd ? launch_fit<2, 1, 2, MD_, true>(p, B, smem, st) : launch_fit<2, 1, 2, MD_, false>(p, B, smem, st); \
    if (nwx == 2 && nwy == 2 && sub == 1) return fd ? launch_fit<2, 2, 1, MD_, true>(p, B, smem, st) : launch_fit<2, 2, 1, MD_, false>(p, B, smem, st); \
    if (nwx == 2 && nwy == 2 && sub == 2) return fd ? launch_fit<2, 2, 2, MD_, true>(p, B, smem, st) : launch_fit<2, 2, 2, MD_, false>(p, B, smem, st); \
    if (nwx == 2 && nwy == 4 && sub == 1) return fd ? launch_fit<2, 4, 1, MD_, true>(p, B, smem, st) : launch_fit<2, 4, 1, MD_, false>(p, B, smem, st); \
    return fail("dispatch_fit: unsupported warp layout");                                                      \
  } while (0)

}  // namespace dnmf
