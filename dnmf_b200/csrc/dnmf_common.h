// Declarations shared by the translation units of libdnmf_b200.so: error plumbing, launch geometry, the
// parameter block of the fused kernel and the per-GPU context behind the C ABI (include/dnmf_b200.h).
#pragma once
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
#ifndef DNMF_AFFINE_BODIES
#define DNMF_AFFINE_BODIES 1  // 1: affine frames with frozen quadratic rows (FitParams::skip_quad) take main loops without
                              // the z^2 Horner term and the z^2 gradient moments (6 packed + 1 scalar op per z step fewer)
#endif
// Checked build (python -m dnmf_b200.build --checked -> _C/libdnmf_b200_checked.so, -DDNMF_CHECKED): device-side
// assertions on every index the kernels form without a clamp -- frame ids, list lengths, the staged-window index of
// the unclamped main loops, table rows of the slice gathers, partial-block slots.  compute-sanitizer is closed on the
// GPU pool this was developed on; tests/test_gpu_checked.py runs a cross-section of the suite through this library.
#ifdef DNMF_CHECKED
#include <cassert>
#define DNMF_DASSERT(cond) assert(cond)
#else
#define DNMF_DASSERT(cond) ((void)0)
#endif
#ifndef DNMF_RESTAGE_BATCH
#define DNMF_RESTAGE_BATCH 4  // slot pairs whose gathers are in flight per thread while the slices are rebuilt
#endif
#ifndef DNMF_SHARE_XZ
#define DNMF_SHARE_XZ 1  // 1: the z-split (dense-list) kernels load the x and z slice entries of a slot pair once for the
                         // voxels A and B of a lane when they coincide (march_rolled SHARE)
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
// error plumbing (text kept per thread, returned by dnmf_last_error)
// ------------------------------------------------------------------------------------------------
int fail(const std::string& s);
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +    \
                  std::to_string(__LINE__) + ")");                                                \
  } while (0)

struct Geom {
  int X, Y, Z, K, T;
  int tx, ty, tz, ntx, nty, ntz;
};

// Per-(frame, tile) partial blocks of the trace statistics, written by a first-stage kernel (fused tiles, SIMT
// panel or tensor-core panel) and summed per frame in ascending tile order by stats_reduce_kernel: tile-frame
// tf = (batch position) * nt + tile holds its list length count[tf], the listed neuron ids[tf][0..L), the
// values vals[tf][j][l] = sum_p A_j A_l (l < L) and vals[tf][j][capL] = sum_p A_j Y; slot_of[b][k][tile] is the
// row of neuron k in that tile's list (0xffff: not listed).
struct StatsPartials {
  float* vals;
  unsigned short* ids;
  int* count;
  unsigned short* slot_of;
  int capL, ld;  // rows available per block, floats per row (capL + 4: rows stay 16-byte aligned)
};

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
  int nwz;  // warp groups splitting the tile's z range (1, 2 or 4)
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
  StatsPartials stats;  // MODE 3 (trace statistics): partial blocks of this launch's tile-frames
  int* mu_overflow;  // MODE 3: set when a tile's list is not fully staged (the caller reruns the generic kernel)
  const int4* windows;  // [B][tiles][2]: (wlo0, wlo1, wlo2, whi0), (whi1, whi2, clipped, 0) from tile_windows_kernel, or NULL
  int skip_quad;  // != 0: gradient rows 4..9 are not wanted (affine fit: Adam freezes them) and are returned as zero
  int chunks_main;  // the first chunks_main chunks of a launch walk fpc frames each, the rest fpc_tail (set by launch_fit)
  unsigned* reserved1;
  int y_pitch;   // floats between x rows of the Y tile in shared memory (>= ty * tile depth)
  int z_skew;    // != 0: lane (lx, ly) starts its z march at ((ly * z_skew) & 3), see march_rolled<SKEW>
  int tmap_ok;   // the frame tile can be fetched with ONE tensor TMA copy (3-D map over [frame][x][y*Z])
  int b_base;    // index of this launch's first frame in the buffer the tensor map describes
  // New fields go HERE, behind everything the kernels have always read: the offsets of the fields above decide which
  // parameters ptxas fetches in pairs, and moving them changed the generated main loops (cfg3 / cfg4 lost 6-7 %
  // when an 8-byte pointer was inserted next to `bg`).
  const float* bg_dev;  // MODE 2: the background is read from the device instead (dnmf_ext_step_begin), or NULL
  int fpc_tail;         // frames per CTA of the last chunks of a launch (short CTAs at the end: the SMs run dry together)
  int cta_slots;        // CTAs resident on the device at once (host-side input of that split)
  alignas(64) CUtensorMap tmap;
};

struct FitSmem {
  int tab_f2;    // float2 count of the staged-table region
  int y_f;       // float count of the Y tile
  int list_u16;  // uint16 count of the list
  size_t bytes;
};

inline FitSmem fit_smem_layout(int nw, int tx, int ty, int tz, int cap, int wsum, int K, int wmax0, int cand_cap,
                               int y_pitch) {
  FitSmem s;
  s.tab_f2 = cap * (wsum + wmax0);  // live slices + the x slice without traces
  s.y_f = tx * std::max(y_pitch, ty * tz) + 4;
  s.list_u16 = (K + 7) & ~7;
  s.bytes = (((size_t)s.tab_f2 * 8 + 127) & ~(size_t)127) + (size_t)s.y_f * 4 + (size_t)nw * kWarpScratch * 4 +
            80 * 4 + 16 +
            (size_t)((cap + 5) & ~3) * 4 + (size_t)cand_cap * 28 + (size_t)((cand_cap + 7) & ~7) * 2 + (size_t)((cap + 7) & ~7) * 2 +
            (size_t)s.list_u16 * 2;
  return s;
}

// One launch sequence of fit_tile_kernel<NWX, NWY, SUB, MODE, FAST_DIV> for B frames (several launches when the
// batch exceeds gridDim.z).  Each MODE lives in its own translation unit (fit_mode<N>.cu) so that they compile
// in parallel; MODE 3 exists for the two-sub-tile layouts with the verified fast division only.
int launch_fit_mode0(int nwx, int nwy, int sub, bool fast_div, const FitParams& p, int B, size_t smem, cudaStream_t st);
int launch_fit_mode1(int nwx, int nwy, int sub, bool fast_div, const FitParams& p, int B, size_t smem, cudaStream_t st);
int launch_fit_mode2(int nwx, int nwy, int sub, bool fast_div, const FitParams& p, int B, size_t smem, cudaStream_t st);
int launch_fit_mode3(int nwx, int nwy, int sub, bool fast_div, const FitParams& p, int B, size_t smem, cudaStream_t st);

// Tensor-core panel Gram (dnmf_gram_tc.cu): one CTA per (frame, 8 x 8 x Z tile), lists of up to 127 neurons.
constexpr int kGramRows = 128, kGramTX = 8, kGramTY = 8;
struct GramTcParams {
  const float* frames;
  const int* frame_ids;
  const float* beta;
  const float2* tab0;
  const float2* tab1;
  const float2* tab2;
  const int* rng;
  StatsPartials out;
  int* overflow;
  int frames_are_batch;
  int X, Y, Z, K, T;
  int ntx, nty;
  int b_base;  // batch position of this launch's first frame in the partial buffers
  int fast_div;  // exact 3-instruction division verified for all three axes (verify_coord_kernel)
  float rcp0, rcp1, rcp2;
  const long long* cand_off;  // the fit kernel's static per-tile candidate lists when its tiles are these 8 x 8 x Z
  const int* cand_ids;        // tiles (NULL otherwise: every CTA scans all K neurons)
  int cand_expand;
};
int launch_gram_tc(const GramTcParams& p, int B, cudaStream_t st);
size_t gram_tc_smem_bytes(int X, int Y, int Z);

}  // namespace dnmf

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
  int nwx = 1, nwy = 1, nwz = 1, tz = 0, cap = 0, user_cap = 0;
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
  long long cand_ids_cap = 0;
  // extension, device-resident shared parameters (dnmf_ext_step_begin / _end): background, Adam state of
  // pos[K][3], sigma[K], b
  float* d_bg = nullptr;
  float* d_ext_m = nullptr;   // [4K+1]
  float* d_ext_v = nullptr;   // [4K+1]
  int cand_expand = 6;
  int cand_cap = 0;
  int affine_grad = 0;   // dnmf_set_affine: dnmf_loss_grad leaves the quadratic gradient rows zero
  int affine_call = 0;   // the same for the duration of one step call made with affine != 0
  int fpc_override = 0;  // DNMF_FPC environment override of the frames-per-CTA heuristic (tuning)
  int fpc_tail_off = 0;  // DNMF_FPC_TAIL_OFF: every chunk fpc frames long (A/B of the short-tail split)
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
  bool video_owned = true;  // false: the slab belongs to the caller (dnmf_attach_frames)
  // scratch
  float* d_partials = nullptr;
  size_t partials_cap = 0;
  int4* d_windows = nullptr;  // tile windows of the current batch (tile_windows_kernel)
  size_t windows_cap = 0;
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
  // frame-id validation (check_ids_kernel): clamped copy of the caller's ids, per-frame stamps, per-call flags and
  // the sticky error word in host-mapped memory
  int* d_ids_safe = nullptr;
  size_t ids_safe_cap = 0;
  int* d_id_mark = nullptr;   // [T]
  int id_stamp = 0;
  int* d_id_flags = nullptr;  // [4]
  int* h_sticky = nullptr;    // cudaHostAllocMapped
  int* d_sticky = nullptr;    // device alias of h_sticky
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
  // partial blocks of the trace statistics (StatsPartials) for one chunk of frames
  float* d_pb_vals = nullptr;
  size_t pb_vals_cap = 0;
  unsigned short* d_pb_ids = nullptr;
  size_t pb_ids_cap = 0;
  int* d_pb_count = nullptr;
  size_t pb_count_cap = 0;
  unsigned short* d_pb_slot = nullptr;
  size_t pb_slot_cap = 0;
  int mu_fused_cap_used = 0;  // staged-slot capacity of the last fused-tile statistics pass
  int mu_prefer_tc = 0;    // dnmf_mu_path bit 3: tensor-core panel kernel instead of the fused tiles
  int mu_last_tc = 0;
  unsigned long long* d_keys = nullptr;
  size_t keys_cap = 0;
  int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace dnmf {

inline Geom geom_of(const dnmf_ctx* c) {
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

}  // namespace dnmf
