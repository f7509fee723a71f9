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
#include "dnmf_common.h"

namespace dnmf {

static thread_local std::string g_err;
int fail(const std::string& s) {
  g_err = s;
  return 1;
}

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

// Kernel 1b, window part, for the fused kernel: one THREAD per (batch position, tile) evaluates the tile's
// conservative sample window under beta_t (tile_window_axis: the device code of the binning kernel, bit-exact twin
// of oracle.tile_window) and stores (wlo[3], whi[3], clipped, 0).  The fused kernel used to run this in its
// per-frame prologue on three lanes of a warp (~130 warp instructions per tile-frame for three lanes of work);
// here the lanes are all busy and the fused kernel fetches eight ints one frame ahead.
__global__ void tile_windows_kernel(Geom g, const float* __restrict__ beta, const int* __restrict__ frame_ids,
                                    long long items, int4* __restrict__ out) {
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items) return;
  const int nt = g.ntx * g.nty * g.ntz;
  const int b = (int)(item / nt), tile = (int)(item - (long long)b * nt);
  const int t = frame_ids[b];
  int x0, y0, z0, x1, y1, z1;
  tile_box(g, tile, x0, y0, z0, x1, y1, z1);
  int wlo[3], whi[3];
  bool clipped[3];
  const int sz[3] = {g.X, g.Y, g.Z};
#pragma unroll
  for (int d = 0; d < 3; ++d)
    tile_window_axis(beta + (size_t)d * g.T + t, 3 * g.T, (float)x0, (float)y0, (float)z0, (float)x1,
                     (float)y1, (float)z1, sz[d], wlo[d], whi[d], clipped[d]);
  out[2 * item] = make_int4(wlo[0], wlo[1], wlo[2], whi[0]);
  out[2 * item + 1] = make_int4(whi[1], whi[2], (clipped[0] || clipped[1] || clipped[2]) ? 1 : 0, 0);
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

// Frame ids handed to the C ABI are checked on the device before any kernel indexes with them: one block walks
// the batch, writes a copy clamped to [0, T) (what every kernel of the call then reads: an id out of range can
// no longer touch memory outside the slab), raises the sticky error bit of the context (host-mapped memory, read
// without a synchronisation at the next entry point) for an id outside [0, T), and tells the second-stage
// reduction whether an id occurs twice in the batch (stamp trick: mark[t] holds the number of the last call that
// listed frame t).  flags[0]: bit 0 = duplicates, bit 1 = out of range; rewritten by every call.
__global__ void check_ids_kernel(const int* __restrict__ ids, int B, int T, int* __restrict__ mark, int stamp,
                                 int* __restrict__ ids_safe, int* __restrict__ flags, int* __restrict__ sticky) {
  __shared__ int s_flags;
  if (threadIdx.x == 0) s_flags = 0;
  __syncthreads();
  int mine = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int t = ids[b];
    if (t < 0 || t >= T) {
      mine |= 2;
      t = min(max(t, 0), T - 1);
    } else if (atomicExch(mark + t, stamp) == stamp) {
      mine |= 1;
    }
    ids_safe[b] = t;
  }
  if (mine) atomicOr(&s_flags, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    flags[0] = s_flags;
    if (s_flags & 2) atomicOr(sticky, 2);
  }
}

// Second stage: per frame, sum the CTA partials in a fixed order (double), scale by 2/(B_global*N),
// write the frame's gradient column and its sum of squared residuals.  grid = B, block = 256.
// A frame id that occurs more than once in the batch (a sampler with replacement) owns ONE gradient column: the
// reference's autograd adds the contributions of all its occurrences (index_put with accumulate,
// Demix/dNMF.py:54,190).  Here the block of the first occurrence sums them in ascending batch position and the
// other blocks leave; every position still reports its own squared residuals (F.mse_loss averages over all B
// entries, :188).  flags[0] bit 0 (check_ids_kernel) says whether the batch has duplicates at all.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, const int* __restrict__ frame_ids, int B,
                                       int nt, int T, double grad_scale, float* __restrict__ grad,
                                       double* __restrict__ sse_out, double* __restrict__ sumr_out,
                                       const double* __restrict__ scale_per_frame, const int* __restrict__ flags) {
  __shared__ double s[8][kNumPartials];
  __shared__ int s_rel[2];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = frame_ids[b];
  DNMF_DASSERT(t >= 0 && t < T);
  int last = b;  // positions b..last may hold further occurrences of t
  if (flags != nullptr && (flags[0] & 1)) {
    if (threadIdx.x < 2) s_rel[threadIdx.x] = 0;
    __syncthreads();
    bool earlier = false, later = false;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const bool same = frame_ids[i] == t;
      earlier |= same && i < b;
      later |= same && i > b;
    }
    if (earlier) s_rel[0] = 1;
    if (later) s_rel[1] = 1;
    __syncthreads();
    if (s_rel[0]) return;          // not the first occurrence: the owner block accounts for this position
    if (s_rel[1]) last = B - 1;
  }
  double gsum = 0.0;
  for (int i = b; i <= last; ++i) {
    if (i != b && frame_ids[i] != t) continue;  // block-uniform
    const float* src = partials + (size_t)i * nt * kNumPartials;
    double acc = 0.0;
    for (int j = warp; j < nt; j += 8) acc += (double)src[(size_t)j * kNumPartials + lane];
    __syncthreads();  // the previous position's s[][] has been read
    s[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s[w][lane];
      const double sc = scale_per_frame != nullptr ? scale_per_frame[i] : grad_scale;
      if (lane < 30) gsum = i == b ? v * sc : gsum + v * sc;
      if (lane == 30) sse_out[i] = v;
      if (lane == 31 && sumr_out != nullptr) sumr_out[i] = v;
    }
  }
  if (warp == 0 && lane < 30) grad[(size_t)lane * T + t] = (float)gsum;
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

extern "C" int dnmf_abi_version(void) { return DNMF_ABI_VERSION; }

namespace dnmf {
__global__ void trip_assert_kernel(int never) { DNMF_DASSERT(never == 12345); }
}  // namespace dnmf

extern "C" int dnmf_build_info(void) {
#ifdef DNMF_CHECKED
  return 1;
#else
  return 0;
#endif
}

extern "C" int dnmf_debug_trip_assert(int device) {
  CU(cudaSetDevice(device));
  dnmf::trip_assert_kernel<<<1, 32>>>(0);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return 0;
}
extern "C" const char* dnmf_last_error(void) { return g_err.c_str(); }

static int create_impl(dnmf_ctx* c, int X, int Y, int Z, int K, int T, int device);

extern "C" void dnmf_destroy(dnmf_ctx* c);

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
  c->device = device;
  if (create_impl(c, X, Y, Z, K, T, device)) {  // the error text is already set; do not leak what was allocated
    dnmf_destroy(c);
    return 1;
  }
  *out = c;
  return 0;
}

static int create_impl(dnmf_ctx* c, int X, int Y, int Z, int K, int T, int device) {
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
  if (const char* ev = getenv("DNMF_FPC_TAIL_OFF")) c->fpc_tail_off = atoi(ev) != 0;
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
  CU(cudaMalloc((void**)&c->d_id_mark, (size_t)T * sizeof(int)));
  CU(cudaMemset(c->d_id_mark, 0, (size_t)T * sizeof(int)));
  CU(cudaMalloc((void**)&c->d_id_flags, 4 * sizeof(int)));
  CU(cudaMemset(c->d_id_flags, 0, 4 * sizeof(int)));
  CU(cudaHostAlloc((void**)&c->h_sticky, sizeof(int), cudaHostAllocMapped));
  *c->h_sticky = 0;
  CU(cudaHostGetDevicePointer((void**)&c->d_sticky, c->h_sticky, 0));
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
  return 0;
}

extern "C" void dnmf_destroy(dnmf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  void* ptrs[] = {c->d_pos,     c->d_sigma, c->d_rng,        c->d_tab[0],      c->d_tab[1],
                  c->d_tab[2],  c->video_owned ? c->d_video : nullptr, c->d_partials,   c->d_grad,        c->d_sse,
                  c->d_windows, c->d_bg, c->d_ext_m, c->d_ext_v,
                  c->d_batch,   c->d_ids,   c->d_loss,       c->d_tmp_counts,  c->d_tmp_offsets,
                  c->d_tmp_max, c->d_G,     c->d_b,          c->d_identity_beta,
                  c->d_Cd[0],   c->d_Cd[1], c->d_keys,       c->d_cand_off,    c->d_cand_ids,
                  c->d_tab_dpos[0], c->d_tab_dpos[1], c->d_tab_dpos[2], c->d_tab_dsig[0], c->d_tab_dsig[1],
                  c->d_tab_dsig[2], c->d_resid, c->d_sumr, c->d_ids_zero,
                  c->d_epoch_batch_of, c->d_epoch_offsets, c->d_epoch_scalars, c->d_epoch_scale,
                  c->d_mu_nbr, c->d_Gc, c->d_ids_safe, c->d_id_mark, c->d_id_flags,
                  c->d_pb_vals, c->d_pb_ids, c->d_pb_count, c->d_pb_slot};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(c->ev_copied[i]);
      cudaEventDestroy(c->ev_done[i]);
    }
  }
  if (c->h_sticky) cudaFreeHost(c->h_sticky);
  delete c;
}

// ---- tiling -----------------------------------------------------------------------------------
static int configure_tiling(dnmf_ctx* c, cudaStream_t st);

extern "C" int dnmf_set_tiling(dnmf_ctx* c, int warps_x, int warps_y, int tz, int slot_capacity, int subtiles_y,
                               int warps_z) {
  if (!c) return fail("dnmf_set_tiling: ctx is NULL");
  if (subtiles_y < 1) subtiles_y = 1;
  if (warps_z < 1) warps_z = 1;
  if (warps_z != 1 && !((warps_z == 2 || warps_z == 4) && warps_x == 1 && warps_y == 1 && subtiles_y == 2))
    return fail("dnmf_set_tiling: warps_z = 2 or 4 is supported for the 1x1 warp layout with subtiles_y = 2");
  if (subtiles_y > 2 || (subtiles_y == 2 && warps_y > 2))
    return fail("dnmf_set_tiling: subtiles_y = 2 is supported for the 1x1, 2x1 and 2x2 warp layouts");
  const bool ok = (warps_x == 1 && warps_y == 1) || (warps_x == 2 && warps_y == 1) ||
                  (warps_x == 2 && warps_y == 2) || (warps_x == 2 && warps_y == 4);
  if (!ok) return fail("dnmf_set_tiling: supported warp layouts are 1x1, 2x1, 2x2, 2x4");
  if (tz < 0) return fail("dnmf_set_tiling: tz must be >= 0");
  c->nwx = warps_x;
  c->nwy = warps_y;
  c->nwz = warps_z;
  c->tz = tz;
  c->user_cap = slot_capacity;
  c->sub = subtiles_y;
  c->auto_tiling = false;
  CU(cudaSetDevice(c->device));
  if (c->have_footprints) return configure_tiling(c, 0);
  return 0;
}

extern "C" int dnmf_set_affine(dnmf_ctx* c, int affine) {
  if (!c) return fail("dnmf_set_affine: ctx is NULL");
  c->affine_grad = affine ? 1 : 0;
  return 0;
}

extern "C" int dnmf_get_tiling(dnmf_ctx* c, int32_t* out) {
  if (!c || !out) return fail("dnmf_get_tiling: NULL argument");
  int32_t v[12] = {c->tx, c->ty, c->tz, c->ntx, c->nty, c->ntz, c->nwx, c->nwy, c->cap, c->sub, c->fast_div, c->nwz};
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
  // {warps_x, warps_y, sub-tiles, warps_z}; the z-split layouts keep the 8 x 8 tile of the one-warp layout (shortest
  // lists) and put 2 or 4 warps on it (long lists: shared memory would leave a one-warp CTA too few warps per SM)
  constexpr int kLayouts = 9;
  static const int layouts[kLayouts][4] = {{1, 1, 2, 1}, {1, 1, 1, 1}, {2, 1, 2, 1}, {2, 2, 2, 1}, {2, 1, 1, 1},
                                          {2, 2, 1, 1}, {2, 4, 1, 1}, {1, 1, 2, 2}, {1, 1, 2, 4}};
  int best = 0;
  double best_cost = 1e300;
  for (int i = 0; i < kLayouts; ++i) {
    c->nwx = layouts[i][0];
    c->nwy = layouts[i][1];
    c->sub = layouts[i][2];
    c->nwz = layouts[i][3];
    if (c->nwz > 1 && c->Z / c->nwz < 4) continue;  // a few z planes per warp: the per-frame setup is not amortised
    if (configure_tiling_fixed(c, st)) return 1;
    const int nw = c->nwx * c->nwy * c->nwz;
    const int ctas = std::min(32, (int)((size_t)227 * 1024 / (c->fit_smem + 1024)));
    const double warps = std::min(64, ctas * nw);
    const double rows = (double)c->sub * c->tz / c->nwz;
    // two sub-tiles per warp run the packed (A, B) march: ~42 fixed + ~7.3 per listed neuron per row
    const double fixed = c->sub == 2 ? 42.0 : 85.0, per = c->sub == 2 ? 7.3 : 9.5;
    // fewer than ~8 resident warps per SM cannot keep the FP32 pipe fed (cfg4, measured: 16 warps 1.0, 8 warps
    // 1.04, 6 warps 1.6, 4 warps 2.0 relative cost)
    // lanes past the volume edge still cost: padded volume / volume
    const double edge = ((double)c->ntx * c->tx / c->X) * ((double)c->nty * c->ty / c->Y);
    // fewer than ~16 resident warps leave the FP32 pipe idle part of the time (cfg4, 8 x 8 tile, measured: 4 CTAs of 4
    // warps 2.80 ms, 4 CTAs of 2 warps 3.31 ms per 100 frames)
    double cost = (fixed + per * c->mean_list_identity + (300.0 + 600.0 / nw) / rows) *
                  std::pow(std::max(1.0, 8.0 / warps), 1.5) * (1.0 + 0.25 * std::max(0.0, 16.0 - warps) / 8.0) * edge;
    if (nw > 1 && c->mean_list_identity < 16.0) cost *= 1.25;  // sharing the staged slices only pays for long lists
    if (cost < best_cost) {
      best_cost = cost;
      best = i;
    }
  }
  c->nwx = layouts[best][0];
  c->nwy = layouts[best][1];
  c->sub = layouts[best][2];
  c->nwz = layouts[best][3];
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
    c->cand_ids_cap = total + total / 2 + 1024;  // slack: the extension refreshes the lists in place as positions move
    CU(cudaMalloc((void**)&c->d_cand_ids, (size_t)c->cand_ids_cap * sizeof(int)));
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
  const int nw = c->nwx * c->nwy * c->nwz;
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
  if (c->d_video && !c->video_owned) {  // a slab attached by the caller is left alone: the context gets its own
    c->d_video = nullptr;
    c->video_owned = true;
    c->tmap_valid = false;
    c->tmap_ptr = nullptr;
  }
  if (!c->d_video) CU(cudaMalloc((void**)&c->d_video, c->N * (size_t)c->T * sizeof(float)));
  float* dst = c->d_video + (size_t)t0 * c->N;
  CU(cudaMemcpyAsync(dst, frames_host, c->N * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
  if (clamp_negative && n > 0) {
    clamp_negative_kernel<<<c->num_sms * 8, 256, 0, st>>>(dst, c->N * (size_t)n);
    CU(cudaGetLastError());
  }
  return 0;
}

extern "C" int dnmf_attach_frames(dnmf_ctx* c, float* frames_dev, int clamp_negative, void* stream) {
  if (!c) return fail("dnmf_attach_frames: ctx is NULL");
  CU(cudaSetDevice(c->device));
  if (c->d_video && c->video_owned) CU(cudaFree(c->d_video));
  c->d_video = frames_dev;             // NULL detaches
  c->video_owned = frames_dev == nullptr;
  c->tmap_valid = false;
  c->tmap_ptr = nullptr;
  if (frames_dev && clamp_negative) {
    if (((uintptr_t)frames_dev & 3) != 0) return fail("dnmf_attach_frames: misaligned pointer");
    clamp_negative_kernel<<<c->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(frames_dev, c->N * (size_t)c->T);
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

// ---- frame-id validation ------------------------------------------------------------------------------
// Errors found on the device by an asynchronous call (a frame id outside [0, T)) are kept in a sticky word in
// host-mapped memory and reported by the next entry point, or by dnmf_check_status.
static int check_sticky(dnmf_ctx* c, const char* who) {
  if (!c->h_sticky) return 0;
  const int bits = *(volatile int*)c->h_sticky;
  if (bits == 0) return 0;
  *(volatile int*)c->h_sticky = 0;
  std::string what;
  if (bits & 2) what += " a frame id outside [0, T) was passed (the call ran on ids clamped into the slab; its results are invalid)";
  if (bits & 8) what += " the candidate lists rebuilt on the device outgrew their buffer (extension: positions moved far; call dnmf_set_footprints)";
  if (bits & ~(2 | 8)) what += " device status " + std::to_string(bits);
  return fail(std::string(who) + ": an earlier asynchronous call failed:" + what);
}

// ids_dev[B] -> the context's clamped copy (what the kernels of this call read) + the duplicate / range flags
static int sanitize_ids(dnmf_ctx* c, const int32_t* ids_dev, int B, cudaStream_t st, const int32_t** out) {
  if (B < 1) return fail("frame id batch is empty");
  if (ensure(&c->d_ids_safe, &c->ids_safe_cap, (size_t)B)) return 1;
  c->id_stamp = c->id_stamp == 0x7fffffff ? 1 : c->id_stamp + 1;
  check_ids_kernel<<<1, 256, 0, st>>>(ids_dev, B, c->T, c->d_id_mark, c->id_stamp, c->d_ids_safe, c->d_id_flags,
                                      c->d_sticky);
  CU(cudaGetLastError());
  c->counters[2] += 1;  // pre-pass launches (id check, tile windows, stand-alone binning)
  *out = c->d_ids_safe;
  return 0;
}

extern "C" int dnmf_check_status(dnmf_ctx* c, void* stream) {
  if (!c) return fail("dnmf_check_status: ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return check_sticky(c, "dnmf_check_status");
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
  if (check_sticky(c, "dnmf_bin_tiles") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
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

template <int MD_>
static int dispatch_fit(dnmf_ctx* c, const FitParams& p, int B, cudaStream_t st) {
  const size_t smem = c->fit_smem;
  if (c->nty > 65535) return fail("dispatch_fit: more than 65535 tiles along y");
  const bool fd = c->fast_div != 0;
  if (MD_ == 0) return launch_fit_mode0(c->nwx, c->nwy, c->sub, fd, p, B, smem, st);
  if (MD_ == 1) return launch_fit_mode1(c->nwx, c->nwy, c->sub, fd, p, B, smem, st);
  return launch_fit_mode2(c->nwx, c->nwy, c->sub, fd, p, B, smem, st);
}

// MODE 3 (trace statistics) exists for the two-sub-tile layouts with the verified fast division only
static bool fused_stats_available(const dnmf_ctx* c) {
  return c->sub == 2 && c->fast_div != 0 && c->nwz == 1 && ((c->nwx == 1 && c->nwy == 1) || (c->nwx == 2 && c->nwy <= 2));
}
static int dispatch_stats(dnmf_ctx* c, const FitParams& p, int B, size_t smem, cudaStream_t st) {
  if (c->nty > 65535) return fail("dispatch_stats: more than 65535 tiles along y");
  if (!fused_stats_available(c)) return fail("dispatch_stats: layout without a fused statistics kernel");
  return launch_fit_mode3(c->nwx, c->nwy, c->sub, true, p, B, smem, st);
}

static int fill_fit_params(dnmf_ctx* c, FitParams& p, const float* frames_dev, const int32_t* ids, int B,
                           const float* beta, const float* C, cudaStream_t st) {
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
  // z-split warp groups exist for the verified fast division only (a singleton axis has none): same tile, one warp
  p.nwz = (c->fast_div != 0 && c->nwx == 1 && c->nwy == 1 && c->sub == 2) ? c->nwz : 1;
  p.cap = c->cap;
  p.wmax0 = c->wmax[0];
  p.wmax1 = c->wmax[1];
  p.wmax2 = c->wmax[2];
  p.full_depth = (c->tz == c->Z) ? 1 : 0;
  p.bulk_ok = (((uintptr_t)p.frames & 15) == 0) && (((size_t)c->Y * c->Z) % 4 == 0) &&
              (((size_t)c->ty * c->Z) % 4 == 0);
  memset(&p.stats, 0, sizeof(p.stats));
  p.mu_overflow = nullptr;
  p.skip_quad = (c->affine_grad || c->affine_call) ? 1 : 0;
  p.chunks_main = 0;  // launch_fit sets the chunk split
  p.fpc_tail = 1;
  p.cta_slots = 0;
  p.reserved1 = nullptr;
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
    p.cta_slots = c->fpc_tail_off ? 0 : c->num_sms * std::max(1, std::min(c->nwx * c->nwy * c->nwz == 1 ? 16 : 32, (int)((size_t)227 * 1024 / (c->fit_smem + 1024))));
  }
  p.yhat = nullptr;
  p.bg = 0.f;
  p.bg_dev = nullptr;
  const size_t need = (size_t)B * c->ntx * c->nty * c->ntz * kNumPartials;
  if (ensure(&c->d_partials, &c->partials_cap, need)) return 1;
  p.partials = c->d_partials;
  {  // the tiles' sample windows of this batch, computed by all lanes ahead of the fused launch (same stream)
    const long long items = (long long)B * c->ntx * c->nty * c->ntz;
    if (ensure(&c->d_windows, &c->windows_cap, (size_t)items * 2)) return 1;
    tile_windows_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(geom_of(c), beta, ids, items, c->d_windows);
    CU(cudaGetLastError());
    p.windows = c->d_windows;
    c->counters[2] += 1;  // pre-pass launches
  }
  return 0;
}

static int launch_fused_fit(dnmf_ctx* c, FitParams& p, int B, cudaStream_t st) { return dispatch_fit<0>(c, p, B, st); }

extern "C" int dnmf_loss_grad(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                              int B_global, const float* beta_dev, const float* C_dev, float* grad_dev,
                              double* sse_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !C_dev || !grad_dev || !sse_dev)
    return fail("dnmf_loss_grad: NULL argument");
  if (B < 1 || B_global < B) return fail("dnmf_loss_grad: need 1 <= B <= B_global");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (check_sticky(c, "dnmf_loss_grad") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
  FitParams p;
  if (fill_fit_params(c, p, frames_dev, frame_ids_dev, B, beta_dev, C_dev, st)) return 1;
  const int nt = c->ntx * c->nty * c->ntz;
  if (launch_fused_fit(c, p, B, st)) return 1;
  const double scale = 2.0 / ((double)B_global * (double)c->N);
  reduce_partials_kernel<<<B, 256, 0, st>>>(c->d_partials, frame_ids_dev, B, nt, c->T, scale, grad_dev, sse_dev, nullptr,
                                            nullptr, c->d_id_flags);
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
  if (check_sticky(c, "dnmf_adam_step")) return 1;
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
  c->affine_call = affine;  // Adam freezes rows 4..9: the fused kernel need not produce their gradient
  const int rc = dnmf_loss_grad(c, frames_dev, frame_ids_dev, B, B_global, beta_dev, C_dev, c->d_grad, c->d_sse, stream);
  c->affine_call = 0;
  if (rc) return 1;
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
    c->affine_call = affine;
    const int frc = fill_fit_params(c, p, nullptr, frame_ids_dev + b_first, (int)Btot, beta_dev, C_dev, st);
    c->affine_call = 0;
    if (frc) return 1;
    const int nt = c->ntx * c->nty * c->ntz;
    if (launch_fused_fit(c, p, (int)Btot, st)) return 1;
    reduce_partials_kernel<<<(unsigned)Btot, 256, 0, st>>>(c->d_partials, frame_ids_dev + b_first, (int)Btot, nt, c->T,
                                                          0.0, c->d_grad, c->d_sse, nullptr, c->d_epoch_scale, nullptr);
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
  if (check_sticky(c, "dnmf_motion_step_host")) return 1;
  {  // the ids are on the host here: validate them before anything is launched
    std::vector<char> seen((size_t)c->T, 0);
    for (int b = 0; b < B; ++b) {
      const int t = frame_ids_host[b];
      if (t < 0 || t >= c->T)
        return fail("dnmf_motion_step_host: frame id " + std::to_string(t) + " outside [0, " + std::to_string(c->T) + ")");
      if (seen[(size_t)t])
        return fail("dnmf_motion_step_host: frame id " + std::to_string(t) +
                    " occurs twice in the batch (the chunked host path keeps one gradient column per frame; use "
                    "dnmf_motion_step with device frames for batches drawn with replacement)");
      seen[(size_t)t] = 1;
    }
  }
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
    c->affine_call = affine;
    const int rc = dnmf_loss_grad(c, buf, c->d_ids + b0, nb, B_global, beta_dev, C_dev, c->d_grad, c->d_sse + b0, stream);
    c->affine_call = 0;
    if (rc) return 1;
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
  if (check_sticky(c, "dnmf_forward") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
  if (AtC_dev) {
    // Yhat does not depend on the video: the WRITE_YHAT instantiation never reads `frames`.
    FitParams p;
    if (fill_fit_params(c, p, AtC_dev, frame_ids_dev, B, beta_dev, C_dev, st)) return 1;
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
