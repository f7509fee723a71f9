// Callers and data formats either side of the fit path, each as its own small kernel set:
//
//   render_cells_kernel        synthetic generator: clean frames of WUtils/Simulator.py:66-77,197-203
//   update_spatial_*           non-parametric footprint update, Demix/dNMF.py:151-160 (+ penalty D of :133-135)
//   dense_temporal_*           the static update_temporal on dense arrays, Demix/dNMF.py:139-149
//   forward_maxz_kernel        max-projection of the deformed footprints along z, demo.py:50-52 (A_t.max(2))
//   frames_maxz_kernel         the same for frames (Y.max(2), Y_i.max(2))
//
// The dense multiplicative updates work in fp64 like the reference's numpy einsum.
#include "dnmf_common.h"

namespace dnmf {

// ------------------------------------------------------------------------------------------------
// Synthetic generator.  video[t][x][y][z] = sum_k tr[k][t] * exp(-|p - P_k(t)|^2 / (2*shape_std)):
// Simulator.py:70-73 evaluates a scipy multivariate_normal pdf with cov = shape_std * I over ALL voxels for
// every (t, k) and rescales it to peak 1 (:203); the cell is separable, so a block of 16 x 16 (x, y) columns
// stages the three 1-D factors of a chunk of neurons for its frame (the x factor carries the trace) and every
// thread marches its column along z.  Neurons whose x or y factor underflows to zero over the whole tile are
// dropped while staging (their contribution is below 1e-38).  Every neuron of a chunk has its own slot and the columns
// add the live ones in ascending k: the video is bitwise reproducible (slots handed out by an atomic counter were not).
// ------------------------------------------------------------------------------------------------
constexpr int kRenTX = 16, kRenTY = 16, kRenChunk = 96, kRenMaxZ = 64;

__global__ void __launch_bounds__(256) render_cells_kernel(const float* __restrict__ pos /*[K][3][T]*/,
                                                           const float* __restrict__ tr /*[K][T]*/, int K, int T,
                                                           int t0, int X, int Y, int Z, float inv2s,
                                                           float* __restrict__ out /*[nT][X][Y][Z]*/) {
  __shared__ float sGX[kRenChunk][kRenTX], sGY[kRenChunk][kRenTY];
  __shared__ __align__(16) float sGZ[kRenChunk][kRenMaxZ];
  __shared__ unsigned char sLive[kRenChunk];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = tid & 15, ly = tid >> 4;
  const int x0 = blockIdx.x * kRenTX, y0 = blockIdx.y * kRenTY, tl = blockIdx.z, t = t0 + tl;
  const int Z4 = (Z + 3) & ~3;
  float acc[kRenMaxZ];
#pragma unroll
  for (int z = 0; z < kRenMaxZ; ++z) acc[z] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kRenChunk) {
    __syncthreads();
    if (tid < kRenChunk) sLive[tid] = 0;
    __syncthreads();
    // one warp per candidate neuron: factors of the tile, kept when the neuron reaches it at all
    for (int kk = warp; kk < kRenChunk && k0 + kk < K; kk += 8) {
      const int k = k0 + kk;
      const float px = pos[((size_t)k * 3 + 0) * T + t], py = pos[((size_t)k * 3 + 1) * T + t],
                  pz = pos[((size_t)k * 3 + 2) * T + t];
      const float c = tr[(size_t)k * T + t];
      float gx = 0.f, gy = 0.f;
      if (lane < 16) {
        const float d = (float)(x0 + lane) - px;
        gx = c * __expf(-d * d * inv2s);
      } else {
        const float d = (float)(y0 + lane - 16) - py;
        gy = __expf(-d * d * inv2s);
      }
      const bool live = __any_sync(0xffffffffu, gx != 0.f) && __any_sync(0xffffffffu, gy != 0.f);
      if (!live) continue;
      const int slot = kk;
      if (lane == 0) sLive[kk] = 1;
      if (lane < 16) sGX[slot][lane] = gx; else sGY[slot][lane - 16] = gy;
      for (int z = lane; z < Z4; z += 32) {
        const float d = (float)z - pz;
        sGZ[slot][z] = z < Z ? __expf(-d * d * inv2s) : 0.f;
      }
    }
    __syncthreads();
    const int n = min(kRenChunk, K - k0);
    for (int j = 0; j < n; ++j) {
      if (!sLive[j]) continue;  // block-uniform
      const float gxy = sGX[j][lx] * sGY[j][ly];
      const float4* gz = reinterpret_cast<const float4*>(sGZ[j]);
#pragma unroll
      for (int z4 = 0; z4 < kRenMaxZ / 4; ++z4) {
        if (4 * z4 < Z) {
          const float4 g = gz[z4];
          acc[4 * z4 + 0] = fmaf(gxy, g.x, acc[4 * z4 + 0]);
          acc[4 * z4 + 1] = fmaf(gxy, g.y, acc[4 * z4 + 1]);
          acc[4 * z4 + 2] = fmaf(gxy, g.z, acc[4 * z4 + 2]);
          acc[4 * z4 + 3] = fmaf(gxy, g.w, acc[4 * z4 + 3]);
        }
      }
    }
  }
  const int gx_ = x0 + lx, gy_ = y0 + ly;
  if (gx_ < X && gy_ < Y) {
    float* dst = out + (((size_t)tl * X + gx_) * Y + gy_) * Z;
#pragma unroll
    for (int z = 0; z < kRenMaxZ; ++z)
      if (z < Z) dst[z] = acc[z];
  }
}

// ------------------------------------------------------------------------------------------------
// update_spatial (Demix/dNMF.py:151-160), voxels p = flattened leading axes:
//   C_s = C C^T;  A1[p,k] = sum_t Y_i[p,t] C[k,t];  A2[p,k] = sum_l A[p,l] C_s[l,k] (+ gamma D[p,k]);
//   A <- A * A1 / (A2 + 1e-32)
// One pass over the registered video: a block owns 64 voxels, streams their Y_i rows once through shared
// memory in t chunks and keeps the [64 x K] accumulators of A1 in registers (8 neurons per thread, K tiled by
// 32), then forms A2 from the block's A rows and the K x K matrix C_s.  D is either the caller's array or
// computed on the fly from the neuron positions, D = 1 - exp(-0.01 |p - pos_k|) (:133-135), never stored.
// ------------------------------------------------------------------------------------------------
__global__ void gram_rows_kernel(const double* __restrict__ C, int K, int T, double* __restrict__ Cs) {
  // Cs[k][l] = sum_t C[k][t] C[l][t]; one warp per entry, fixed summation order
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (idx >= K * K) return;
  const int k = idx / K, l = idx - k * K;
  double acc = 0.0;
  for (int t = lane; t < T; t += 32) acc = fma(C[(size_t)k * T + t], C[(size_t)l * T + t], acc);
  acc = warp_sum_d(acc);
  if (lane == 0) Cs[idx] = acc;
}

constexpr int kUsVox = 64, kUsKT = 32, kUsTC = 32;

__global__ void __launch_bounds__(256) update_spatial_kernel(const double* __restrict__ A, const double* __restrict__ C,
                                                             const double* __restrict__ Yi, const double* __restrict__ Cs,
                                                             const double* __restrict__ D, const float* __restrict__ pos,
                                                             int gX, int gY, int gZ, double gamma, int use_D,
                                                             long long P, int K, int T, double* __restrict__ out) {
  __shared__ double sY[kUsVox][kUsTC + 1];
  __shared__ double sC[kUsKT][kUsTC + 1];
  const int tid = threadIdx.x, v = tid & 63, kg = tid >> 6;  // 4 groups of 8 neurons
  const long long p0 = (long long)blockIdx.x * kUsVox, p = p0 + v;
  for (int kt = 0; kt < K; kt += kUsKT) {
    double a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a1[i] = 0.0;
    for (int tc = 0; tc < T; tc += kUsTC) {
      __syncthreads();
      for (int e = tid; e < kUsVox * kUsTC; e += 256) {
        const int vv = e / kUsTC, tt = e - vv * kUsTC;
        sY[vv][tt] = (p0 + vv < P && tc + tt < T) ? Yi[(size_t)(p0 + vv) * T + tc + tt] : 0.0;
      }
      for (int e = tid; e < kUsKT * kUsTC; e += 256) {
        const int kk = e / kUsTC, tt = e - kk * kUsTC;
        sC[kk][tt] = (kt + kk < K && tc + tt < T) ? C[(size_t)(kt + kk) * T + tc + tt] : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int tt = 0; tt < kUsTC; ++tt) {
        const double y = sY[v][tt];
#pragma unroll
        for (int i = 0; i < 8; ++i) a1[i] = fma(y, sC[kg * 8 + i][tt], a1[i]);
      }
    }
    if (p < P) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kt + kg * 8 + i;
        if (k >= K) continue;
        double a2 = 0.0;
        for (int l = 0; l < K; ++l) a2 = fma(A[(size_t)p * K + l], Cs[(size_t)l * K + k], a2);
        if (use_D == 1) {
          a2 += gamma * D[(size_t)p * K + k];
        } else if (use_D == 2) {
          const long long z = p % gZ, y = (p / gZ) % gY, x = p / ((long long)gZ * gY);
          const double dx = (double)x - (double)pos[k * 3 + 0], dy = (double)y - (double)pos[k * 3 + 1],
                       dz = (double)z - (double)pos[k * 3 + 2];
          a2 += gamma * (1.0 - exp(-0.01 * sqrt(dx * dx + dy * dy + dz * dz)));
        }
        const double a = A[(size_t)p * K + k];
        out[(size_t)p * K + k] = a * a1[i] / (a2 + 1e-32);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Static update_temporal on dense arrays (Demix/dNMF.py:139-149): A_t[p][k][t] (T innermost, the reference's
// [X,Y,Z,K,T] array), Y[p][t], C[k][t].  Statistics per frame in fp64 with a fixed summation order (one warp per
// (k, l, t) triple walks the voxels in order), then the elementwise multiplicative update.
// ------------------------------------------------------------------------------------------------
__global__ void dense_temporal_stats_kernel(const double* __restrict__ At, const double* __restrict__ Y, long long P,
                                            int K, int T, double* __restrict__ G /*[K][K][T]*/,
                                            double* __restrict__ b /*[K][T]*/) {
  // thread = (k, l, t) with t fastest: consecutive threads read consecutive t of the same (p, k) row
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)K * (K + 1) * T;  // l == K: the b column
  if (idx >= total) return;
  const int t = (int)(idx % T);
  const int l = (int)((idx / T) % (K + 1));
  const int k = (int)(idx / ((long long)T * (K + 1)));
  double acc = 0.0;
  if (l < K) {
    for (long long p = 0; p < P; ++p) acc = fma(At[((size_t)p * K + k) * T + t], At[((size_t)p * K + l) * T + t], acc);
    G[((size_t)k * K + l) * T + t] = acc;
  } else {
    for (long long p = 0; p < P; ++p) acc = fma(At[((size_t)p * K + k) * T + t], Y[(size_t)p * T + t], acc);
    b[(size_t)k * T + t] = acc;
  }
}

__global__ void dense_temporal_update_kernel(const double* __restrict__ G, const double* __restrict__ b,
                                             const double* __restrict__ C, int K, int T, double gamma, int use_gamma,
                                             double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * T) return;
  const int k = idx / T, t = idx - k * T;
  double c2 = 0.0;
  for (int l = 0; l < K; ++l) c2 = fma(G[((size_t)k * K + l) * T + t], C[(size_t)l * T + t], c2);
  double c1 = b[idx];
  const double ck = C[idx];
  if (use_gamma) {
    const double prev = t > 0 ? C[idx - 1] : ck, next = t < T - 1 ? C[idx + 1] : ck;
    c1 += gamma * (prev + next);
    c2 += 2.0 * gamma * ck;
  }
  out[idx] = ck * c1 / (c2 + 1e-32);
}

// ------------------------------------------------------------------------------------------------
// Max-projections along z (demo.py:50-52).  forward_maxz: out[b][k][x][y] = max_z A_t(p, k) straight from the
// per-axis tables -- the dense A_t[B,K,X,Y,Z] is never formed.  One thread per (b, x, y) column: the cells of
// its Z samples are computed once, every neuron is first tested against the column's cell ranges (its truncated
// footprint cannot reach most columns) and only then evaluated.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxzZ = 64;

__global__ void forward_maxz_kernel(Geom g, const int* __restrict__ frame_ids, int B, const float* __restrict__ beta,
                                    const float2* __restrict__ tab0, const float2* __restrict__ tab1,
                                    const float2* __restrict__ tab2, const int* __restrict__ rng,
                                    float* __restrict__ out) {
  const size_t XY = (size_t)g.X * g.Y;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= XY * B) return;
  const int b = (int)(idx / XY);
  const size_t c = idx - (size_t)b * XY;
  const int y = (int)(c % g.Y), x = (int)(c / g.Y);
  const int t = frame_ids[b];
  const float xf = (float)x, yf = (float)y;
  const int sz[3] = {g.X, g.Y, g.Z};
  short ii[3][kMaxzZ];
  float ff[3][kMaxzZ];
  int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
  for (int z = 0; z < g.Z; ++z) {
    const float zf = (float)z;
    const float phi[kBasis] = {1.f, xf, yf, zf, xf * xf, yf * yf, zf * zf, xf * yf, xf * zf, yf * zf};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float q = 0.f;
#pragma unroll
      for (int a = 0; a < kBasis; ++a) q = __fadd_rn(q, __fmul_rn(phi[a], beta[((size_t)a * 3 + d) * g.T + t]));
      int i;
      float f;
      split_coord(sample_coord(q, (float)(sz[d] - 1)), sz[d], i, f);
      ii[d][z] = (short)i;
      ff[d][z] = f;
      lo[d] = min(lo[d], i);
      hi[d] = max(hi[d], i);
    }
  }
  for (int k = 0; k < g.K; ++k) {
    const int* r = rng + (size_t)k * 6;
    float m = 0.f;
    const bool reach = r[0] <= r[1] && r[2] <= r[3] && r[4] <= r[5] && r[0] - 1 <= hi[0] && r[1] >= lo[0] &&
                       r[2] - 1 <= hi[1] && r[3] >= lo[1] && r[4] - 1 <= hi[2] && r[5] >= lo[2];
    if (reach) {
      for (int z = 0; z < g.Z; ++z) {
        const float2 ex = tab0[(size_t)k * (g.X + 3) + ii[0][z] + 2];
        const float2 ey = tab1[(size_t)k * (g.Y + 3) + ii[1][z] + 2];
        const float2 ez = tab2[(size_t)k * (g.Z + 3) + ii[2][z] + 2];
        m = fmaxf(m, fmaf(ff[0][z], ex.y, ex.x) * (fmaf(ff[1][z], ey.y, ey.x) * fmaf(ff[2][z], ez.y, ez.x)));
      }
    }
    out[((size_t)b * g.K + k) * XY + c] = m;
  }
}

__global__ void frames_maxz_kernel(const float* __restrict__ frames, size_t columns, int Z, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= columns) return;
  const float* src = frames + idx * Z;
  float m = src[0];
  for (int z = 1; z < Z; ++z) m = fmaxf(m, src[z]);
  out[idx] = m;
}

}  // namespace dnmf

using namespace dnmf;

extern "C" int dnmf_render_cells(const float* pos_dev, const float* traces_dev, int K, int T, int t0, int nT, int X,
                                 int Y, int Z, float shape_std, float* out_dev, void* stream) {
  if (!pos_dev || !traces_dev || !out_dev) return fail("dnmf_render_cells: NULL argument");
  if (K < 1 || T < 1 || t0 < 0 || nT < 1 || t0 + nT > T || X < 1 || Y < 1 || Z < 1)
    return fail("dnmf_render_cells: bad sizes");
  if (Z > kRenMaxZ) return fail("dnmf_render_cells: Z > 64 is not supported by the column march");
  if (!(shape_std > 0.f)) return fail("dnmf_render_cells: shape_std must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((X + kRenTX - 1) / kRenTX), gy = (unsigned)((Y + kRenTY - 1) / kRenTY);
  if (gy > 65535) return fail("dnmf_render_cells: more than 65535 tiles along y");
  for (int b0 = 0; b0 < nT; b0 += 65535) {
    const int nb = std::min(65535, nT - b0);
    render_cells_kernel<<<dim3(gx, gy, (unsigned)nb), 256, 0, st>>>(
        pos_dev, traces_dev, K, T, t0 + b0, X, Y, Z, 1.f / (2.f * shape_std), out_dev + (size_t)b0 * X * Y * Z);
    CU(cudaGetLastError());
  }
  return 0;
}

extern "C" int dnmf_update_spatial(const double* A_dev, const double* C_dev, const double* Yi_dev, const double* D_dev,
                                   const float* pos_dev, int gX, int gY, int gZ, double gamma, int use_D,
                                   int64_t P, int K, int T, double* scratch_KK_dev, double* out_dev, void* stream) {
  if (!A_dev || !C_dev || !Yi_dev || !scratch_KK_dev || !out_dev) return fail("dnmf_update_spatial: NULL argument");
  if (P < 1 || K < 1 || T < 1) return fail("dnmf_update_spatial: bad sizes");
  if (use_D == 1 && !D_dev) return fail("dnmf_update_spatial: use_D = 1 needs D_dev");
  if (use_D == 2 && (!pos_dev || (int64_t)gX * gY * gZ != P))
    return fail("dnmf_update_spatial: use_D = 2 needs positions and a grid with X*Y*Z = P voxels");
  if (use_D < 0 || use_D > 2) return fail("dnmf_update_spatial: use_D must be 0, 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  gram_rows_kernel<<<(unsigned)(((size_t)K * K + 7) / 8), 256, 0, st>>>(C_dev, K, T, scratch_KK_dev);
  CU(cudaGetLastError());
  update_spatial_kernel<<<(unsigned)((P + kUsVox - 1) / kUsVox), 256, 0, st>>>(
      A_dev, C_dev, Yi_dev, scratch_KK_dev, D_dev, pos_dev, gX, gY, gZ, gamma, use_D, (long long)P, K, T, out_dev);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_update_temporal_dense(const double* At_dev, const double* C_dev, const double* Y_dev, double gamma,
                                          int use_gamma, int64_t P, int K, int T, double* scratch_dev, double* out_dev,
                                          void* stream) {
  if (!At_dev || !C_dev || !Y_dev || !scratch_dev || !out_dev) return fail("dnmf_update_temporal_dense: NULL argument");
  if (P < 1 || K < 1 || T < 1) return fail("dnmf_update_temporal_dense: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  double* G = scratch_dev;                       // [K][K][T]
  double* b = scratch_dev + (size_t)K * K * T;   // [K][T]
  const long long total = (long long)K * (K + 1) * T;
  dense_temporal_stats_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(At_dev, Y_dev, (long long)P, K, T, G, b);
  CU(cudaGetLastError());
  dense_temporal_update_kernel<<<(unsigned)(((size_t)K * T + 127) / 128), 128, 0, st>>>(G, b, C_dev, K, T, gamma,
                                                                                       use_gamma, out_dev);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_forward_maxz(dnmf_ctx* c, const int32_t* frame_ids_dev, int B, const float* beta_dev,
                                 float* out_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !out_dev) return fail("dnmf_forward_maxz: NULL argument");
  if (!c->have_footprints) return fail("dnmf_forward_maxz: call dnmf_set_footprints first");
  if (c->Z > kMaxzZ) return fail("dnmf_forward_maxz: Z > 64 is not supported");
  if (B < 1) return fail("dnmf_forward_maxz: B must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  const size_t total = (size_t)c->X * c->Y * B;
  forward_maxz_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(geom_of(c), frame_ids_dev, B, beta_dev,
                                                                      c->d_tab[0], c->d_tab[1], c->d_tab[2], c->d_rng,
                                                                      out_dev);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_frames_maxz(const float* frames_dev, int64_t columns, int Z, float* out_dev, void* stream) {
  if (!frames_dev || !out_dev || columns < 1 || Z < 1) return fail("dnmf_frames_maxz: bad argument");
  frames_maxz_kernel<<<(unsigned)((columns + 255) / 256), 256, 0, (cudaStream_t)stream>>>(frames_dev, (size_t)columns,
                                                                                         Z, out_dev);
  CU(cudaGetLastError());
  return 0;
}
