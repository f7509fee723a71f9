// Trace update (kernel 3b) -- included at the end of dnmf_kernels.cu (same translation unit).
//
// mu_stats_kernel: per (frame, 16x8xtz tile) CTA of 128 threads.  Phase 1: every thread evaluates
// the closed-form footprint value A_j(p) of each listed neuron at its voxel and parks it in shared
// memory next to the voxel's Y value (pseudo-neuron).  Phase 2: the CTA computes the small SYRK
// [A|Y]^T [A|Y] of those 128 voxels with 4x4 register blocks and split-K over the voxel index.
// Accumulators live in registers for the whole tile and are flushed once with fp64 atomics into
// the dense per-frame G_t[K][K], b_t[K].   Reference: Demix/dNMF.py:141-142 (fp64 einsum).
namespace dnmf {

constexpr int kMuThreads = 128;
constexpr int kMuTX = 16, kMuTY = 8;
constexpr int kMuVS = 129;  // row stride of the A panel in floats (odd: conflict-free column walks)

struct MuParams {
  const float* frames;
  const int* frame_ids;
  const float* beta;
  const float2* tab0;
  const float2* tab1;
  const float2* tab2;
  const int* rng;
  double* G;
  double* bvec;
  int* overflow;
  int frames_are_batch;
  int X, Y, Z, K, T;
  int tz, ntx, nty, ntz;
  int capM;  // rows available in the A panel (multiple of 4), must be >= L + 1
  int full_depth;
};

// BS = 4: panel [row][voxel] (stride kMuVS), 4x4 register blocks -- short lists.
// BS = 8: panel [voxel][row] (stride capM + 4 floats, 16-byte aligned rows), 8x8 register blocks fed by four
// LDS.128 per voxel (64 FMAs per 4 shared-memory instructions instead of 16 per 8 scalar ones, whose row
// stride put the 4x4 variant's lanes on 8 banks) -- long lists (cfg4: ~100 neurons per tile).
static size_t mu_smem_bytes(int capM, int tz, int K, int bs) {
  const int mb = (capM + bs - 1) / bs;
  const size_t panel = bs == 8 ? (size_t)kMuThreads * (capM + 4) * 4 : (size_t)capM * kMuVS * 4;
  return panel + (size_t)(kMuTX * kMuTY * tz + 4) * 4 + 64 * 4 +
         (size_t)((K + 7) & ~7) * 2 + (size_t)((mb * (mb + 1) / 2 + 3) & ~3) * 2;
}

template <int R, int BS>
__global__ void __launch_bounds__(kMuThreads) mu_stats_kernel(const __grid_constant__ MuParams p) {
  constexpr int NW = kMuThreads / 32, NWX = 2;
  constexpr int TX = kMuTX, TY = kMuTY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);
  const int PS = p.capM + 4;  // BS == 8: floats per voxel row of the panel (capM is a multiple of 8 there)
  const int panel_floats = BS == 8 ? kMuThreads * PS : p.capM * kMuVS;
  float* sY = sA + (size_t)panel_floats;
  float* sBeta = sY + (TX * TY * p.tz + 4);
  int* sInt = reinterpret_cast<int*>(sBeta + 32);
  unsigned short* sList = reinterpret_cast<unsigned short*>(sInt + 32);
  unsigned short* sBlk = sList + ((p.K + 7) & ~7);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = p.ntx * p.nty * p.ntz;
  const int b = blockIdx.x / nt, tile = blockIdx.x - b * nt;
  const int bx = tile % p.ntx, by = (tile / p.ntx) % p.nty, bz = tile / (p.ntx * p.nty);
  const int t = p.frame_ids[b];
  const int x0 = bx * TX, y0 = by * TY, z0 = bz * p.tz;
  const int nx = min(TX, p.X - x0), ny = min(TY, p.Y - y0), nz = min(p.tz, p.Z - z0);
  const float* __restrict__ frame = p.frames + (size_t)(p.frames_are_batch ? b : t) * ((size_t)p.X * p.Y * p.Z);
  const int zs = p.full_depth ? p.Z : p.tz;
  const int RS = TY * zs;

  if (tid < 30) sBeta[tid] = p.beta[(size_t)tid * p.T + t];
  if (p.full_depth) {
    const int run = ny * p.Z;
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
  for (int e = tid; e < panel_floats; e += kMuThreads) sA[e] = 0.f;
  __syncthreads();
  if (tid < 3) {
    const int s = tid == 0 ? p.X : (tid == 1 ? p.Y : p.Z);
    int wlo, whi;
    tile_window_axis(sBeta + tid, 3, (float)x0, (float)y0, (float)z0, (float)(x0 + nx - 1),
                     (float)(y0 + ny - 1), (float)(z0 + nz - 1), s, wlo, whi);
    sInt[tid] = wlo;
    sInt[3 + tid] = whi;
  }
  __syncthreads();
  int wlo[3], whi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wlo[d] = sInt[d];
    whi[d] = sInt[3 + d];
  }
  const int per = ((p.K + NW * 32 - 1) / (NW * 32)) * 32;
  const int kb = warp * per;
  {
    int cnt = 0;
    for (int k0 = kb; k0 < kb + per; k0 += 32) {
      int k = k0 + lane;
      bool ok = (k < p.K) && neuron_in_window(p.rng + (size_t)k * 6, wlo, whi);
      cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) sInt[8 + warp] = cnt;
  }
  __syncthreads();
  int L = 0;
  {
    int off = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      int c = sInt[8 + w];
      if (w < warp) off += c;
      L += c;
    }
    for (int k0 = kb; k0 < kb + per; k0 += 32) {
      int k = k0 + lane;
      bool ok = (k < p.K) && neuron_in_window(p.rng + (size_t)k * 6, wlo, whi);
      unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) sList[off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)k;
      off += __popc(m);
    }
  }
  if (L + 1 > p.capM) {  // loud failure: the host checks this flag after the launch
    if (tid == 0) atomicMax(p.overflow, L + 1);
    return;
  }
  const int M = L + 1;               // neurons + the Y pseudo-neuron
  const int mb = (M + BS - 1) / BS;  // BS x BS blocks per side
  const int nblk = mb * (mb + 1) / 2;
  for (int e = tid; e < nblk; e += kMuThreads) {  // decode triangular block index -> (bi, bj), bi <= bj
    int bi = 0, rem = e;
    while (rem >= mb - bi) {
      rem -= mb - bi;
      ++bi;
    }
    sBlk[e] = (unsigned short)((bi << 8) | (bi + rem));
  }
  __syncthreads();

  int ks = 1;
  while (ks < 32 && nblk * (ks * 2) <= kMuThreads) ks *= 2;
  const int groups = kMuThreads / ks;  // blocks processed concurrently per register slot
  const int grp = tid / ks, ksl = tid - grp * ks;

  const int lx = (warp % NWX) * kWarpX + (lane & 7);
  const int ly = (warp / NWX) * kWarpY + (lane >> 3);
  const int gx = x0 + lx, gy = y0 + ly;
  const bool valid = (gx < p.X) && (gy < p.Y);
  const float xf = (float)gx, yf = (float)gy;
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
  }
  const float sm1x = (float)(p.X - 1), sm1y = (float)(p.Y - 1), sm1z = (float)(p.Z - 1);
  const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
  const int ybase = lx * RS + ly * zs;
  const size_t gbase = (size_t)t * p.K * p.K;

  for (int base = 0; base < nblk; base += groups * R) {
    float acc[R][BS * BS];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < BS * BS; ++i) acc[r][i] = 0.f;

    for (int zz = 0; zz < nz; ++zz) {
      // phase 1: A_j(p) for every listed neuron at this thread's voxel
      const float zf = (float)(z0 + zz);
      int i0, i1, i2;
      float f0, f1, f2;
      split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[0], c1[0]), c0[0]), sm1x), p.X, i0, f0);
      split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[1], c1[1]), c0[1]), sm1y), p.Y, i1, f1);
      split_coord(sample_coord(fmaf(zf, fmaf(zf, c2[2], c1[2]), c0[2]), sm1z), p.Z, i2, f2);
      for (int j = 0; j < L; ++j) {
        const int k = sList[j];
        float2 ex = __ldg(p.tab0 + (size_t)k * sX3 + (i0 + 2));
        float2 ey = __ldg(p.tab1 + (size_t)k * sY3 + (i1 + 2));
        float2 ez = __ldg(p.tab2 + (size_t)k * sZ3 + (i2 + 2));
        float a = (fmaf(f0, ex.y, ex.x) * fmaf(f1, ey.y, ey.x)) * fmaf(f2, ez.y, ez.x);
        sA[BS == 8 ? tid * PS + j : j * kMuVS + tid] = valid ? a : 0.f;
      }
      sA[BS == 8 ? tid * PS + L : L * kMuVS + tid] = valid ? sY[ybase + zz] : 0.f;
      __syncthreads();
      // phase 2: BS x BS register blocks of [A|Y]^T [A|Y], split-K over the 128 voxels
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int blk = base + r * groups + grp;
        if (blk < nblk) {
          const int code = sBlk[blk];
          if constexpr (BS == 8) {
            const float4* ra = reinterpret_cast<const float4*>(sA + (code >> 8) * 8);
            const float4* rb = reinterpret_cast<const float4*>(sA + (code & 255) * 8);
            const int ps4 = PS >> 2;
#pragma unroll 2
            for (int v = ksl; v < kMuThreads; v += ks) {
              const float4 a0 = ra[v * ps4], a1 = ra[v * ps4 + 1], b0 = rb[v * ps4], b1 = rb[v * ps4 + 1];
              const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][i * 8 + c] = fmaf(a[i], bb[c], acc[r][i * 8 + c]);
            }
          } else {
            const float* ra = sA + (size_t)((code >> 8) * 4) * kMuVS;
            const float* rb = sA + (size_t)((code & 255) * 4) * kMuVS;
            for (int v = ksl; v < kMuThreads; v += ks) {
              float a[4], bb[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                a[i] = ra[i * kMuVS + v];
                bb[i] = rb[i * kMuVS + v];
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][i * 4 + c] = fmaf(a[i], bb[c], acc[r][i * 4 + c]);
            }
          }
        }
      }
      __syncthreads();
    }

    // flush: reduce the split-K lanes, then fp64 atomics into G_t / b_t
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int blk = base + r * groups + grp;
#pragma unroll
      for (int i = 0; i < BS * BS; ++i) {
        float v = acc[r][i];
        for (int o = ks >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][i] = v;
      }
      if (blk < nblk && ksl == 0) {
        const int code = sBlk[blk];
        const int bi = code >> 8, bj = code & 255;
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          const int jr = bi * BS + i;
          if (jr >= L) continue;  // Y pseudo-row or padding
          const int kr = sList[jr];
#pragma unroll
          for (int c = 0; c < BS; ++c) {
            const int jc = bj * BS + c;
            const double v = (double)acc[r][i * BS + c];
            if (jc < L) {
              const int kc = sList[jc];
              atomicAdd(p.G + gbase + (size_t)kr * p.K + kc, v);
              if (bi != bj) atomicAdd(p.G + gbase + (size_t)kc * p.K + kr, v);
            } else if (jc == L) {
              atomicAdd(p.bvec + (size_t)t * p.K + kr, v);
            }
          }
        }
      }
    }
  }
}

__global__ void mu_zero_kernel(double* __restrict__ G, double* __restrict__ bvec, const int* __restrict__ frame_ids,
                               int K) {
  const int t = frame_ids[blockIdx.x];
  double* g = G + (size_t)t * K * K;
  for (size_t i = threadIdx.x; i < (size_t)K * K; i += blockDim.x) g[i] = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) bvec[(size_t)t * K + i] = 0.0;
}

__global__ void mu_load_kernel(const float* __restrict__ C, double* __restrict__ Cd, int K, int T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)K * T) return;
  const int t = (int)(i / K), k = (int)(i - (size_t)t * K);
  Cd[i] = (double)C[(size_t)k * T + t];
}

__global__ void mu_store_kernel(const double* __restrict__ Cd, float* __restrict__ C, int K, int T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)K * T) return;
  const int k = (int)(i / T), t = (int)(i - (size_t)k * T);
  C[i] = (float)Cd[(size_t)t * K + k];
}

// One warp per (t, k): C <- C (b + g (C[t-1] + C[t+1])) / (G C + 2 g C + 1e-32), Jacobi over frames.
__global__ void mu_sweep_kernel(const double* __restrict__ G, const double* __restrict__ bvec,
                                const double* __restrict__ Cin, double* __restrict__ Cout, int T, int K,
                                double gamma, int use_gamma, const double* __restrict__ halo_prev,
                                const double* __restrict__ halo_next) {
  // grid = (T, ceil(K / warps per block))
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x, k = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K) return;
  const double* g = G + ((size_t)t * K + k) * K;
  const double* c = Cin + (size_t)t * K;
  double dot = 0.0;
  for (int l = lane; l < K; l += 32) dot = fma(g[l], c[l], dot);
  dot = warp_sum_d(dot);
  if (lane == 0) {
    const double ck = c[k];
    double c1 = bvec[(size_t)t * K + k];
    double c2 = dot;
    if (use_gamma) {
      const double prev = t > 0 ? Cin[(size_t)(t - 1) * K + k] : (halo_prev ? halo_prev[k] : ck);
      const double next = t < T - 1 ? Cin[(size_t)(t + 1) * K + k] : (halo_next ? halo_next[k] : ck);
      c1 += gamma * (prev + next);
      c2 += 2.0 * gamma * ck;
    }
    Cout[(size_t)t * K + k] = ck * c1 / (c2 + 1e-32);
  }
}

// Sparse sweeps.  G_t[k][l] can only be non-zero when the truncated supports of neurons k and l overlap, which
// does not depend on the deformation: the static neighbour list nbr[k][0..W) (ascending, -1 padded).  One warp per
// (t, k) row gathers the row's neighbour entries from the dense statistics into Gc[t][k][0..W) and checks that
// every other entry of the row is exactly zero (flag otherwise: the caller keeps the dense sweeps).
__global__ void mu_compact_kernel(const double* __restrict__ G, const int* __restrict__ nbr, int W, int T, int K,
                                  double* __restrict__ Gc, int* __restrict__ violation) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)T * K) return;
  const int k = (int)(row % K);
  const double* g = G + (size_t)row * K;
  const int* nb = nbr + (size_t)k * W;
  bool bad = false;
  for (int l = lane; l < K; l += 32) {
    if (g[l] != 0.0) {
      int lo = 0, hi = W - 1;  // is l in nb[] ? (-1 padding sorts last: treat as +inf)
      bool found = false;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = nb[mid];
        if (v == l) {
          found = true;
          break;
        }
        if (v < 0 || v > l) hi = mid - 1; else lo = mid + 1;
      }
      bad |= !found;
    }
  }
  if (bad) atomicOr(violation, 1);
  for (int s_ = lane; s_ < W; s_ += 32) {
    const int l = nb[s_];
    Gc[(size_t)row * W + s_] = l >= 0 ? g[l] : 0.0;
  }
}

// Ws lanes per neuron k (the power of two covering the row length W, at most 32); one thread walks kMuTB
// consecutive frames of its (k, list slot): the neighbour id is loaded once and the kMuTB (G, C) pairs are
// independent loads in flight (one row per warp and frame was latency-bound: nbr -> C gather -> reduce -> store).
// grid = (ceil(T / kMuTB), ceil(K / rows per block)).
constexpr int kMuTB = 4;
__global__ void mu_sweep_sparse_kernel(const double* __restrict__ Gc, const int* __restrict__ nbr, int W, int Ws,
                                       const double* __restrict__ bvec, const double* __restrict__ Cin,
                                       double* __restrict__ Cout, int T, int K, double gamma, int use_gamma,
                                       const double* __restrict__ halo_prev, const double* __restrict__ halo_next) {
  const int sub = threadIdx.x & (Ws - 1);
  const int wshift = 31 - __clz(Ws);
  const int t0 = blockIdx.x * kMuTB;
  const int k = blockIdx.y * (blockDim.x >> wshift) + (threadIdx.x >> wshift);
  const bool live = k < K;
  double dot[kMuTB];
#pragma unroll
  for (int u = 0; u < kMuTB; ++u) dot[u] = 0.0;
  // the epilogue's operands of frame t0 + sub (lanes sub < kMuTB), requested before the dot products
  const int te = t0 + sub;
  const bool epi = live && sub < kMuTB && te < T;
  double ck = 0.0, c1 = 0.0, prev = 0.0, next = 0.0;
  if (epi) {
    ck = Cin[(size_t)te * K + k];
    c1 = bvec[(size_t)te * K + k];
    if (use_gamma) {
      prev = te > 0 ? Cin[(size_t)(te - 1) * K + k] : (halo_prev ? halo_prev[k] : ck);
      next = te < T - 1 ? Cin[(size_t)(te + 1) * K + k] : (halo_next ? halo_next[k] : ck);
    }
  }
  if (live) {
    for (int s_ = sub; s_ < W; s_ += Ws) {
      const int l = max(nbr[(size_t)k * W + s_], 0);  // padding slots hold G = 0
#pragma unroll
      for (int u = 0; u < kMuTB; ++u) {
        const int t = min(t0 + u, T - 1);
        dot[u] = fma(Gc[((size_t)t * K + k) * W + s_], Cin[(size_t)t * K + l], dot[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kMuTB; ++u)
    for (int o = Ws >> 1; o > 0; o >>= 1) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
  if (epi) {
    double c2 = dot[0];
#pragma unroll
    for (int u = 1; u < kMuTB; ++u) c2 = sub == u ? dot[u] : c2;
    if (use_gamma) {
      c1 += gamma * (prev + next);
      c2 += 2.0 * gamma * ck;
    }
    Cout[(size_t)te * K + k] = ck * c1 / (c2 + 1e-32);
  }
}

// All `iters` sweeps of one frame in one CTA, for the update without temporal coupling (gamma = None or 0:
// demo.py:46 runs update_footprints(gamma_c=0, iter_c=50)): the frames are independent.  The CTA stages its
// frame's compacted statistics and the neighbour ids in shared memory, transposed to [slot][k] so that one
// thread per neuron walks its row conflict-free with four independent accumulators (no shuffles, no global
// memory inside the sweep loop); traces and b_t live in shared memory, one barrier per sweep.
__global__ void mu_sweeps_local_kernel(const double* __restrict__ Gc, const int* __restrict__ nbr, int W,
                                       const double* __restrict__ bvec, const double* __restrict__ Cin,
                                       double* __restrict__ Cout, int K, int iters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sG = reinterpret_cast<double*>(smem_raw);  // [W][K]
  double* sC0 = sG + (size_t)W * K;
  double* sC1 = sC0 + K;
  double* sB = sC1 + K;
  int* sN = reinterpret_cast<int*>(sB + K);           // [W][K]
  const int t = blockIdx.x;
  const double* g = Gc + (size_t)t * K * W;
  for (int e = threadIdx.x; e < K * W; e += blockDim.x) {
    const int k = e / W, s_ = e - k * W;
    sG[s_ * K + k] = g[e];
    sN[s_ * K + k] = max(nbr[e], 0);  // padding slots hold G = 0
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    sC0[k] = Cin[(size_t)t * K + k];
    sB[k] = bvec[(size_t)t * K + k];
  }
  __syncthreads();
  double* cur = sC0;
  double* nxt = sC1;
  for (int it = 0; it < iters; ++it) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
      for (int s_ = 0; s_ < W; s_ += 4) {  // W is a multiple of 4
        d0 = fma(sG[(s_ + 0) * K + k], cur[sN[(s_ + 0) * K + k]], d0);
        d1 = fma(sG[(s_ + 1) * K + k], cur[sN[(s_ + 1) * K + k]], d1);
        d2 = fma(sG[(s_ + 2) * K + k], cur[sN[(s_ + 2) * K + k]], d2);
        d3 = fma(sG[(s_ + 3) * K + k], cur[sN[(s_ + 3) * K + k]], d3);
      }
      nxt[k] = cur[k] * sB[k] / (((d0 + d1) + (d2 + d3)) + 1e-32);
    }
    __syncthreads();
    double* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) Cout[(size_t)t * K + k] = cur[k];
}

__global__ void mu_boundary_kernel(const double* __restrict__ Cd, int T, int K, double* first, double* last) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  if (first) first[k] = Cd[k];
  if (last) last[k] = Cd[(size_t)(T - 1) * K + k];
}

// Nearest-neighbour registration (ExponentialFP.image_iwarp, Demix/dNMF.py:81-83,95-103): the
// reference scatters frame values at the deformed points f(p) = ((u+1)/2)*s (note: s, not s-1) and
// reads the nearest scattered point at every integer voxel.  Here every scattered point votes for
// the integer voxels in its 3x3x3 neighbourhood with a packed (distance, source index) key and an
// atomicMin; voxels that receive no vote fall back to an exact brute-force search.
__global__ void iwarp_vote_kernel(Geom g, const int* __restrict__ frame_ids, int B, const float* __restrict__ beta,
                                  unsigned long long* __restrict__ keys) {
  const size_t N = (size_t)g.X * g.Y * g.Z;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * B) return;
  const int b = (int)(idx / N);
  const size_t v = idx - (size_t)b * N;
  const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((size_t)g.Z * g.Y));
  const int t = frame_ids[b];
  const float xf = (float)x, yf = (float)y, zf = (float)z;
  const float phi[kBasis] = {1.f, xf, yf, zf, xf * xf, yf * yf, zf * zf, xf * yf, xf * zf, yf * zf};
  const int sz[3] = {g.X, g.Y, g.Z};
  float f[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float q = 0.f;
#pragma unroll
    for (int a = 0; a < kBasis; ++a) q = __fadd_rn(q, __fmul_rn(phi[a], beta[((size_t)a * 3 + d) * g.T + t]));
    const float sm1 = (float)(sz[d] - 1);
    float u = sm1 == 0.f ? 0.f : __fsub_rn(__fdiv_rn(__fmul_rn(2.f, q), sm1), 1.f);
    f[d] = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)sz[d]);
  }
  const int cx = (int)floorf(fminf(fmaxf(f[0], -2.f), (float)g.X + 1.f));
  const int cy = (int)floorf(fminf(fmaxf(f[1], -2.f), (float)g.Y + 1.f));
  const int cz = (int)floorf(fminf(fmaxf(f[2], -2.f), (float)g.Z + 1.f));
  for (int dx = -1; dx <= 2; ++dx)
    for (int dy = -1; dy <= 2; ++dy)
      for (int dz = -1; dz <= 2; ++dz) {
        const int px = cx + dx, py = cy + dy, pz = cz + dz;
        if (px < 0 || py < 0 || pz < 0 || px >= g.X || py >= g.Y || pz >= g.Z) continue;
        const double ddx = (double)f[0] - px, ddy = (double)f[1] - py, ddz = (double)f[2] - pz;
        const float dist = (float)(ddx * ddx + ddy * ddy + ddz * ddz);
        const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)v;
        atomicMin(keys + (size_t)b * N + ((size_t)px * g.Y + py) * g.Z + pz, key);
      }
}

__global__ void iwarp_gather_kernel(Geom g, const float* __restrict__ frames, int frames_are_batch,
                                    const int* __restrict__ frame_ids, int B, const float* __restrict__ beta,
                                    const unsigned long long* __restrict__ keys, float* __restrict__ out) {
  const size_t N = (size_t)g.X * g.Y * g.Z;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * B) return;
  const int b = (int)(idx / N);
  const size_t v = idx - (size_t)b * N;
  const int t = frame_ids[b];
  const float* frame = frames + (size_t)(frames_are_batch ? b : t) * N;
  unsigned long long key = keys[idx];
  if (key != ~0ull) {
    out[idx] = frame[(unsigned)(key & 0xffffffffu)];
    return;
  }
  // no scattered point within reach: exact nearest by brute force (rare; large deformations only)
  const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((size_t)g.Z * g.Y));
  const int sz[3] = {g.X, g.Y, g.Z};
  double best = 1e300;
  size_t besti = 0;
  for (size_t s = 0; s < N; ++s) {
    const int sz_ = (int)(s % g.Z), sy = (int)((s / g.Z) % g.Y), sx = (int)(s / ((size_t)g.Z * g.Y));
    const float xf = (float)sx, yf = (float)sy, zf = (float)sz_;
    const float phi[kBasis] = {1.f, xf, yf, zf, xf * xf, yf * yf, zf * zf, xf * yf, xf * zf, yf * zf};
    double d2 = 0.0;
    const int pt[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float q = 0.f;
#pragma unroll
      for (int a = 0; a < kBasis; ++a) q = __fadd_rn(q, __fmul_rn(phi[a], beta[((size_t)a * 3 + d) * g.T + t]));
      const float sm1 = (float)(sz[d] - 1);
      float u = sm1 == 0.f ? 0.f : __fsub_rn(__fdiv_rn(__fmul_rn(2.f, q), sm1), 1.f);
      const double fd = (double)__fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)sz[d]) - pt[d];
      d2 += fd * fd;
    }
    if (d2 < best) {
      best = d2;
      besti = s;
    }
  }
  out[idx] = frame[besti];
}

}  // namespace dnmf

using namespace dnmf;

static int mu_alloc(dnmf_ctx* c) {
  if (!c->d_G) {
    CU(cudaMalloc((void**)&c->d_G, (size_t)c->T * c->K * c->K * sizeof(double)));
    CU(cudaMemset(c->d_G, 0, (size_t)c->T * c->K * c->K * sizeof(double)));
  }
  if (!c->d_b) {
    CU(cudaMalloc((void**)&c->d_b, (size_t)c->T * c->K * sizeof(double)));
    CU(cudaMemset(c->d_b, 0, (size_t)c->T * c->K * sizeof(double)));
  }
  return 0;
}

template <int R, int BS>
static int launch_mu(const MuParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = mu_stats_kernel<R, BS>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kMuThreads, smem, st>>>(p);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_mu_stats(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                             const float* beta_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev) return fail("dnmf_mu_stats: NULL argument");
  if (!c->have_footprints) return fail("dnmf_mu_stats: call dnmf_set_footprints first");
  if (!frames_dev && !c->d_video) return fail("dnmf_mu_stats: no resident video and frames_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (mu_alloc(c)) return 1;
  c->gc_valid = false;
  // Fast path: the fused kernel's tiles, lists and staged slices (fit_tile_kernel<MODE=3>).  A tile whose list
  // is longer than the staged capacity under the current deformation sets the overflow flag and everything is
  // redone by the panel kernel below.
  while (fused_stats_available(c) && c->mu_force_panel == 0 && c->mu_fused_off == 0) {  // at most one pass
    FitParams p;
    if (fill_fit_params(c, p, frames_dev, frame_ids_dev, B, beta_dev, nullptr)) return 1;
    p.muG = c->d_G;
    p.mub = c->d_b;
    p.mu_overflow = c->d_tmp_max;
    // Every slot must be staged here (there is no global-table tail as in the fit): capacity for the longest
    // identity-deformation list + 2 when that fits the shared-memory budget of ~2 CTAs per SM.
    size_t smem = c->fit_smem;
    {
      int cap = (std::min(c->K + 1, std::max(c->lmax_identity, c->mu_fused_need) + 2) + 1) & ~1;
      if ((cap & 3) == 0) cap += 2;
      if (cap > c->cap) {
        const int wsum = c->wmax[0] + c->wmax[1] + c->wmax[2];
        const size_t need = fit_smem_layout(c->nwx * c->nwy, c->tx, c->ty, c->tz, cap, wsum, c->K, c->wmax[0],
                                            c->cand_cap, c->y_pitch).bytes;
        if (need <= std::min<size_t>((size_t)c->max_smem_optin, (size_t)113 * 1024)) {
          p.cap = cap;
          smem = need;
        } else if (c->mu_fused_need > 0) {
          c->mu_fused_off = 1;  // known to overflow and no room to grow
          break;
        }
      }
    }
    CU(cudaMemsetAsync(c->d_tmp_max, 0, sizeof(int), st));
    mu_zero_kernel<<<B, 256, 0, st>>>(c->d_G, c->d_b, frame_ids_dev, c->K);
    CU(cudaGetLastError());
    if (dispatch_stats(c, p, B, smem, st)) return 1;
    int over = 0;
    CU(cudaMemcpyAsync(&over, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->counters[6] += 1;
    if (over == 0) {
      c->mu_last_path = 1;
      return 0;
    }
    // Remembered until the next dnmf_set_footprints.  A list longer than the capacity: stage that many slots
    // next time.  Otherwise a window wider than the staged slices raised the flag: not a matter of capacity.
    if (over > p.cap) c->mu_fused_need = over;
    else c->mu_fused_off = 1;
    break;
  }
  c->mu_last_path = 0;
  // geometry of the statistics kernel (16 x 8 x tz tiles) and its longest identity-deformation list
  dnmf_ctx g = *c;  // shallow copy used only for geometry helpers
  g.tx = kMuTX;
  g.ty = kMuTY;
  g.ntx = (c->X + kMuTX - 1) / kMuTX;
  g.nty = (c->Y + kMuTY - 1) / kMuTY;
  const int nt = g.ntx * g.nty * g.ntz;
  if (c->mu_capM == 0) {
    if (ensure(&c->d_tmp_counts, &c->tmp_counts_cap, (size_t)nt)) return 1;
    if (c->d_tmp_offsets) cudaFree(c->d_tmp_offsets);
    c->d_tmp_offsets = nullptr;
    CU(cudaMalloc((void**)&c->d_tmp_offsets, ((size_t)nt + 1) * sizeof(long long)));
    int zero = 0;
    if (ensure(&c->d_ids, &c->ids_cap, (size_t)1)) return 1;
    CU(cudaMemcpyAsync(c->d_ids, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    g.d_tmp_counts = c->d_tmp_counts;
    if (run_bin_count(&g, c->d_identity_beta, 1, c->d_ids, 1, c->d_tmp_counts, nullptr, st)) return 1;
    scan_counts_kernel<<<1, 1024, 0, st>>>(c->d_tmp_counts, nt, c->d_tmp_offsets, c->d_tmp_max);
    CU(cudaGetLastError());
    int lmax = 0;
    CU(cudaMemcpyAsync(&lmax, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int capM = std::min(c->K, lmax + lmax / 2 + 8) + 1;
    c->mu_capM = (capM + 7) & ~7;
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int capM = c->mu_capM;
    // 8x8 register blocks from 32 rows up (enough blocks to occupy the CTA), 4x4 below
    const int bs = (capM >= 32 && c->mu_block4 == 0) ? 8 : 4;
    const size_t smem = mu_smem_bytes(capM, c->tz, c->K, bs);
    if (smem > (size_t)c->max_smem_optin)
      return fail("dnmf_mu_stats: neuron lists too long for the shared-memory A panel (K_eff too large)");
    MuParams p;
    p.frames = frames_dev ? frames_dev : c->d_video;
    p.frames_are_batch = frames_dev ? 1 : 0;
    p.frame_ids = frame_ids_dev;
    p.beta = beta_dev;
    p.tab0 = c->d_tab[0];
    p.tab1 = c->d_tab[1];
    p.tab2 = c->d_tab[2];
    p.rng = c->d_rng;
    p.G = c->d_G;
    p.bvec = c->d_b;
    p.overflow = c->d_tmp_max;
    p.X = c->X;
    p.Y = c->Y;
    p.Z = c->Z;
    p.K = c->K;
    p.T = c->T;
    p.tz = c->tz;
    p.ntx = g.ntx;
    p.nty = g.nty;
    p.ntz = g.ntz;
    p.capM = capM;
    p.full_depth = (c->tz == c->Z) ? 1 : 0;
    CU(cudaMemsetAsync(c->d_tmp_max, 0, sizeof(int), st));
    mu_zero_kernel<<<B, 256, 0, st>>>(c->d_G, c->d_b, frame_ids_dev, c->K);
    CU(cudaGetLastError());
    const int mb = (capM + bs - 1) / bs;
    const int nblk = mb * (mb + 1) / 2;
    const int need = (nblk + kMuThreads - 1) / kMuThreads;
    int rc;
    if (bs == 8) {
      if (need <= 1) rc = launch_mu<1, 8>(p, B * nt, smem, st);
      else rc = launch_mu<2, 8>(p, B * nt, smem, st);
    } else if (need <= 1) rc = launch_mu<1, 4>(p, B * nt, smem, st);
    else if (need <= 2) rc = launch_mu<2, 4>(p, B * nt, smem, st);
    else rc = launch_mu<4, 4>(p, B * nt, smem, st);
    if (rc) return rc;
    int over = 0;
    CU(cudaMemcpyAsync(&over, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c->counters[6] += 1;
    if (over == 0) return 0;
    // a tile's list outgrew the panel under the current deformation: grow once and redo
    c->mu_capM = (std::min(c->K + 1, over + over / 4 + 4) + 7) & ~7;
  }
  return fail("dnmf_mu_stats: neuron list longer than the A panel after regrowth");
}

extern "C" int dnmf_mu_path(dnmf_ctx* c, int force_panel, int* last_path_out) {
  if (!c) return fail("dnmf_mu_path: NULL context");
  if (force_panel >= 0) {
    c->mu_force_panel = (force_panel & 1) != 0;
    c->mu_dense_sweeps = (force_panel & 2) != 0;
    c->mu_sweep_per_launch = (force_panel & 4) != 0;
  }
  if (last_path_out) *last_path_out = c->mu_last_path | (c->mu_last_sparse << 1);
  return 0;
}

extern "C" int dnmf_get_mu_stats(dnmf_ctx* c, int t, double* G_host, double* b_host) {
  if (!c || t < 0 || t >= c->T) return fail("dnmf_get_mu_stats: bad argument");
  if (!c->d_G) return fail("dnmf_get_mu_stats: call dnmf_mu_stats first");
  CU(cudaSetDevice(c->device));
  if (G_host) CU(cudaMemcpy(G_host, c->d_G + (size_t)t * c->K * c->K, (size_t)c->K * c->K * sizeof(double), cudaMemcpyDeviceToHost));
  if (b_host) CU(cudaMemcpy(b_host, c->d_b + (size_t)t * c->K, (size_t)c->K * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// Static neighbour lists from the integer node ranges of the truncated footprints: A_k(ix) != 0 needs
// lo_k - 1 < ix_d < hi_k + 1 on every axis, so G[k][l] != 0 needs the open intervals of k and l to meet.
static int mu_build_neighbours(dnmf_ctx* c) {
  c->mu_nbr_built = true;
  c->mu_nbrw = 0;
  const int K = c->K;
  std::vector<int> rng((size_t)K * 6);
  CU(cudaMemcpy(rng.data(), c->d_rng, rng.size() * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<std::vector<int>> lists((size_t)K);
  size_t longest = 1;
  for (int k = 0; k < K; ++k) {
    const int* a = &rng[(size_t)k * 6];
    for (int l = 0; l < K; ++l) {
      const int* b = &rng[(size_t)l * 6];
      bool meet = true;
      for (int d = 0; d < 3; ++d)
        meet = meet && a[2 * d] <= a[2 * d + 1] && b[2 * d] <= b[2 * d + 1] &&  // non-empty ranges
               a[2 * d] - 1 <= b[2 * d + 1] + 1 && b[2 * d] - 1 <= a[2 * d + 1] + 1;
      if (meet || l == k) lists[(size_t)k].push_back(l);
    }
    longest = std::max(longest, lists[(size_t)k].size());
    if (longest * 2 > (size_t)K) return 0;  // dense overlap: the dense sweep reads less
  }
  const int W = std::max(4, (int)((longest + 3) & ~(size_t)3));  // row length of the compacted statistics
  std::vector<int> flat((size_t)K * W, -1);
  for (int k = 0; k < K; ++k) std::copy(lists[(size_t)k].begin(), lists[(size_t)k].end(), flat.begin() + (size_t)k * W);
  if (c->d_mu_nbr) cudaFree(c->d_mu_nbr);
  c->d_mu_nbr = nullptr;
  CU(cudaMalloc((void**)&c->d_mu_nbr, flat.size() * sizeof(int)));
  CU(cudaMemcpy(c->d_mu_nbr, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
  c->mu_nbrw = W;
  return 0;
}

extern "C" int dnmf_mu_begin(dnmf_ctx* c, const float* C_dev, void* stream) {
  if (!c || !C_dev) return fail("dnmf_mu_begin: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  // compact the statistics to the static neighbour lists (once per set of statistics)
  c->mu_last_sparse = 0;
  if (c->d_G && c->have_footprints && c->mu_dense_sweeps == 0) {
    if (!c->mu_nbr_built && mu_build_neighbours(c)) return 1;
    if (c->mu_nbrw > 0) {
      if (!c->gc_valid) {
        if (ensure(&c->d_Gc, &c->gc_cap, (size_t)c->T * c->K * c->mu_nbrw)) return 1;
        CU(cudaMemsetAsync(c->d_tmp_max, 0, sizeof(int), st));
        const long long threads = (long long)c->T * c->K * 32;
        mu_compact_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(c->d_G, c->d_mu_nbr, c->mu_nbrw, c->T, c->K,
                                                                           c->d_Gc, c->d_tmp_max);
        CU(cudaGetLastError());
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (bad) c->mu_nbrw = 0;  // a non-zero outside the lists: never expected; keep the dense sweeps
        else c->gc_valid = true;
      }
      c->mu_last_sparse = c->mu_nbrw > 0 ? 1 : 0;
    }
  }
  const size_t n = (size_t)c->K * c->T;
  if (!c->d_Cd[0]) {
    CU(cudaMalloc((void**)&c->d_Cd[0], n * sizeof(double)));
    CU(cudaMalloc((void**)&c->d_Cd[1], n * sizeof(double)));
  }
  c->cd_cur = 0;
  mu_load_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C_dev, c->d_Cd[0], c->K, c->T);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_mu_sweep(dnmf_ctx* c, double gamma, int use_gamma, const double* halo_prev_dev,
                             const double* halo_next_dev, void* stream) {
  if (!c) return fail("dnmf_mu_sweep: ctx is NULL");
  if (!c->d_G || !c->d_b) return fail("dnmf_mu_sweep: call dnmf_mu_stats first");
  if (!c->d_Cd[0]) return fail("dnmf_mu_sweep: call dnmf_mu_begin first");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (c->mu_last_sparse && c->gc_valid && c->mu_nbrw > 0) {
    const int W = c->mu_nbrw;
    int Ws = 8;  // lanes per row: the power of two covering W, 8..32 (lanes past W idle, longer rows loop)
    while (Ws < 32 && Ws < W) Ws *= 2;
    const int rows_per_block = 256 / Ws;
    const dim3 grid((unsigned)((c->T + kMuTB - 1) / kMuTB), (unsigned)((c->K + rows_per_block - 1) / rows_per_block));
    mu_sweep_sparse_kernel<<<grid, 256, 0, st>>>(
        c->d_Gc, c->d_mu_nbr, W, Ws, c->d_b, c->d_Cd[c->cd_cur], c->d_Cd[c->cd_cur ^ 1], c->T, c->K, gamma, use_gamma,
        halo_prev_dev, halo_next_dev);
  } else {
    const dim3 grid((unsigned)c->T, (unsigned)((c->K + 7) / 8));
    mu_sweep_kernel<<<grid, 256, 0, st>>>(
        c->d_G, c->d_b, c->d_Cd[c->cd_cur], c->d_Cd[c->cd_cur ^ 1], c->T, c->K, gamma, use_gamma, halo_prev_dev,
        halo_next_dev);
  }
  CU(cudaGetLastError());
  c->cd_cur ^= 1;
  c->counters[7] += 1;
  return 0;
}

extern "C" int dnmf_mu_boundary(dnmf_ctx* c, double* first_dev, double* last_dev, void* stream) {
  if (!c) return fail("dnmf_mu_boundary: ctx is NULL");
  if (!c->d_Cd[0]) return fail("dnmf_mu_boundary: call dnmf_mu_begin first");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  mu_boundary_kernel<<<(c->K + 127) / 128, 128, 0, st>>>(c->d_Cd[c->cd_cur], c->T, c->K, first_dev, last_dev);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_mu_end(dnmf_ctx* c, float* C_dev, void* stream) {
  if (!c || !C_dev) return fail("dnmf_mu_end: NULL argument");
  if (!c->d_Cd[0]) return fail("dnmf_mu_end: call dnmf_mu_begin first");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  const size_t n = (size_t)c->K * c->T;
  mu_store_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->d_Cd[c->cd_cur], C_dev, c->K, c->T);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int dnmf_mu_sweeps(dnmf_ctx* c, float* C_dev, double gamma, int use_gamma, int iters, void* stream) {
  if (dnmf_mu_begin(c, C_dev, stream)) return 1;
  // no temporal coupling (gamma None or exactly 0) and compacted statistics: all sweeps of a frame in one CTA
  const size_t local_smem = (size_t)c->K * c->mu_nbrw * 12 + (size_t)c->K * 24;
  if ((use_gamma == 0 || gamma == 0.0) && iters > 0 && c->mu_last_sparse && c->gc_valid && c->mu_nbrw > 0 &&
      c->mu_sweep_per_launch == 0 && local_smem <= (size_t)c->max_smem_optin) {
    cudaStream_t st = (cudaStream_t)stream;
    if (local_smem > 48 * 1024)
      CU(cudaFuncSetAttribute(mu_sweeps_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)local_smem));
    const int threads = std::min(1024, (c->K + 31) & ~31);
    mu_sweeps_local_kernel<<<c->T, threads, local_smem, st>>>(c->d_Gc, c->d_mu_nbr, c->mu_nbrw, c->d_b,
                                                             c->d_Cd[c->cd_cur], c->d_Cd[c->cd_cur ^ 1], c->K, iters);
    CU(cudaGetLastError());
    c->cd_cur ^= 1;
    c->counters[7] += 1;
    return dnmf_mu_end(c, C_dev, stream);
  }
  for (int i = 0; i < iters; ++i)
    if (dnmf_mu_sweep(c, gamma, use_gamma, nullptr, nullptr, stream)) return 1;
  return dnmf_mu_end(c, C_dev, stream);
}

extern "C" int dnmf_iwarp(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                          const float* beta_dev, float* out_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev || !out_dev) return fail("dnmf_iwarp: NULL argument");
  if (!frames_dev && !c->d_video) return fail("dnmf_iwarp: no resident video and frames_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  const size_t total = c->N * (size_t)B;
  if (ensure(&c->d_keys, &c->keys_cap, total)) return 1;
  CU(cudaMemsetAsync(c->d_keys, 0xff, total * sizeof(unsigned long long), st));
  Geom g = geom_of(c);
  iwarp_vote_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, frame_ids_dev, B, beta_dev, c->d_keys);
  CU(cudaGetLastError());
  iwarp_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      g, frames_dev ? frames_dev : c->d_video, frames_dev ? 1 : 0, frame_ids_dev, B, beta_dev, c->d_keys, out_dev);
  CU(cudaGetLastError());
  return 0;
}

// ---- FP32 peak microbenchmark (roofline denominator for the FP32-bound fused kernel) ---------------
namespace dnmf {
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
        x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b);
      x1 = fmaf(x1, a, b);
      x2 = fmaf(x2, a, b);
      x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b);
      x5 = fmaf(x5, a, b);
      x6 = fmaf(x6, a, b);
      x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace dnmf

extern "C" int dnmf_measure_fp32_peak(int device, int repeats, double* tflops_out) {
  if (!tflops_out) return fail("dnmf_measure_fp32_peak: NULL argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float* d = nullptr;
  CU(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < repeats + 2; ++r) {
    CU(cudaEventRecord(e0, 0));
    fp32_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    if (r >= 2) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return 0;
}
