// Trace update (kernel 3b) -- included at the end of dnmf_kernels.cu (same translation unit).
//
// mu_stats_kernel: per (frame, 16x8xtz tile) CTA of 128 threads.  Phase 1: every thread evaluates
// the closed-form footprint value A_j(p) of each listed neuron at its voxel and parks it in shared
// memory next to the voxel's Y value (pseudo-neuron).  Phase 2: the CTA computes the small SYRK
// [A|Y]^T [A|Y] of those 128 voxels with 4x4 register blocks and split-K over the voxel index.
// Accumulators live in registers for the whole tile and are stored once into the tile-frame's partial block
// (StatsPartials); stats_reduce_kernel sums the blocks of a frame in ascending tile order in fp64 into the
// dense per-frame G_t[K][K], b_t[K].   Reference: Demix/dNMF.py:141-142 (fp64 einsum).
// This SIMT form serves lists longer than the 127 rows of the tensor-core panel kernel (dnmf_gram_tc.cu).
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
  StatsPartials out;
  int b_base;  // batch position of this launch's first frame in the partial buffers
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
  if (L + 1 > p.capM || L > p.out.capL) {  // loud failure: the host checks this flag after the launch
    if (tid == 0) atomicMax(p.overflow, L + 1);
    return;
  }
  const size_t tf = (size_t)(p.b_base + b) * nt + tile;
  if (tid == 0) p.out.count[tf] = L;
  if (L == 0) return;
  __syncthreads();  // sList complete
  for (int pos = tid; pos < L; pos += kMuThreads) {
    const int k = sList[pos];
    p.out.ids[tf * p.out.capL + pos] = (unsigned short)k;
    p.out.slot_of[((size_t)(p.b_base + b) * p.K + k) * nt + tile] = (unsigned short)pos;
  }
  float* pblock = p.out.vals + tf * (size_t)p.out.capL * p.out.ld;
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

    // flush: reduce the split-K lanes, then plain stores into the tile-frame's partial block (every entry of the
    // block is produced by exactly one register block; diagonal blocks hold both triangles)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int blk_i = base + r * groups + grp;
#pragma unroll
      for (int i = 0; i < BS * BS; ++i) {
        float v = acc[r][i];
        for (int o = ks >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][i] = v;
      }
      if (blk_i < nblk && ksl == 0) {
        const int code = sBlk[blk_i];
        const int bi = code >> 8, bj = code & 255;
#pragma unroll
        for (int i = 0; i < BS; ++i) {
          const int jr = bi * BS + i;
          if (jr >= L) continue;  // Y pseudo-row or padding
#pragma unroll
          for (int c = 0; c < BS; ++c) {
            const int jc = bj * BS + c;
            const float v = acc[r][i * BS + c];
            if (jc < L) {
              pblock[(size_t)jr * p.out.ld + jc] = v;
              if (bi != bj) pblock[(size_t)jc * p.out.ld + jr] = v;
            } else if (jc == L) {
              pblock[(size_t)jr * p.out.ld + p.out.capL] = v;
            }
          }
        }
      }
    }
  }
}

// Second stage of the trace statistics, one warp per (frame of the chunk, neuron k) = one row of G_t: the warp
// walks the tiles in ascending order, finds the ones that list k (slot_of), and adds that tile's partial row to
// its row accumulators in shared memory (fp64), then writes the whole row G_t[k][:] and b_t[k] once.  A row has
// one owner and tiles are visited in a fixed order: bitwise reproducible, no atomics, and exactly symmetric --
// every first stage stores bitwise symmetric blocks (the fused tiles and the SIMT panel mirror their off-diagonal
// register blocks, the tensor-core panel stores both triangles from the upper one), so a row reads its own row of
// each block: contiguous (reading element (min slot, max slot) instead cost 8x the sectors below the diagonal).
__global__ void stats_reduce_kernel(StatsPartials sp, const int* __restrict__ frame_ids, int nt, int K,
                                    double* __restrict__ G, double* __restrict__ bvec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int k = blockIdx.x * wpb + warp, b = blockIdx.y;
  if (k >= K) return;
  double* row = reinterpret_cast<double*>(smem_raw) + (size_t)warp * K;
  for (int l = lane; l < K; l += 32) row[l] = 0.0;
  double bsum = 0.0;
  __syncwarp();
  const unsigned short* slots = sp.slot_of + ((size_t)b * K + k) * nt;
  constexpr int R = 8;  // tiles per lane and round: the slot and list-length loads of a round are all in flight together
  for (int tile0 = 0; tile0 < nt; tile0 += 32 * R) {
    unsigned s[R];
    int cnt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int tl = tile0 + r * 32 + lane;
      s[r] = tl < nt ? slots[tl] : 0xffffu;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int tl = tile0 + r * 32 + lane;
      cnt[r] = s[r] != 0xffffu ? sp.count[(size_t)b * nt + tl] : 0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      unsigned mask = __ballot_sync(0xffffffffu, s[r] != 0xffffu);
      while (mask) {  // ascending tile order: the fixed summation order of every entry of the row
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int j = (int)__shfl_sync(0xffffffffu, s[r], src);
        const int L = __shfl_sync(0xffffffffu, cnt[r], src);
        const size_t tf = (size_t)b * nt + tile0 + r * 32 + src;
        const float* blk = sp.vals + tf * (size_t)sp.capL * sp.ld + (size_t)j * sp.ld;  // this neuron's row of the block
        const unsigned short* ids = sp.ids + tf * sp.capL;
        const float bj = lane == 0 ? blk[sp.capL] : 0.f;
        for (int i0 = 0; i0 < L; i0 += 128) {  // four entries per lane in flight: one memory round trip per 128 entries
          float v[4];
          int id[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            id[u] = 0;
            v[u] = 0.f;
            if (i < L) {
              id[u] = ids[i];
              DNMF_DASSERT(id[u] < K && j < L && L <= sp.capL);
              v[u] = blk[i];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * 32 + lane < L) row[id[u]] += (double)v[u];
        }
        if (lane == 0) bsum += (double)bj;
        __syncwarp();  // (four listing tiles per round with all their loads in flight measured 20 % slower: not latency-bound)
      }
    }
  }
  const int t = frame_ids[b];
  DNMF_DASSERT(t >= 0);
  double* g = G + ((size_t)t * K + k) * K;
  for (int l = lane; l < K; l += 32) g[l] = row[l];
  if (lane == 0) bvec[(size_t)t * K + k] = bsum;
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
  // every non-zero of the row must be one of the listed entries: the list holds distinct neurons, so it is enough that
  // the row has as many non-zeros as its listed entries have (no search per non-zero: W ~ 300 of them at K = 1000)
  int nnz_row = 0, nnz_list = 0;
  for (int l = lane; l < K; l += 32) nnz_row += g[l] != 0.0 ? 1 : 0;
  for (int s_ = lane; s_ < W; s_ += 32) {
    const int l = nb[s_];
    const double v = l >= 0 ? g[l] : 0.0;
    Gc[(size_t)row * W + s_] = v;
    nnz_list += v != 0.0 ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nnz_row += __shfl_xor_sync(0xffffffffu, nnz_row, o);
    nnz_list += __shfl_xor_sync(0xffffffffu, nnz_list, o);
  }
  if (lane == 0 && nnz_row != nnz_list) atomicOr(violation, 1);
}

// Ws lanes per neuron k (the power of two covering the row length W, at most 32); one thread walks kMuTB
// consecutive frames of its (k, list slot): the neighbour id is loaded once and the kMuTB (G, C) pairs are
// independent loads in flight (one row per warp and frame was latency-bound: nbr -> C gather -> reduce -> store).
// grid = (ceil(T / kMuTB), ceil(K / rows per block)).
constexpr int kMuTB = 4;  // 8 frames per thread measured slower (cfg4 50 sweeps 3.78 -> 4.32 ms per 100 frames)
template <int kMuSB>
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
    // kMuSB list slots per round: their neighbour ids are requested together, then the kMuSB x kMuTB (G, C) pairs --
    // two dependent memory round trips per round instead of per slot (long rows: W ~ 300 at K = 1000: 50 sweeps over
    // 100 frames 3.78 -> 2.49 ms; rows of one round, W <= 32, run the kMuSB = 1 instantiation)
    for (int s0 = sub; s0 < W; s0 += kMuSB * Ws) {
      int l[kMuSB];
#pragma unroll
      for (int v = 0; v < kMuSB; ++v) {
        const int s_ = s0 + v * Ws;
        l[v] = s_ < W ? max(nbr[(size_t)k * W + s_], 0) : 0;  // padding slots hold G = 0
        DNMF_DASSERT(l[v] < K);
      }
      double g[kMuSB][kMuTB], cc[kMuSB][kMuTB];
#pragma unroll
      for (int v = 0; v < kMuSB; ++v) {
        const int s_ = s0 + v * Ws;
#pragma unroll
        for (int u = 0; u < kMuTB; ++u) {
          const int t = min(t0 + u, T - 1);
          g[v][u] = s_ < W ? Gc[((size_t)t * K + k) * W + s_] : 0.0;
          cc[v][u] = Cin[(size_t)t * K + l[v]];
        }
      }
#pragma unroll
      for (int v = 0; v < kMuSB; ++v)   // ascending slot order: the same sum as one slot per round
#pragma unroll
        for (int u = 0; u < kMuTB; ++u) dot[u] = fma(g[v][u], cc[v][u], dot[u]);
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
// reads the nearest scattered point at every integer voxel (scipy NearestNDInterpolator: exact nearest
// neighbour on float64 coordinates).  Here every scattered point votes for the integer voxels within
// 2 of it on every axis: pass 0 keeps the smallest squared distance (fp64, compared as its bit pattern),
// pass 1 the lowest source index among the points at exactly that distance.  A point that did not vote for
// a voxel is at least 2 away from it on some axis, so a winning distance below 4 is the true nearest
// neighbour; voxels without a vote or with a winner at distance >= 4 take an exact brute-force search
// (ties: lowest source index, like the votes).
__device__ __forceinline__ void iwarp_point(const Geom& g, const float* __restrict__ beta, int t, int x, int y, int z,
                                            float (&f)[3]) {
  const float xf = (float)x, yf = (float)y, zf = (float)z;
  const float phi[kBasis] = {1.f, xf, yf, zf, xf * xf, yf * yf, zf * zf, xf * yf, xf * zf, yf * zf};
  const int sz[3] = {g.X, g.Y, g.Z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float q = 0.f;
#pragma unroll
    for (int a = 0; a < kBasis; ++a) q = __fadd_rn(q, __fmul_rn(phi[a], beta[((size_t)a * 3 + d) * g.T + t]));
    const float sm1 = (float)(sz[d] - 1);
    const float u = sm1 == 0.f ? 0.f : __fsub_rn(__fdiv_rn(__fmul_rn(2.f, q), sm1), 1.f);
    f[d] = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)sz[d]);
  }
}

template <int PASS>
__global__ void iwarp_vote_kernel(Geom g, const int* __restrict__ frame_ids, int B, const float* __restrict__ beta,
                                  unsigned long long* __restrict__ best, unsigned* __restrict__ winner) {
  const size_t N = (size_t)g.X * g.Y * g.Z;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * B) return;
  const int b = (int)(idx / N);
  const size_t v = idx - (size_t)b * N;
  const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((size_t)g.Z * g.Y));
  float f[3];
  iwarp_point(g, beta, frame_ids[b], x, y, z, f);
  const int cx = (int)floorf(fminf(fmaxf(f[0], -2.f), (float)g.X + 1.f));
  const int cy = (int)floorf(fminf(fmaxf(f[1], -2.f), (float)g.Y + 1.f));
  const int cz = (int)floorf(fminf(fmaxf(f[2], -2.f), (float)g.Z + 1.f));
  for (int dx = -1; dx <= 2; ++dx)
    for (int dy = -1; dy <= 2; ++dy)
      for (int dz = -1; dz <= 2; ++dz) {
        const int px = cx + dx, py = cy + dy, pz = cz + dz;
        if (px < 0 || py < 0 || pz < 0 || px >= g.X || py >= g.Y || pz >= g.Z) continue;
        const double ddx = (double)f[0] - px, ddy = (double)f[1] - py, ddz = (double)f[2] - pz;
        const unsigned long long key = (unsigned long long)__double_as_longlong(ddx * ddx + ddy * ddy + ddz * ddz);
        const size_t o = (size_t)b * N + ((size_t)px * g.Y + py) * g.Z + pz;
        if (PASS == 0) atomicMin(best + o, key);
        else if (best[o] == key) atomicMin(winner + o, (unsigned)v);
      }
}

__global__ void iwarp_gather_kernel(Geom g, const float* __restrict__ frames, int frames_are_batch,
                                    const int* __restrict__ frame_ids, int B, const float* __restrict__ beta,
                                    const unsigned long long* __restrict__ best, const unsigned* __restrict__ winner,
                                    float* __restrict__ out) {
  const size_t N = (size_t)g.X * g.Y * g.Z;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * B) return;
  const int b = (int)(idx / N);
  const size_t v = idx - (size_t)b * N;
  const int t = frame_ids[b];
  const float* frame = frames + (size_t)(frames_are_batch ? b : t) * N;
  const unsigned long long key = best[idx];
  if (key != ~0ull && __longlong_as_double((long long)key) < 4.0) {
    out[idx] = frame[winner[idx]];
    return;
  }
  // no scattered point within reach: exact nearest by brute force (rare; large deformations only)
  const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((size_t)g.Z * g.Y));
  double bestd = 1e300;
  size_t besti = 0;
  for (size_t s = 0; s < N; ++s) {
    const int sz_ = (int)(s % g.Z), sy = (int)((s / g.Z) % g.Y), sx = (int)(s / ((size_t)g.Z * g.Y));
    float f[3];
    iwarp_point(g, beta, t, sx, sy, sz_, f);
    const double ddx = (double)f[0] - x, ddy = (double)f[1] - y, ddz = (double)f[2] - z;
    const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
    if (d2 < bestd) {
      bestd = d2;
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

// ---- trace statistics: first stage (one of three kernels) + row-owner second stage ------------------------
// Scratch for the partial blocks of `frames` frames with nt tiles each and blocks of capL rows.
static size_t stats_bytes_per_frame(int nt, int capL, int K) {
  const size_t ld = (size_t)capL + 4;
  return (size_t)nt * ((size_t)capL * ld * 4 + (size_t)capL * 2 + 4) + (size_t)K * nt * 2;
}

static int stats_scratch(dnmf_ctx* c, int frames, int nt, int capL, StatsPartials& sp, cudaStream_t st) {
  sp.capL = capL;
  sp.ld = capL + 4;
  if (ensure(&c->d_pb_vals, &c->pb_vals_cap, (size_t)frames * nt * capL * sp.ld)) return 1;
  if (ensure(&c->d_pb_ids, &c->pb_ids_cap, (size_t)frames * nt * capL)) return 1;
  if (ensure(&c->d_pb_count, &c->pb_count_cap, (size_t)frames * nt)) return 1;
  if (ensure(&c->d_pb_slot, &c->pb_slot_cap, (size_t)frames * c->K * nt)) return 1;
  sp.vals = c->d_pb_vals;
  sp.ids = c->d_pb_ids;
  sp.count = c->d_pb_count;
  sp.slot_of = c->d_pb_slot;
  CU(cudaMemsetAsync(sp.slot_of, 0xff, (size_t)frames * c->K * nt * sizeof(unsigned short), st));
  return 0;
}

static int stats_reduce(dnmf_ctx* c, const StatsPartials& sp, const int32_t* ids_dev, int frames, int nt,
                        cudaStream_t st) {
  const size_t row_bytes = (size_t)c->K * sizeof(double);
  int wpb = (int)std::min<size_t>(8, ((size_t)c->max_smem_optin - 1024) / row_bytes);
  if (wpb < 1) return fail("dnmf_mu_stats: K too large for the row accumulators of the second stage");
  const size_t smem = (size_t)wpb * row_bytes;
  static size_t configured[64] = {0};
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > configured[dev])) {
    CU(cudaFuncSetAttribute(stats_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  for (int b0 = 0; b0 < frames; b0 += 65535) {
    const int nb = std::min(65535, frames - b0);
    StatsPartials q = sp;
    q.vals += (size_t)b0 * nt * sp.capL * sp.ld;
    q.ids += (size_t)b0 * nt * sp.capL;
    q.count += (size_t)b0 * nt;
    q.slot_of += (size_t)b0 * c->K * nt;
    stats_reduce_kernel<<<dim3((unsigned)((c->K + wpb - 1) / wpb), (unsigned)nb), 32 * wpb, smem, st>>>(
        q, ids_dev + b0, nt, c->K, c->d_G, c->d_b);
    CU(cudaGetLastError());
  }
  return 0;
}

// One pass over the batch with one first-stage kernel.  which: 0 = fused tiles (fit_tile_kernel<MODE 3>),
// 1 = tensor-core panel, 2 = SIMT panel.  *over receives the kernel's overflow value (0 = the pass is complete;
// otherwise nothing of this pass may be used and the caller moves on to the next kernel).
static int stats_pass(dnmf_ctx* c, int which, const float* frames_dev, const int32_t* ids_dev, int B,
                      const float* beta_dev, cudaStream_t st, int* over) {
  *over = 0;
  const size_t budget = (size_t)3 << 29;  // 1.5 GB of partial blocks in flight
  int nt, capL;
  size_t smem = 0;
  MuParams mp;
  GramTcParams tp;
  int mu_bs = 4, mu_need = 1;
  int fused_cap = c->cap;
  if (which == 0) {
    // Every slot must be staged here (there is no global-table tail as in the fit): capacity for the longest
    // identity-deformation list + 2 when that fits the shared-memory budget of ~2 CTAs per SM.
    smem = c->fit_smem;
    int cap = (std::min(c->K + 1, std::max(c->lmax_identity, c->mu_fused_need) + 2) + 1) & ~1;
    if ((cap & 3) == 0) cap += 2;
    if (cap > c->cap) {
      const int wsum = c->wmax[0] + c->wmax[1] + c->wmax[2];
      const size_t need = fit_smem_layout(c->nwx * c->nwy * c->nwz, c->tx, c->ty, c->tz, cap, wsum, c->K, c->wmax[0],
                                          c->cand_cap, c->y_pitch).bytes;
      if (need <= std::min<size_t>((size_t)c->max_smem_optin, (size_t)113 * 1024)) {
        fused_cap = cap;
        smem = need;
      } else if (c->mu_fused_need > 0) {
        c->mu_fused_off = 1;  // known to overflow and no room to grow
        *over = c->mu_fused_need;
        return 0;
      }
    }
    c->mu_fused_cap_used = fused_cap;
    nt = c->ntx * c->nty * c->ntz;
    capL = (fused_cap + 3) & ~3;
  } else if (which == 1) {
    tp.beta = beta_dev;
    tp.tab0 = c->d_tab[0];
    tp.tab1 = c->d_tab[1];
    tp.tab2 = c->d_tab[2];
    tp.rng = c->d_rng;
    tp.overflow = c->d_tmp_max;
    tp.frames_are_batch = frames_dev ? 1 : 0;
    tp.X = c->X;
    tp.Y = c->Y;
    tp.Z = c->Z;
    tp.K = c->K;
    tp.T = c->T;
    tp.fast_div = c->fast_div;
    tp.rcp0 = c->rcp[0];
    tp.rcp1 = c->rcp[1];
    tp.rcp2 = c->rcp[2];
    tp.ntx = (c->X + kGramTX - 1) / kGramTX;
    tp.nty = (c->Y + kGramTY - 1) / kGramTY;
    const bool same_tiles = c->tx == kGramTX && c->ty == kGramTY && c->ntz == 1 && c->d_cand_off && c->d_cand_ids;
    tp.cand_off = same_tiles ? c->d_cand_off : nullptr;
    tp.cand_ids = same_tiles ? c->d_cand_ids : nullptr;
    tp.cand_expand = c->cand_expand;
    if (gram_tc_smem_bytes(c->X, c->Y, c->Z) > (size_t)c->max_smem_optin) {
      *over = 1 << 30;  // volume too deep for whole-depth tiles in shared memory
      return 0;
    }
    nt = tp.ntx * tp.nty;
    capL = kGramRows;
  } else {
    dnmf_ctx g = *c;  // shallow copy used only for geometry helpers
    g.tx = kMuTX;
    g.ty = kMuTY;
    g.ntx = (c->X + kMuTX - 1) / kMuTX;
    g.nty = (c->Y + kMuTY - 1) / kMuTY;
    nt = g.ntx * g.nty * g.ntz;
    if (c->mu_capM == 0) {
      if (ensure(&c->d_tmp_counts, &c->tmp_counts_cap, (size_t)nt)) return 1;
      if (c->d_tmp_offsets) cudaFree(c->d_tmp_offsets);
      c->d_tmp_offsets = nullptr;
      CU(cudaMalloc((void**)&c->d_tmp_offsets, ((size_t)nt + 1) * sizeof(long long)));
      if (run_bin_count(&g, c->d_identity_beta, 1, c->d_ids_zero, 1, c->d_tmp_counts, nullptr, st)) return 1;
      scan_counts_kernel<<<1, 1024, 0, st>>>(c->d_tmp_counts, nt, c->d_tmp_offsets, c->d_tmp_max);
      CU(cudaGetLastError());
      int lmax = 0;
      CU(cudaMemcpyAsync(&lmax, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      const int capM = std::min(c->K, lmax + lmax / 2 + 8) + 1;
      c->mu_capM = (capM + 7) & ~7;
    }
    const int capM = c->mu_capM;
    mu_bs = (capM >= 32 && c->mu_block4 == 0) ? 8 : 4;  // 8x8 register blocks from 32 rows up, 4x4 below
    smem = mu_smem_bytes(capM, c->tz, c->K, mu_bs);
    if (smem > (size_t)c->max_smem_optin)
      return fail("dnmf_mu_stats: neuron lists too long for the shared-memory A panel (K_eff too large)");
    mp.beta = beta_dev;
    mp.tab0 = c->d_tab[0];
    mp.tab1 = c->d_tab[1];
    mp.tab2 = c->d_tab[2];
    mp.rng = c->d_rng;
    mp.overflow = c->d_tmp_max;
    mp.frames_are_batch = frames_dev ? 1 : 0;
    mp.X = c->X;
    mp.Y = c->Y;
    mp.Z = c->Z;
    mp.K = c->K;
    mp.T = c->T;
    mp.tz = c->tz;
    mp.ntx = g.ntx;
    mp.nty = g.nty;
    mp.ntz = g.ntz;
    mp.capM = capM;
    mp.full_depth = (c->tz == c->Z) ? 1 : 0;
    const int mb = (capM + mu_bs - 1) / mu_bs;
    mu_need = (mb * (mb + 1) / 2 + kMuThreads - 1) / kMuThreads;
    capL = (capM + 3) & ~3;
  }
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, budget / stats_bytes_per_frame(nt, capL, c->K)));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    StatsPartials sp;
    if (stats_scratch(c, nb, nt, capL, sp, st)) return 1;
    CU(cudaMemsetAsync(c->d_tmp_max, 0, sizeof(int), st));
    const float* fr = frames_dev ? frames_dev + (size_t)b0 * c->N : nullptr;
    if (which == 0) {
      FitParams q;
      if (fill_fit_params(c, q, fr, ids_dev + b0, nb, beta_dev, nullptr, st)) return 1;
      q.mu_overflow = c->d_tmp_max;
      q.cap = fused_cap;
      q.stats = sp;
      if (dispatch_stats(c, q, nb, smem, st)) return 1;
    } else if (which == 1) {
      GramTcParams q = tp;
      q.frames = frames_dev ? fr : c->d_video;
      q.frame_ids = ids_dev + b0;
      q.out = sp;
      q.b_base = 0;
      if (launch_gram_tc(q, nb, st)) return 1;
    } else {
      MuParams q = mp;
      q.frames = frames_dev ? fr : c->d_video;
      q.frame_ids = ids_dev + b0;
      q.out = sp;
      q.b_base = 0;
      int rc;
      if (mu_bs == 8) rc = mu_need <= 1 ? launch_mu<1, 8>(q, nb * nt, smem, st) : launch_mu<2, 8>(q, nb * nt, smem, st);
      else if (mu_need <= 1) rc = launch_mu<1, 4>(q, nb * nt, smem, st);
      else if (mu_need <= 2) rc = launch_mu<2, 4>(q, nb * nt, smem, st);
      else rc = launch_mu<4, 4>(q, nb * nt, smem, st);
      if (rc) return rc;
    }
    c->counters[6] += 1;
    CU(cudaMemcpyAsync(over, c->d_tmp_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (*over != 0) return 0;
    if (stats_reduce(c, sp, ids_dev + b0, nb, nt, st)) return 1;
  }
  return 0;
}

extern "C" int dnmf_mu_stats(dnmf_ctx* c, const float* frames_dev, const int32_t* frame_ids_dev, int B,
                             const float* beta_dev, void* stream) {
  if (!c || !frame_ids_dev || !beta_dev) return fail("dnmf_mu_stats: NULL argument");
  if (!c->have_footprints) return fail("dnmf_mu_stats: call dnmf_set_footprints first");
  if (!frames_dev && !c->d_video) return fail("dnmf_mu_stats: no resident video and frames_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaSetDevice(c->device));
  if (check_sticky(c, "dnmf_mu_stats") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
  {  // statistics are stored per frame id: an id listed twice has no defined result; out of range is an error
    int flags = 0;
    CU(cudaMemcpyAsync(&flags, c->d_id_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flags & 2) {
      *(volatile int*)c->h_sticky = 0;
      return fail("dnmf_mu_stats: frame id outside [0, T)");
    }
    if (flags & 1) return fail("dnmf_mu_stats: a frame id occurs twice in the batch");
  }
  if (mu_alloc(c)) return 1;
  c->gc_valid = false;
  c->mu_last_path = 0;
  c->mu_last_tc = 0;
  int over = 0;
  // 1. the fused kernel's tiles, lists and staged slices (short lists).  A tile whose list is longer than the
  //    staged capacity under the current deformation raises the overflow flag: remembered until the next
  //    dnmf_set_footprints (capacity hint, or "does not fit").
  if (fused_stats_available(c) && c->mu_force_panel == 0 && c->mu_prefer_tc == 0 && c->mu_fused_off == 0) {
    if (stats_pass(c, 0, frames_dev, frame_ids_dev, B, beta_dev, st, &over)) return 1;
    if (over == 0) {
      c->mu_last_path = 1;
      return 0;
    }
    // A list longer than the capacity: stage that many slots next time.  Otherwise a window wider than the
    // staged slices raised the flag: not a matter of capacity.
    if (over > c->mu_fused_cap_used) c->mu_fused_need = over;
    else c->mu_fused_off = 1;
  }
  // 2. tensor-core panel kernel: lists of up to 127 neurons per 8 x 8 x Z tile
  if (c->mu_force_panel == 0 || c->mu_prefer_tc != 0) {
    if (stats_pass(c, 1, frames_dev, frame_ids_dev, B, beta_dev, st, &over)) return 1;
    if (over == 0) {
      c->mu_last_tc = 1;
      return 0;
    }
  }
  // 3. SIMT panel kernel: any list length the shared-memory panel holds; grows once when a list outgrew it
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (stats_pass(c, 2, frames_dev, frame_ids_dev, B, beta_dev, st, &over)) return 1;
    if (over == 0) return 0;
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
    c->mu_prefer_tc = (force_panel & 8) != 0;
  }
  if (last_path_out) *last_path_out = c->mu_last_path | (c->mu_last_sparse << 1) | (c->mu_last_tc << 2);
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
    if (W > 2 * Ws)
      mu_sweep_sparse_kernel<4><<<grid, 256, 0, st>>>(
          c->d_Gc, c->d_mu_nbr, W, Ws, c->d_b, c->d_Cd[c->cd_cur], c->d_Cd[c->cd_cur ^ 1], c->T, c->K, gamma, use_gamma,
          halo_prev_dev, halo_next_dev);
    else
      mu_sweep_sparse_kernel<1><<<grid, 256, 0, st>>>(
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
  if (check_sticky(c, "dnmf_iwarp") || sanitize_ids(c, frame_ids_dev, B, st, &frame_ids_dev)) return 1;
  const size_t total = c->N * (size_t)B;
  if (c->N > 0xffffffffull) return fail("dnmf_iwarp: more than 2^32 voxels per frame");
  if (ensure(&c->d_keys, &c->keys_cap, total + (total + 1) / 2)) return 1;  // fp64 distances + 32-bit winners
  unsigned* winner = reinterpret_cast<unsigned*>(c->d_keys + total);
  CU(cudaMemsetAsync(c->d_keys, 0xff, (total + (total + 1) / 2) * sizeof(unsigned long long), st));
  Geom g = geom_of(c);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  iwarp_vote_kernel<0><<<blocks, 256, 0, st>>>(g, frame_ids_dev, B, beta_dev, c->d_keys, winner);
  CU(cudaGetLastError());
  iwarp_vote_kernel<1><<<blocks, 256, 0, st>>>(g, frame_ids_dev, B, beta_dev, c->d_keys, winner);
  CU(cudaGetLastError());
  iwarp_gather_kernel<<<blocks, 256, 0, st>>>(g, frames_dev ? frames_dev : c->d_video, frames_dev ? 1 : 0,
                                             frame_ids_dev, B, beta_dev, c->d_keys, winner, out_dev);
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
