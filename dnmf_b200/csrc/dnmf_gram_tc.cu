// Trace statistics of dense neuron lists on the 5th-generation tensor cores (kernel 3b, dense form).
//
//   G_t = A_t^T A_t,  b_t = A_t^T Y_t          Demix/dNMF.py:141-142 (numpy fp64 einsum in the reference)
//
// With ~100 neurons reaching every tile (BASELINE configuration 4: K = 1000, sigma = 6) the per-tile Gram
// [A|Y]^T [A|Y] is a real dense contraction: M = N = listed neurons + the Y pseudo-neuron (<= 128), K = the
// tile's voxels.  One CTA (9 warps) per (frame, 8 x 8 x Z tile):
//   * prologue: beta_t -> conservative window -> neuron list (same device code as the binning kernel) -> the
//     listed neurons' table slices staged in shared memory, two slots per float4 (G_j, G_j+1, D_j, D_j+1);
//   * producers: four warps per 8 x 4 (x, y) half of the tile walk the z planes in ascending order and split the
//     rows.  Per plane they evaluate the closed-form footprint value of every listed neuron at the half's 32 voxels
//     from the staged slices (3 LDS.128 + 5 packed FP32x2 operations per slot pair and voxel, no global memory, no
//     transcendental) and write them as ONE K-major panel stage [128 rows][32 voxels] in the canonical
//     SWIZZLE_128B layout -- twice: hi = tf32(a) and lo = a - hi.  A half alternates between two stages (even / odd
//     planes; four stages in all), so the tensor pipe reads plane z while plane z + 1 is being written;
//   * tensor cores: one thread issues tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x N x 8, N = list length
//     rounded up to 16) on each stage-plane: SYRK has ONE operand, so the same shared-memory panel serves the A and
//     the B descriptor.  fp32-accurate products come from the 3xTF32 split hi hi^T + hi lo^T + lo hi^T, of which the
//     last term is the transpose of the second: P = hi hi^T and Q = hi lo^T are accumulated (two MMAs per 8 voxels,
//     not three) and the epilogue forms P + Q + Q^T.  The stage-planes are dealt round-robin over nP = 4 (N <= 96) or 3
//     TMEM accumulators for P, Q has one: the truncating fp32 accumulation of the tensor pipe (measured with
//     tools/experiments/syrk_tf32_umma.cu: relative error 3e-6 per 256 accumulated voxels, growing linearly) stays
//     at 1/nP of the chain length; Q is 2^-11 of P and its chain length does not matter.  tcgen05.commit ->
//     mbarrier hands the stage back to its producers;
//   * epilogue: Q is transposed through shared memory (the panel stages are free by then); every thread reads its
//     row of the accumulators with tcgen05.ld, adds them in a fixed order and writes the tile-frame's partial block;
//     the row-owner second stage (stats_reduce_kernel) sums the blocks of a frame in ascending tile order in fp64:
//     deterministic, no atomics.
#include "dnmf_common.h"

namespace dnmf {

namespace {

constexpr int kTcProducers = 8;                       // four per (x, y) half of the tile, they split the rows
constexpr int kTcWarps = kTcProducers + 1;            // + the MMA warp
constexpr int kTcThreads = 32 * kTcWarps;
constexpr int kTcStageBytes = kGramRows * 128;        // 128 rows x 32 voxels of tf32
constexpr int kTcPairs = kGramRows / 2;               // slot pairs per slice entry
constexpr int kTcEntryBytes = (kTcPairs + 1) * 16;    // padded: consecutive entries start 16 B apart mod 128

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// The staged slices are read-only once the CTA barrier after staging has passed: the loads are plain (not
// volatile, no memory clobber) so that the compiler may hoist them over the panel stores of the previous slot
// pairs; the panel stores are volatile (ordered among themselves and with the fences) but do not clobber memory.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(addr), "f"(v), "n"(OFF));
}
__device__ __forceinline__ void sts32r(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}

// hi = the value rounded to tf32 (round to nearest, ties away: what cvt.rna.tf32.f32 does for finite inputs; two
// integer instructions instead of the guarded conversion sequence), lo = the exact remainder.  lo is handed to the
// tensor core as it is: the MMA reads the upper 19 bits, a truncation of 2^-11 |lo| <= 2^-22 |a|.
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
  lo = a - hi;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcSmem {
  size_t slices, ytile, misc, bytes;
};

}  // namespace

__host__ __device__ static TcSmem gram_tc_layout(int wsum, int Z) {
  TcSmem s;
  s.slices = (size_t)4 * 2 * kTcStageBytes;                        // panel stages come first (1024-byte aligned)
  s.ytile = s.slices + (size_t)(wsum + 1) * kTcEntryBytes;  // + one all-zero entry (x slice of lanes outside the volume)
  s.misc = s.ytile + (((size_t)kGramTX * kGramTY * Z + 3) & ~(size_t)3) * 4;
  s.bytes = s.misc + kGramRows * 2 * (2 + kTcWarps) + 32 * 4 + 32 * 4 + 8 * 8 + 16 + 1024;  // list + per-warp lists, beta, ints, barriers, tmem ptr, slack
  return s;
}

size_t gram_tc_smem_bytes(int X, int Y, int Z) {
  const int w0 = std::min(kGramTX + 4, X + 3), w1 = std::min(kGramTY + 4, Y + 3), w2 = Z + 3;
  return gram_tc_layout(w0 + w1 + w2, Z).bytes;
}

__global__ void __launch_bounds__(kTcThreads, 1) gram_tc_kernel(const __grid_constant__ GramTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  const int wmax0 = min(kGramTX + 4, p.X + 3), wmax1 = min(kGramTY + 4, p.Y + 3), wmax2 = p.Z + 3;
  const int wsum = wmax0 + wmax1 + wmax2;
  const TcSmem lay = gram_tc_layout(wsum, p.Z);
  unsigned char* sSl = base + lay.slices;
  float* sY = reinterpret_cast<float*>(base + lay.ytile);
  unsigned short* sList = reinterpret_cast<unsigned short*>(base + lay.misc);
  unsigned short* sWList = sList + kGramRows;                        // [warps][kGramRows]: matches found by each warp
  float* sBeta = reinterpret_cast<float*>(sWList + kTcWarps * kGramRows + (kTcWarps & 1) * kGramRows);
  int* sInt = reinterpret_cast<int*>(sBeta + 32);
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sInt + 32);   // full[4], empty[4]
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 8);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.y;
#ifdef DNMF_TC_TIMING
  const bool tprint = blockIdx.x == 200 && (blockIdx.y == 3 || blockIdx.y == 17);
  const long long T0 = clock64();
  long long Twait = 0, Tmma0 = 0, Tloop0 = 0, Tloop1 = 0;
#define TC_T(x) x
#else
#define TC_T(x)
#endif
  const int nt = p.ntx * p.nty;
  const int bx = tile % p.ntx, by = tile / p.ntx;
  const int t = p.frame_ids[b];
  DNMF_DASSERT(t >= 0 && t < p.T);
  const int x0 = bx * kGramTX, y0 = by * kGramTY;
  const int nx = min(kGramTX, p.X - x0), ny = min(kGramTY, p.Y - y0), nz = p.Z;
  const size_t Nvox = (size_t)p.X * p.Y * p.Z;
  const float* __restrict__ frame = p.frames + (size_t)(p.frames_are_batch ? b : t) * Nvox;
  const size_t tf = (size_t)(p.b_base + b) * nt + tile;  // tile-frame index in the partial buffers

  // ---- beta, Y tile, window ----
  if (tid < 30) sBeta[tid] = p.beta[(size_t)tid * p.T + t];
  {
    const int run = ny * p.Z;
    for (int lx = warp; lx < nx; lx += kTcWarps) {
      const float* src = frame + ((size_t)(x0 + lx) * p.Y + y0) * p.Z;
      for (int e = lane; e < run; e += 32) sY[lx * kGramTY * p.Z + e] = __ldg(src + e);
    }
  }
  for (int e = tid; e < kTcEntryBytes / 16; e += kTcThreads)   // the all-zero slice entry
    reinterpret_cast<float4*>(sSl + (size_t)wsum * kTcEntryBytes)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (tid < 3) {
    const int s = tid == 0 ? p.X : (tid == 1 ? p.Y : p.Z);
    int wlo, whi;
    tile_window_axis(sBeta + tid, 3, (float)x0, (float)y0, 0.f, (float)(x0 + nx - 1), (float)(y0 + ny - 1),
                     (float)(nz - 1), s, wlo, whi);
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
  // ---- neuron list (ascending k): every warp scans its share of the neurons ONCE into its own segment, the
  // segments are then concatenated in warp order ----
  {
    // candidates: the fit kernel's static per-tile lists (identity windows expanded by cand_expand nodes, ascending k)
    // when they were built for these very tiles and the window stays inside the expanded box; else all K neurons
    long long r0 = 0;
    int r1 = p.K;
    const int* __restrict__ cand = nullptr;
    if (p.cand_off != nullptr) {
      const int e = p.cand_expand;
      const bool inside = wlo[0] >= max(x0 - 1, -2) - e && whi[0] <= min(x0 + nx, p.X) + e &&
                          wlo[1] >= max(y0 - 1, -2) - e && whi[1] <= min(y0 + ny, p.Y) + e &&
                          wlo[2] >= -1 - e && whi[2] <= p.Z + e;
      if (inside) {
        r0 = p.cand_off[tile];
        r1 = (int)(p.cand_off[tile + 1] - r0);
        cand = p.cand_ids + r0;
      }
    }
    const int per = ((r1 + kTcThreads - 1) / kTcThreads) * 32;
    const int kb = warp * per;
    int cnt = 0;
    for (int k0 = kb; k0 < kb + per; k0 += 32) {
      const int idx = k0 + lane;
      bool ok = false;
      int k = 0;
      if (idx < r1) {  // the six range bounds as three independent 8-byte loads, then the test
        k = cand ? __ldg(cand + idx) : idx;
        const int2* r2 = reinterpret_cast<const int2*>(p.rng + (size_t)k * 6);
        const int2 rx = __ldg(r2), ry = __ldg(r2 + 1), rz = __ldg(r2 + 2);
        const int r[6] = {rx.x, rx.y, ry.x, ry.y, rz.x, rz.y};
        ok = neuron_in_window(r, wlo, whi);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (ok && pos < kGramRows) sWList[warp * kGramRows + pos] = (unsigned short)k;
      cnt += __popc(m);
    }
    if (lane == 0) sInt[8 + warp] = cnt;
  }
  __syncthreads();
  int L = 0, off = 0;
#pragma unroll
  for (int w = 0; w < kTcWarps; ++w) {
    const int c = sInt[8 + w];
    if (w < warp) off += c;
    L += c;
  }
  const int W0 = whi[0] - wlo[0] + 1, W1 = whi[1] - wlo[1] + 1, W2 = whi[2] - wlo[2] + 1;
  if (L + 1 > kGramRows || W0 > wmax0 || W1 > wmax1 || W2 > wmax2) {  // loud: the host reruns the SIMT panel kernel
    if (tid == 0) atomicMax(p.overflow, max(L + 1, kGramRows + 1));
    return;
  }
  DNMF_DASSERT(wlo[0] >= -2 && whi[0] <= p.X && wlo[1] >= -2 && whi[1] <= p.Y && wlo[2] >= -2 && whi[2] <= p.Z);  // table rows
  DNMF_DASSERT(L + 1 <= p.out.capL + 1 && p.out.capL < p.out.ld);
  if (tid == 0) p.out.count[tf] = L;
  if (L == 0) return;
  for (int i = lane; i < sInt[8 + warp]; i += 32) {
    const int k = sWList[warp * kGramRows + i], pos = off + i;
    sList[pos] = (unsigned short)k;
    p.out.ids[tf * p.out.capL + pos] = (unsigned short)k;
    p.out.slot_of[((size_t)(p.b_base + b) * p.K + k) * nt + tile] = (unsigned short)pos;
  }

  // ---- TMEM (all 512 columns: one 128-column accumulator per panel stage), stage barriers ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(sTmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&sBar[w])), "n"(4));  // full: the four producer warps of the stage's (x, y) half
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sBar[4 + w])));  // empty: tcgen05.commit
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();  // also: sList complete
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *sTmem;

  // ---- table slices of the listed neurons: [entry][slot pair] float4 (G_2p, G_2p+1, D_2p, D_2p+1) ----
  const int npair = (L + 2) >> 1;  // rows 0 .. L (the Y pseudo-row L overwrites its half of the last pair)
  {
    const int sX3 = p.X + 3, sY3 = p.Y + 3, sZ3 = p.Z + 3;
    const int Wt = W0 + W1 + W2;
    // work items (table entry, slot pair) dealt round-robin to all threads, four items (eight 8-byte gathers) in flight
    // per thread before the first store: two dependent rounds of L2 latency instead of the eleven of one entry per
    // warp and 32 pairs per pass
    const int items = Wt * npair;
    const unsigned recN = npair > 1 ? 0xFFFFFFFFu / (unsigned)npair + 1u : 0u;  // floor(i / npair) == umulhi(i, recN)
    constexpr int kBatch = 4;
    for (int i0 = tid; i0 < items; i0 += kBatch * kTcThreads) {
      float2 va[kBatch], vb[kBatch];
      float4* dst[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * kTcThreads;
        va[u] = vb[u] = make_float2(0.f, 0.f);
        dst[u] = nullptr;
        if (i < items) {
          const int e = npair > 1 ? (int)__umulhi((unsigned)i, recN) : i;
          const int pp = i - e * npair;
          const float2* src;
          int row, ent;
          if (e < W0) {
            src = p.tab0 + (wlo[0] + 2 + e);
            row = sX3;
            ent = e;
          } else if (e < W0 + W1) {
            src = p.tab1 + (wlo[1] + 2 + (e - W0));
            row = sY3;
            ent = wmax0 + (e - W0);
          } else {
            src = p.tab2 + (wlo[2] + 2 + (e - W0 - W1));
            row = sZ3;
            ent = wmax0 + wmax1 + (e - W0 - W1);
          }
          const int j = 2 * pp;
          DNMF_DASSERT(j >= L || sList[j] < p.K);
          if (j < L) va[u] = __ldg(src + (size_t)sList[j] * row);
          if (j + 1 < L) vb[u] = __ldg(src + (size_t)sList[j + 1] * row);
          dst[u] = reinterpret_cast<float4*>(sSl + (size_t)ent * kTcEntryBytes) + pp;
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
        if (dst[u]) *dst[u] = make_float4(va[u].x, vb[u].x, va[u].y, vb[u].y);
    }
  }
  __syncthreads();
  TC_T(const long long T1 = clock64();)

  // ---- producers (warps 0..7) and the MMA warp (warp 8).  Stage s belongs to the 8 x 4 (x, y) half s & 1 and the z
  // planes of parity s >> 1; the four producer warps of the half arrive on full[s], the MMA warp issues the stage's
  // MMAs and tcgen05.commit hands the stage back through empty[s].  (stg, colq) are the epilogue's roles: TMEM lane
  // quadrant and column half ----
  const int stg = warp & 3, colq = (warp >> 2) & 1;
  const int npad = (L + 1 + 15) & ~15;  // MMA N
  // TMEM: nP accumulators for P = hi hi^T (the truncating fp32 accumulation of the tensor pipe makes the chain length
  // matter: the stage-planes are dealt over nP chains) + ONE for the cross term Q = hi lo^T (2^-11 of P: its chain
  // length does not matter).  The third product of the 3xTF32 split, lo hi^T, is Q^T: the epilogue adds it from a
  // shared-memory transpose, and the tensor pipe and its shared-memory operand reads do 2/3 of the work.
  const int accStride = npad <= 96 ? 96 : 128;
  const int nP = npad <= 96 ? 4 : 3;
  const int nz0 = (nz + 1) >> 1, nz1 = nz >> 1;  // planes of parity 0 / 1
  if (warp == kTcProducers) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kGramRows >> 4) << 24);
      const uint32_t sbase = smem_addr(base), bar0 = smem_addr(&sBar[0]);
      const uint32_t accQ = tmem + (uint32_t)(nP * accStride);
      int chain = 0;  // stage-planes are dealt round-robin to the P accumulators: nP chains of equal length
      for (int u = 0; u < nz0; ++u) {
#pragma unroll
        for (int sI = 0; sI < 4; ++sI) {
          if (u >= ((sI >> 1) ? nz1 : nz0)) continue;
          TC_T(const long long w0 = clock64();)
          mbar_wait(bar0 + 8u * sI, (uint32_t)(u & 1));  // the producer warps have written plane u of this stage
          TC_T(Twait += clock64() - w0; if (u == 0 && sI == 0) Tmma0 = clock64();)
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint64_t dh = make_desc(sbase + (uint32_t)sI * 2u * kTcStageBytes);
          const uint64_t dl = make_desc(sbase + (uint32_t)sI * 2u * kTcStageBytes + kTcStageBytes);
          const int a = chain % nP;
          const uint32_t accP = tmem + (uint32_t)(a * accStride);
          const bool first_p = chain < nP, first_q = chain == 0;
          ++chain;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // UMMA_K = 8 tf32 = 32 bytes along the swizzled row: +2 in the address field
            const uint64_t ah = dh + (uint64_t)(2 * k), al = dl + (uint64_t)(2 * k);
            umma_tf32(accP, ah, ah, idesc, (!first_p || k > 0) ? 1u : 0u);
            umma_tf32(accQ, ah, al, idesc, (!first_q || k > 0) ? 1u : 0u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 32u + 8u * sI)
                       : "memory");
        }
      }
    }
    __syncwarp();
  } else {
    // Producer warp (sub, quarter): the 8 x 4 (x, y) half `sub` of the tile, every z plane in ascending order, a quarter
    // of the 8-row atoms.  Plane z goes to panel stage sub + 2 (z & 1): the four warps of a half alternate between two
    // stages, so the tensor pipe works on plane z while they write plane z + 1.
    const int sub = warp & 1, quarter = warp >> 1;
    const int lx = lane & 7, ly = (lane >> 3) + 4 * sub;
    const int gx = x0 + lx, gy = y0 + ly;
    const bool valid = (gx < p.X) && (gy < p.Y);
    const float xf = (float)gx, yf = (float)gy;
    float c0[3], c1[3], c2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {  // same operation order as the fused kernel's Horner form in z
      float v = sBeta[d];
      v = fmaf(sBeta[3 + d], xf, v);
      v = fmaf(sBeta[6 + d], yf, v);
      v = fmaf(sBeta[12 + d], xf * xf, v);
      v = fmaf(sBeta[15 + d], yf * yf, v);
      v = fmaf(sBeta[21 + d], xf * yf, v);
      c0[d] = v;
      c1[d] = fmaf(sBeta[27 + d], yf, fmaf(sBeta[24 + d], xf, sBeta[9 + d]));
      c2[d] = sBeta[18 + d];
      if (p.fast_div) {  // exact doubling: the plane loop then evaluates 2q, what sample_coord_fast expects
        c0[d] += c0[d];
        c1[d] += c1[d];
        c2[d] += c2[d];
      }
    }
    const float sm1[3] = {(float)(p.X - 1), (float)(p.Y - 1), (float)(p.Z - 1)};
    const float rcp[3] = {p.rcp0, p.rcp1, p.rcp2};
    const int sz[3] = {p.X, p.Y, p.Z};
    const uint32_t hi_base = smem_addr(base) + (uint32_t)sub * 2u * kTcStageBytes;  // stage `sub` (even planes)
    constexpr uint32_t kOddPlanes = 4u * kTcStageBytes;                                 // stage sub + 2 is this much further
    const uint32_t slx = smem_addr(sSl), sly = slx + (uint32_t)wmax0 * kTcEntryBytes,
                   slz = sly + (uint32_t)wmax1 * kTcEntryBytes, slzero = slx + (uint32_t)wsum * kTcEntryBytes;
    const uint32_t full0 = smem_addr(&sBar[sub]), empty0 = smem_addr(&sBar[4 + sub]);  // odd planes: + 16 bytes (stage + 2)
    const int natom = (npair + 3) >> 2;
    const int at0 = (natom * quarter) >> 2, at1 = (natom * (quarter + 1)) >> 2;
    const bool owns_y = quarter == 3;   // the Y pseudo-row lives in the last atom
    // address of this lane's column in row r8 of atom at0 of the hi stage (16-byte chunks XOR-swizzled with the
    // row); the lo stage is kTcStageBytes further, the next atom 1024 bytes further
    uint32_t rowaddr[8];
#pragma unroll
    for (int r8 = 0; r8 < 8; ++r8)
      rowaddr[r8] = hi_base + (uint32_t)at0 * 1024u + (uint32_t)(r8 * 128 + (((lane >> 2) ^ r8) << 4) + ((lane & 3) << 2));
    const int yL_r8 = L & 7;
    const uint32_t yaddr = hi_base + (uint32_t)(L >> 3) * 1024u + (uint32_t)(yL_r8 * 128 + (((lane >> 2) ^ yL_r8) << 4) + ((lane & 3) << 2));
    for (int z = 0; z < nz; ++z) {
      const int u = z >> 1;
      const uint32_t zoff = (z & 1) ? kOddPlanes : 0u, boff = (z & 1) ? 16u : 0u;
      const float zf = (float)z;
      int ii[3];
      float ff[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float q = fmaf(zf, fmaf(zf, c2[d], c1[d]), c0[d]);
        const float ix = p.fast_div ? sample_coord_fast(q, sm1[d], rcp[d], 0.5f * sm1[d]) : sample_coord(q, sm1[d]);
        split_coord(ix, sz[d], ii[d], ff[d]);
      }
      const uint32_t ax = (valid ? slx + (uint32_t)min(max(ii[0] - wlo[0], 0), W0 - 1) * kTcEntryBytes : slzero) + (uint32_t)at0 * 64u;
      const uint32_t ay = sly + (uint32_t)min(max(ii[1] - wlo[1], 0), W1 - 1) * kTcEntryBytes + (uint32_t)at0 * 64u;
      const uint32_t az = slz + (uint32_t)min(max(ii[2] - wlo[2], 0), W2 - 1) * kTcEntryBytes + (uint32_t)at0 * 64u;
      const float2 ff0 = make_float2(ff[0], ff[0]), ff1 = make_float2(ff[1], ff[1]), ff2 = make_float2(ff[2], ff[2]);
      TC_T(const long long w0 = clock64();)
      if (u > 0) mbar_wait(empty0 + boff, (uint32_t)((u - 1) & 1));  // the MMAs that read this stage have completed
      TC_T(Twait += clock64() - w0; if (z == 0) Tloop0 = clock64();)
      uint32_t pofs = 0, rofs = zoff;
#pragma unroll 2
      for (int at = at0; at < at1; ++at, pofs += 64u, rofs += 1024u) {
        float4 ex[4], ey[4], ez[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ex[q] = lds128(ax + pofs + q * 16);
          ey[q] = lds128(ay + pofs + q * 16);
          ez[q] = lds128(az + pofs + q * 16);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a0 = __ffma2_rn(ff0, make_float2(ex[q].z, ex[q].w), make_float2(ex[q].x, ex[q].y));
          const float2 a1 = __ffma2_rn(ff1, make_float2(ey[q].z, ey[q].w), make_float2(ey[q].x, ey[q].y));
          const float2 a2 = __ffma2_rn(ff2, make_float2(ez[q].z, ez[q].w), make_float2(ez[q].x, ez[q].y));
          const float2 a = __fmul2_rn(__fmul2_rn(a0, a1), a2);
          float hx, lx_, hy, ly_;
          split_tf32(a.x, hx, lx_);
          split_tf32(a.y, hy, ly_);
          sts32<0>(rowaddr[2 * q] + rofs, hx);
          sts32<0>(rowaddr[2 * q + 1] + rofs, hy);
          sts32<kTcStageBytes>(rowaddr[2 * q] + rofs, lx_);
          sts32<kTcStageBytes>(rowaddr[2 * q + 1] + rofs, ly_);
        }
      }
      if (owns_y) {  // the Y pseudo-neuron (row L): b_t = A_t^T Y_t rides in column L of the same product
        const float y = valid ? sY[(lx * kGramTY + ly) * p.Z + z] : 0.f;
        float h, l;
        split_tf32(y, h, l);
        sts32r(yaddr + zoff, h);
        sts32<kTcStageBytes>(yaddr + zoff, l);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + boff) : "memory");
    }
    // a stage's phases can only be followed by a waiter that saw every one of them: its own producers
    TC_T(Tloop1 = clock64();)
    if (nz0 > 0) mbar_wait(empty0, (uint32_t)((nz0 - 1) & 1));
    if (nz1 > 0) mbar_wait(empty0 + 16u, (uint32_t)((nz1 - 1) & 1));
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();  // all four stages' last commits are now known to every warp
  TC_T(const long long T2 = clock64();)
  const int nacc = min(nP, 2 * (nz0 + nz1));  // P accumulators that received a stage-plane
  asm volatile("tcgen05.fence::after_thread_sync;");
  // a warp reads the TMEM lanes 32 * (warp % 4) ..: warps w and w + 4 share a row quadrant and split the columns
  const int row = stg * 32 + lane;
  float* sQ = reinterpret_cast<float*>(base);  // [128][kQPitch]: the panel stages are free (every MMA has completed)
  constexpr int kQPitch = 129;
  const uint32_t tq = tmem + ((uint32_t)(stg * 32) << 16) + (uint32_t)(nP * accStride);
  if (warp < kTcProducers) {
    for (int cb = 32 * colq; cb < npad; cb += 32 * (kTcProducers / 4)) {
      uint32_t r[32];
      tmem_ld32(tq + (uint32_t)cb, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) sQ[row * kQPitch + cb + i] = __uint_as_float(r[i]);
    }
  }
  __syncthreads();
  // the finished block goes through shared memory once more (behind sQ; the slices are free as well), so that every
  // thread can store BOTH triangles of its row from the upper one: element (row, col) = S[min][max].  The stored block
  // is bitwise symmetric by construction and the second stage reads rows only (coalesced; reading element
  // (min slot, max slot) column-wise cost it 8x the sectors for half of its entries).
  float* sS = sQ + kGramRows * kQPitch;
  if (warp < kTcProducers) {
    for (int cb = 32 * colq; cb < npad; cb += 32 * (kTcProducers / 4)) {
      float sum[32];
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(stg * 32) << 16) + (uint32_t)cb;
      tmem_ld32(taddr, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) sum[i] = __uint_as_float(r[i]);
      for (int w = 1; w < nacc; ++w) {
        tmem_ld32(taddr + (uint32_t)(w * accStride), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) sum[i] += __uint_as_float(r[i]);
      }
      // the cross terms, Q[row][col] + Q[col][row]
#pragma unroll
      for (int i = 0; i < 32; ++i) sS[row * kQPitch + cb + i] = sum[i] + (sQ[row * kQPitch + cb + i] + sQ[(cb + i) * kQPitch + row]);
    }
  }
  __syncthreads();
  if (warp < kTcProducers && row < L) {
    float* out = p.out.vals + tf * (size_t)p.out.capL * p.out.ld + (size_t)row * p.out.ld;
    for (int cb = 32 * colq; cb < npad; cb += 32 * (kTcProducers / 4)) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const int col = cb + i;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = col + q;
          v[q] = c >= row ? sS[row * kQPitch + c] : sS[c * kQPitch + row];
        }
        if (col + 3 < L) {
          *reinterpret_cast<float4*>(out + col) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col + q < L) out[col + q] = v[q];
            else if (col + q == L) out[p.out.capL] = v[q];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
  TC_T(if (tprint && lane == 0 && (warp == 0 || warp == 3 || warp == 8)) printf("tc-timing frame %d warp %d L %d: prologue %lld, loop %lld (first plane at +%lld, producer done at +%lld), waits %lld, epilogue %lld\n", (int)blockIdx.y, warp, L, T1 - T0, T2 - T1, (warp == 8 ? Tmma0 : Tloop0) - T1, Tloop1 - T1, Twait, clock64() - T2);)
}

int launch_gram_tc(const GramTcParams& p, int B, cudaStream_t st) {
  const size_t smem = gram_tc_smem_bytes(p.X, p.Y, p.Z);
  static size_t configured[64] = {0};
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    CU(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {
    GramTcParams q = p;
    const int nb = std::min(65535, B - b0);
    q.b_base = p.b_base + b0;
    q.frame_ids = p.frame_ids + b0;
    if (p.frames_are_batch) q.frames = p.frames + (size_t)b0 * p.X * p.Y * p.Z;
    gram_tc_kernel<<<dim3((unsigned)(p.ntx * p.nty), (unsigned)nb), kTcThreads, smem, st>>>(q);
    CU(cudaGetLastError());
  }
  return 0;
}

}  // namespace dnmf
