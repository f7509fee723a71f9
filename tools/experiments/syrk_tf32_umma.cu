// EXPERIMENT for round 2 -- NOT part of the product library, NOT yet run on hardware:
// it was written after the round's GPU budget was spent and has only been through nvcc/ptxas for sm_100a.
// First thing to do with it: `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o syrk_tf32_umma
// tools/experiments/syrk_tf32_umma.cu && ./syrk_tf32_umma` on a B200 under a short `timeout`.
//
// What it is: the Gram matrix G = P P^T of a [128 x KV] fp32 panel (rows = listed neurons + the Y pseudo-neuron of
// one tile, columns = voxels) on the 5th-generation tensor cores, the way the dense-neuron trace statistics
// (mu_stats_kernel, Demix/dNMF.py:141-142) would use them:
//   * SYRK has ONE operand: the panel is K-major for A (M x K) and for B (N x K), so one shared-memory buffer
//     (canonical K-major SWIZZLE_128B layout, 32 voxels = 128 B per row, 8-row atoms of 1 KB) serves both
//     descriptors of tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x 128 x 8);
//   * fp32-accurate products from the 3xTF32 split: hi = tf32(a), lo = tf32(a - hi), D += hi hi^T + hi lo^T + lo hi^T;
//   * D[128 x 128] fp32 lives in 128 TMEM columns for the whole panel and is read once with tcgen05.ld;
//   * all 128 threads produce panel stages (two stages), one thread issues the MMAs, tcgen05.commit -> mbarrier
//     recycles a stage.
// The host checks G against a double-precision reference and prints the achieved TFLOP/s (counting 2*128*128*KV
// per panel, i.e. the useful fp32-equivalent work, not the 3 MMAs).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

namespace {

constexpr int kRows = 128;       // M = N
constexpr int kChunk = 32;       // voxels per stage: one 128-byte swizzle row of tf32
constexpr int kStageFloats = kRows * kChunk;
constexpr int kTmemCols = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor): K-major, SWIZZLE_128B,
// 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row atoms, bits [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // layout type: SWIZZLE_128B
  return d;
}

// InstrDescriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate)
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

// One CTA (128 threads) per panel.  P: [panels][128][KV] fp32, KV a multiple of 32.  G: [panels][128][128].
__global__ void __launch_bounds__(128) syrk_tf32_kernel(const float* __restrict__ P, float* __restrict__ G, int KV) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment of every stage: the swizzle pattern is a function of address bits [4,10)
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // the launch reserves 1 KB of slack
  float* sHi = reinterpret_cast<float*>(base);                // [2][128][32]
  float* sLo = sHi + 2 * kStageFloats;                         // [2][128][32]
  __shared__ __align__(8) unsigned long long sBar[2];
  __shared__ uint32_t sTmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* panel = P + (size_t)blockIdx.x * kRows * KV;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sTmem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sBar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = sTmem;

  const int nch = KV / kChunk;
  // thread t owns panel row t: byte offset of its 128-byte row inside a stage, and its swizzle key
  const int atom = tid >> 3, r8 = tid & 7;
  for (int c = 0; c < nch; ++c) {
    const int s = c & 1;
    if (c >= 2) mbar_wait(smem_u32(&sBar[s]), (uint32_t)(((c >> 1) - 1) & 1));  // MMAs of chunk c-2 are done with stage s
    const float4* src = reinterpret_cast<const float4*>(panel + (size_t)tid * KV + (size_t)c * kChunk);
    unsigned char* hi_row = reinterpret_cast<unsigned char*>(sHi + s * kStageFloats) + atom * 1024 + r8 * 128;
    unsigned char* lo_row = reinterpret_cast<unsigned char*>(sLo + s * kStageFloats) + atom * 1024 + r8 * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // eight 16-byte chunks of the row, XOR-swizzled with the row index inside the atom
      const float4 v = __ldg(src + j);
      float4 h, l;
      h.x = to_tf32(v.x), h.y = to_tf32(v.y), h.z = to_tf32(v.z), h.w = to_tf32(v.w);
      l.x = to_tf32(v.x - h.x), l.y = to_tf32(v.y - h.y), l.z = to_tf32(v.z - h.z), l.w = to_tf32(v.w - h.w);
      const int off = ((j ^ r8) << 4);
      *reinterpret_cast<float4*>(hi_row + off) = h;
      *reinterpret_cast<float4*>(lo_row + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA's async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint64_t dh = make_desc(smem_u32(sHi + s * kStageFloats));
      const uint64_t dl = make_desc(smem_u32(sLo + s * kStageFloats));
#pragma unroll
      for (int k = 0; k < kChunk / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes along the swizzled row: +2 in the address field
        const uint64_t ah = dh + (uint64_t)(2 * k), al = dl + (uint64_t)(2 * k);
        umma_tf32(tmem, ah, ah, (c > 0 || k > 0) ? 1u : 0u);
        umma_tf32(tmem, ah, al, 1u);
        umma_tf32(tmem, al, ah, 1u);
      }
      // arrives on the stage's barrier when every MMA issued so far has completed (implies fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sBar[s]))
                   : "memory");
    }
  }
  {  // commits complete in issue order: the last one covers everything
    const int last = nch - 1;
    mbar_wait(smem_u32(&sBar[last & 1]), (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;");

  // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (= rows of D), 32 columns at a time
  float* out = G + (size_t)blockIdx.x * kRows * kRows + (size_t)(warp * 32 + lane) * kRows;
#pragma unroll 1
  for (int c0 = 0; c0 < kRows; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
#pragma unroll
    for (int i = 0; i < 32; ++i) out[c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

}  // namespace

int main(int argc, char** argv) {
  const int KV = argc > 1 ? atoi(argv[1]) : 2688;      // 16 x 8 x 21 voxels of a cfg4 tile
  const int panels = argc > 2 ? atoi(argv[2]) : 148 * 8;
  if (KV % kChunk != 0 || KV <= 0) {
    fprintf(stderr, "KV must be a positive multiple of %d\n", kChunk);
    return 1;
  }
  std::vector<float> h((size_t)panels * kRows * KV);
  uint32_t st = 12345u;
  for (auto& x : h) {
    st = st * 1664525u + 1013904223u;
    x = (float)(st >> 8) * (1.0f / 16777216.0f);       // footprint values live in [0, 1)
  }
  float *dP = nullptr, *dG = nullptr;
  CU(cudaMalloc(&dP, h.size() * sizeof(float)));
  CU(cudaMalloc(&dG, (size_t)panels * kRows * kRows * sizeof(float)));
  CU(cudaMemcpy(dP, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)4 * kStageFloats * sizeof(float) + 1024;  // + slack for the 1 KB alignment
  CU(cudaFuncSetAttribute(syrk_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  syrk_tf32_kernel<<<panels, 128, smem>>>(dP, dG, KV);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  std::vector<float> g((size_t)kRows * kRows);
  CU(cudaMemcpy(g.data(), dG, g.size() * sizeof(float), cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (int i = 0; i < kRows; i += 7)
    for (int j = 0; j < kRows; j += 5) {
      double ref = 0.0;
      for (int k = 0; k < KV; ++k) ref += (double)h[(size_t)i * KV + k] * (double)h[(size_t)j * KV + k];
      worst = std::max(worst, std::fabs((double)g[(size_t)i * kRows + j] - ref) / std::fabs(ref));
    }
  printf("panel 0: max relative error of G vs fp64 = %.3e (3xTF32 should be ~1e-6; plain TF32 ~1e-4)\n", worst);
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  const int reps = 10;
  CU(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) syrk_tf32_kernel<<<panels, 128, smem>>>(dP, dG, KV);
  CU(cudaEventRecord(e1));
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  const double flop = 2.0 * kRows * kRows * (double)KV * panels * reps;
  printf("%d panels of 128 x %d: %.3f ms per launch, %.1f TFLOP/s fp32-equivalent (the FP32 pipe peaks at ~72)\n", panels,
         KV, ms / reps, flop / (ms * 1e-3) * 1e-12);
  return worst < 1e-4 ? 0 : 2;
}
