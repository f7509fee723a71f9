// Does FFMA2 (packed fp32x2) free issue slots on sm_100a?  Interleave FP work with integer ALU work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_ffma2.cu && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, int c) {
  float x[8];
  int y[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {  // 8 scalar FFMA
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
      } else if (MODE == 1) {  // 4 FFMA2 (same flops)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
          x[i] = r.x; x[i + 1] = r.y;
        }
      } else if (MODE == 2) {  // 8 FFMA + 8 integer ops
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = fmaf(x[i], a, b); asm volatile("add.s32 %0, %0, %1;" : "+r"(y[i]) : "r"(c)); }
      } else {  // 4 FFMA2 + 8 integer ops
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float2 r = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
          x[i] = r.x; x[i + 1] = r.y;
          asm volatile("add.s32 %0, %0, %1;" : "+r"(y[i]) : "r"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(y[i + 1]) : "r"(c));
        }
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name) {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int blocks = p.multiProcessorCount * 8, iters = 2048;
  float* d; cudaMalloc(&d, blocks * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f, 12345); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double fma = 64.0 * iters * blocks * 256;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, 2 * fma / (best * 1e-3) * 1e-12);
  cudaFree(d);
}
int main() {
  run<0>("8 FFMA");
  run<1>("4 FFMA2");
  run<2>("8 FFMA + 8 IADD");
  run<3>("4 FFMA2 + 8 IADD");
  return 0;
}
