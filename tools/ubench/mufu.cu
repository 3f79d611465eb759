// Micro-benchmark: issue rate of MUFU.EX2 / F2FP.BF16 pack / FFMA2 per SM sub-partition on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/mufu.cu -o build/ubench_mufu && build/ubench_mufu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 1 || MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2(v[i]);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(v[i], v[i + 1]);
    }
    if (MODE == 3) {   // bf16 pack with integer ops instead of F2FP
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned a = __float_as_uint(v[i]), b = __float_as_uint(v[i + 1]);
        a += 0x7fffu + ((a >> 16) & 1u);
        b += 0x7fffu + ((b >> 16) & 1u);
        acc ^= __byte_perm(a, b, 0x7632);
      }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += 1.0f;
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  printf("%-34s %4d threads/SM: %7.2f cycles per 16 elements per warp, %6.2f cycles per warp-element per SMSP\n", name, threads,
         (double)h / iters, (double)h / iters / 16 / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("16 MUFU.EX2", threads);
    run<1>("16 MUFU.EX2 + 8 F2FP pack", threads);
    run<2>("8 F2FP pack + 16 FADD", threads);
    run<3>("16 MUFU.EX2 + integer bf16 pack", threads);
  }
  return 0;
}
