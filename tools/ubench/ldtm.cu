// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM on sm_100a (bytes per clock as a function of resident warps
// and of the vector width).  Each warp reads its own lane quarter of a 512-column allocation.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I tcow_b200/csrc tools/ubench/ldtm.cu -o build/ubench_ldtm && build/ubench_ldtm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int W>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= r[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= r[i];
  return x;
}
__device__ __forceinline__ void st8(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}

// MODE 0: loads of width W, wait after every `batch` loads.  MODE 1: x8 stores.  MODE 2: load x16 + store x8 alternating.
template <int W, int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters, int batch) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  // each warp sharing a lane quarter walks its own column range
  const int sharers = blockDim.x / 128, me = warp >> 2;
  const int span = 512 / sharers, col0 = me * span;
  uint32_t acc = 0;
  // initialise
  for (int c = 0; c < span; c += 8) st8(base + col0 + c, threadIdx.x);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  int c = 0;
  for (int it = 0; it < iters; ++it) {
    for (int b = 0; b < batch; ++b) {
      if (MODE == 0) acc ^= ld<W>(base + col0 + c);
      if (MODE == 1) st8(base + col0 + c, acc + it);
      if (MODE == 2) { acc ^= ld<W>(base + col0 + c); }
      c += W;
      if (c + W > span) c = 0;
    }
    if (MODE == 0 || MODE == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (MODE == 2) {
      for (int b = 0; b < batch; ++b) st8(base + col0 + ((8 * b) % span), acc);
    }
    if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if (MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int W, int MODE>
void run(const char* name, int threads, int batch) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<W, MODE><<<148, threads>>>(out, cyc, iters, batch);
  k<W, MODE><<<148, threads>>>(out, cyc, iters, batch);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double warps = threads / 32;
  double bytes = 0;
  if (MODE == 0) bytes = warps * batch * W * 128.0;
  if (MODE == 1) bytes = warps * batch * 8 * 128.0;
  if (MODE == 2) bytes = warps * batch * (W + 8) * 128.0;
  printf("%-28s %4d threads/SM batch %d: %8.1f cycles/iter, %7.1f B/clk/SM\n", name, threads, batch, (double)h / iters,
         bytes * iters / (double)h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512}) {
    for (int batch : {1, 2, 4}) {
      run<16, 0>("tcgen05.ld 32x32b.x16", threads, batch);
      run<32, 0>("tcgen05.ld 32x32b.x32", threads, batch);
    }
    run<16, 1>("tcgen05.st 32x32b.x8", threads, 4);
    run<16, 2>("ld x16 + st x8", threads, 2);
  }
  return 0;
}
