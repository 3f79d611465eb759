// LayerNorm (fp32 statistics) -> bf16, one warp per token row.  HBM-bound: reads 4*D, writes 2*D bytes/row.
// Replaces nn.LayerNorm(eps=1e-6) at vit.py:135 (norm1), :142 (temporal_norm1), :150 (norm2) and the
// optional final norm (vision_tf.py:152-153); gamma == nullptr degenerates to the fp32 -> bf16 cast that
// feeds the head GEMM when norm_embeddings is off.
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

template <int NV>  // D = NV * 128 : each lane owns NV float4
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ y, int rows, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_per_grid) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcs(xr + i * 32 + lane);
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
    if (gamma == nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) yr[i * 32 + lane] = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
      continue;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
      const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
      yr[i * 32 + lane] = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
    }
  }
}

template <int NV>
static int launch_ln(const float* x, const float* g, const float* b, void* y, int rows, float eps, cudaStream_t s) {
  const int warps = 8;
  long long blocks = (static_cast<long long>(rows) + warps - 1) / warps;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  layernorm_bf16_kernel<NV><<<static_cast<int>(blocks), warps * 32, 0, s>>>(x, g, b, static_cast<__nv_bfloat16*>(y), rows, eps);
  return check_launch("layernorm_bf16_kernel");
}

}  // namespace tcow

extern "C" int tcow_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int rows, int D,
                                   float eps, void* stream) {
  using namespace tcow;
  if (!x || !y || rows <= 0) return set_error(TCOW_ERR_ARG, "layernorm: bad pointer or row count");
  if ((gamma == nullptr) != (beta == nullptr)) return set_error(TCOW_ERR_ARG, "layernorm: gamma and beta go together");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 768: return launch_ln<6>(x, gamma, beta, y, rows, eps, s);
    case 896: return launch_ln<7>(x, gamma, beta, y, rows, eps, s);
    case 1024: return launch_ln<8>(x, gamma, beta, y, rows, eps, s);
  }
  return set_error(TCOW_ERR_ARG, "layernorm: unsupported width %d (768, 896 or 1024)", D);
}
