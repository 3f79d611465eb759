// Mask-loss reductions of the training step in two bandwidth passes.  The reference builds the
// weighted BCE and the soft-Jaccard / Tversky terms of loss.py:19-31,164-212 from about a dozen elementwise torch kernels
// over the (B,Q,3,T,H,W) logits and again as many in autograd; here
//   pass 1 (tcow_mask_loss_sums):  S = { sum w*bce(x,y), sum p*y, sum p*(1-y), sum (1-p)*y, sum y },  p = sigmoid(x)
//   pass 2 (tcow_mask_loss_grad):  g = c0 * w*(p - y) + p*(1-p) * (c1*y + c2)
// where the scalars c0..c2 carry the upstream gradients and the Tversky quotient rule (tcow_b200/loss.py).
// Deterministic: per-block partial sums, then a fixed-order second stage.
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {
constexpr int ML_BLOCKS = 1184;  // 8 x 148
constexpr int ML_TERMS = 5;

__device__ __forceinline__ void loss_terms(float x, float y, float w, float (&acc)[ML_TERMS]) {
  const float e = __expf(-fabsf(x));
  const float bce = fmaxf(x, 0.f) - x * y + log1pf(e);          // BCEWithLogits, the stable form torch uses
  const float p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  acc[0] += w * bce;
  acc[1] += p * y;
  acc[2] += p * (1.f - y);
  acc[3] += (1.f - p) * y;
  acc[4] += y;
}
}  // namespace

__global__ void __launch_bounds__(256) mask_loss_sums_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ w, int64_t n4, float* __restrict__ partial) {
  float acc[ML_TERMS] = {0.f, 0.f, 0.f, 0.f, 0.f};
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i);
    const float4 wv = w ? __ldcs(w4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    loss_terms(xv.x, yv.x, wv.x, acc);
    loss_terms(xv.y, yv.y, wv.y, acc);
    loss_terms(xv.z, yv.z, wv.z, acc);
    loss_terms(xv.w, yv.w, wv.w, acc);
  }
  __shared__ float s_red[8][ML_TERMS];
#pragma unroll
  for (int k = 0; k < ML_TERMS; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < ML_TERMS) {
    float s = 0.f;
    for (int wi = 0; wi < 8; ++wi) s += s_red[wi][threadIdx.x];
    partial[blockIdx.x * ML_TERMS + threadIdx.x] = s;
  }
}

__global__ void mask_loss_finalize_kernel(const float* __restrict__ partial, int blocks, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= ML_TERMS) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += static_cast<double>(partial[b * ML_TERMS + k]);
  out[k] = s;
}

__global__ void __launch_bounds__(256) mask_loss_grad_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ w, int64_t n4,
                                                             const float* __restrict__ coef, float* __restrict__ g) {
  const float c0 = __ldg(coef), c1 = __ldg(coef + 1), c2 = __ldg(coef + 2);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* g4 = reinterpret_cast<float4*>(g);
  auto one = [&](float xv, float yv, float wv) {
    const float e = __expf(-fabsf(xv));
    const float p = xv >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    return c0 * wv * (p - yv) + p * (1.f - p) * fmaf(c1, yv, c2);
  };
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i);
    const float4 wv = w ? __ldcs(w4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    __stcs(g4 + i, make_float4(one(xv.x, yv.x, wv.x), one(xv.y, yv.y, wv.y), one(xv.z, yv.z, wv.z), one(xv.w, yv.w, wv.w)));
  }
}

}  // namespace tcow

extern "C" int64_t tcow_mask_loss_workspace_floats(void) { return static_cast<int64_t>(tcow::ML_BLOCKS) * tcow::ML_TERMS; }

extern "C" int tcow_mask_loss_sums(const float* logits, const float* target, const float* weights, int64_t n, float* workspace,
                                   double* sums, void* stream) {
  using namespace tcow;
  if (!logits || !target || !workspace || !sums || n <= 0 || (n % 4)) return set_error(TCOW_ERR_ARG, "mask_loss_sums: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n4 = n / 4;
  int blocks = static_cast<int>((n4 + 255) / 256 < ML_BLOCKS ? (n4 + 255) / 256 : ML_BLOCKS);
  mask_loss_sums_kernel<<<blocks, 256, 0, s>>>(logits, target, weights, n4, workspace);
  mask_loss_finalize_kernel<<<1, 32, 0, s>>>(workspace, blocks, sums);
  return check_launch("mask_loss_sums_kernel");
}

extern "C" int tcow_mask_loss_grad(const float* logits, const float* target, const float* weights, int64_t n, const float* coef,
                                   float* grad, void* stream) {
  using namespace tcow;
  if (!logits || !target || !coef || !grad || n <= 0 || (n % 4)) return set_error(TCOW_ERR_ARG, "mask_loss_grad: bad argument");
  const int64_t n4 = n / 4;
  int blocks = static_cast<int>((n4 + 255) / 256 < ML_BLOCKS * 2 ? (n4 + 255) / 256 : ML_BLOCKS * 2);
  mask_loss_grad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, target, weights, n4, coef, grad);
  return check_launch("mask_loss_grad_kernel");
}
