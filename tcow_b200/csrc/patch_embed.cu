// Patch-embedding front end: im2col gather with the query mask concatenated as the 4th channel, fused with
// the fp32 -> bf16 cast and the optional RGB normalisation, plus the residual-stream initialisation that
// carries conv bias + positional + temporal embeddings.  The contraction itself is the tcgen05 GEMM
// (gemm_tcgen05.cu) accumulating into the stream with the TMA reduce-add epilogue.
// Reference: model/mask_tracker.py:103-108 (cast, clone, cat), model/vision_tf.py:81-89 (normalise),
// vit.py:235-241 (Conv2d k=s=16 as patches), model/vision_tf.py:99-138 (cls/pos/time embeddings).
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

// One thread per 8 consecutive K elements (16 bytes of bf16 out, 32 bytes of fp32 in).
// K index = c*P*P + r*P + w (Conv2d weight layout (D, C, P, P) flattened), P % 8 == 0.
__global__ void __launch_bounds__(256) patch_gather_kernel(const float* __restrict__ frames, const float* __restrict__ query,
                                                           __nv_bfloat16* __restrict__ Pm, int B, int T, int Hf, int Wf,
                                                           int P, int normalize, int qpv, int sample0) {
  const int Ho = Hf / P, Wo = Wf / P, N = Ho * Wo;
  const int K = 4 * P * P, KC = K / 8;
  const long long total = static_cast<long long>(B) * N * T * KC;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int kc = static_cast<int>(i % KC);
    const long long row = i / KC;
    const int t = static_cast<int>(row % T);
    const int n = static_cast<int>((row / T) % N);
    const int b = static_cast<int>(row / (static_cast<long long>(T) * N));
    const int k = kc * 8;
    const int c = k / (P * P), r = (k / P) % P, w = k % P;
    const int y = (n / Wo) * P + r, x = (n % Wo) * P + w;
    const int vid = (sample0 + b) / qpv;  // queries of one video share its RGB frames (pipeline.py:134-158)
    const float* src = (c < 3) ? frames + (((static_cast<long long>(vid) * 3 + c) * T + t) * Hf + y) * Wf + x
                               : query + ((static_cast<long long>(b) * T + t) * Hf + y) * Wf + x;
    float4 v0 = __ldcs(reinterpret_cast<const float4*>(src));
    float4 v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
    if (normalize && c < 3) {
      const float m = 0.45f, s = 0.225f;  // vision_tf.py:23-24
      v0.x = (v0.x - m) / s; v0.y = (v0.y - m) / s; v0.z = (v0.z - m) / s; v0.w = (v0.w - m) / s;
      v1.x = (v1.x - m) / s; v1.y = (v1.y - m) / s; v1.z = (v1.z - m) / s; v1.w = (v1.w - m) / s;
    }
    uint4 o = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
    reinterpret_cast<uint4*>(Pm)[i] = o;
  }
}

// X[(b*N+n)*T+t,:] = conv_bias + pos_embed[1+n] + time_embed[t];  X[M+b,:] = cls_token + pos_embed[0].
__global__ void __launch_bounds__(256) embed_init_kernel(float* __restrict__ X, const float* __restrict__ conv_bias,
                                                         const float* __restrict__ pos, const float* __restrict__ tim,
                                                         const float* __restrict__ cls, int B, int N, int T, int D) {
  const int DV = D / 4;
  const long long M = static_cast<long long>(B) * N * T;
  const long long total = (M + B) * DV;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int dv = static_cast<int>(i % DV);
    const long long row = i / DV;
    float4 o;
    if (row < M) {
      const int t = static_cast<int>(row % T);
      const int n = static_cast<int>((row / T) % N);
      const float4 a = __ldg(reinterpret_cast<const float4*>(conv_bias) + dv);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(1 + n) * D) + dv);
      const float4 q = __ldg(reinterpret_cast<const float4*>(tim + static_cast<long long>(t) * D) + dv);
      o = make_float4(a.x + p.x + q.x, a.y + p.y + q.y, a.z + p.z + q.z, a.w + p.w + q.w);
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + dv);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + dv);
      o = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
    reinterpret_cast<float4*>(X)[i] = o;
  }
}

static int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 32;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace tcow

extern "C" int tcow_patch_gather(const float* frames, const float* query, void* P, int B, int T, int Hf, int Wf,
                                 int patch, int normalize, int queries_per_video, int sample0, void* stream) {
  using namespace tcow;
  if (!frames || !query || !P || B <= 0 || T <= 0 || queries_per_video < 1 || sample0 < 0)
    return set_error(TCOW_ERR_ARG, "patch_gather: bad argument");
  if (patch % 8 != 0 || Hf % patch != 0 || Wf % patch != 0)
    return set_error(TCOW_ERR_ARG, "patch_gather: frame %dx%d not divisible by patch %d (or patch %% 8 != 0)", Hf, Wf, patch);
  if ((reinterpret_cast<uintptr_t>(frames) & 15) || (reinterpret_cast<uintptr_t>(query) & 15) || (Wf % 4))
    return set_error(TCOW_ERR_ARG, "patch_gather: inputs must be 16-byte aligned, contiguous, width %% 4 == 0");
  const long long total = static_cast<long long>(B) * (Hf / patch) * (Wf / patch) * T * (4 * patch * patch / 8);
  patch_gather_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, query, static_cast<__nv_bfloat16*>(P), B, T, Hf, Wf, patch, normalize, queries_per_video, sample0);
  return check_launch("patch_gather_kernel");
}

extern "C" int tcow_embed_init(float* X, const float* conv_bias, const float* pos_embed, const float* time_embed,
                               const float* cls_token, int B, int N, int T, int D, void* stream) {
  using namespace tcow;
  if (!X || !conv_bias || !pos_embed || !time_embed || !cls_token || B <= 0 || N <= 0 || T <= 0 || D % 4)
    return set_error(TCOW_ERR_ARG, "embed_init: bad argument");
  const long long total = (static_cast<long long>(B) * N * T + B) * (D / 4);
  embed_init_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, conv_bias, pos_embed, time_embed,
                                                                                      cls_token, B, N, T, D);
  return check_launch("embed_init_kernel");
}
