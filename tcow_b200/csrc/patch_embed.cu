// Patch-embedding front end: im2col gather with the query mask concatenated as the 4th channel, fused with
// the fp32 -> bf16 cast and the optional RGB normalisation, plus the residual-stream initialisation that
// carries conv bias + positional + temporal embeddings.  The contraction itself is the tcgen05 GEMM
// (gemm_tcgen05.cu) accumulating into the stream with the TMA reduce-add epilogue.
// Reference: model/mask_tracker.py:103-108 (cast, clone, cat), model/vision_tf.py:81-89 (normalise),
// vit.py:235-241 (Conv2d k=s=16 as patches), model/vision_tf.py:99-138 (cls/pos/time embeddings).
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

// One thread per 8 consecutive K elements (16 bytes of bf16 out; 32 bytes of fp32 or 8 bytes of uint8 in).
// K index = c*P*P + r*P + w (Conv2d weight layout (D, C, P, P) flattened), P % 8 == 0.
// FT / QT: element type of the frames / the query mask — float, or uint8_t as a video decoder and the data loader leave
// them (data/data_plugin.py:174 divides the integer frames by 255 on the host; `frame_scale` does it here instead).
__device__ __forceinline__ void load8(const float* src, float (&v)[8]) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
  const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const uint8_t* src, float (&v)[8]) {
  const uint2 a = __ldcs(reinterpret_cast<const uint2*>(src));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = static_cast<float>((a.x >> (8 * i)) & 0xffu);
    v[4 + i] = static_cast<float>((a.y >> (8 * i)) & 0xffu);
  }
}

template <typename FT, typename QT>
__global__ void __launch_bounds__(256) patch_gather_kernel(const FT* __restrict__ frames, const QT* __restrict__ query,
                                                           __nv_bfloat16* __restrict__ Pm, int B, int T, int Hf, int Wf,
                                                           int P, int normalize, float frame_scale, int qpv, int sample0) {
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
    float v[8];
    if (c < 3) {
      load8(frames + (((static_cast<long long>(vid) * 3 + c) * T + t) * Hf + y) * Wf + x, v);
      if (frame_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= frame_scale;
      }
      if (normalize) {
        const float m = 0.45f, s = 0.225f;  // vision_tf.py:23-24
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (v[j] - m) / s;
      }
    } else {
      load8(query + ((static_cast<long long>(b) * T + t) * Hf + y) * Wf + x, v);
    }
    reinterpret_cast<uint4*>(Pm)[i] =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
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

extern "C" int tcow_patch_gather_typed(const void* frames, int frames_dtype, const void* query, int query_dtype, void* P,
                                       int B, int T, int Hf, int Wf, int patch, int normalize, float frame_scale,
                                       int queries_per_video, int sample0, void* stream) {
  using namespace tcow;
  if (!frames || !query || !P || B <= 0 || T <= 0 || queries_per_video < 1 || sample0 < 0)
    return set_error(TCOW_ERR_ARG, "patch_gather: bad argument");
  if ((frames_dtype != TCOW_DTYPE_F32 && frames_dtype != TCOW_DTYPE_U8) ||
      (query_dtype != TCOW_DTYPE_F32 && query_dtype != TCOW_DTYPE_U8))
    return set_error(TCOW_ERR_ARG, "patch_gather: dtype must be TCOW_DTYPE_F32 or TCOW_DTYPE_U8");
  if (patch % 8 != 0 || Hf % patch != 0 || Wf % patch != 0)
    return set_error(TCOW_ERR_ARG, "patch_gather: frame %dx%d not divisible by patch %d (or patch %% 8 != 0)", Hf, Wf, patch);
  const uintptr_t fa = frames_dtype == TCOW_DTYPE_F32 ? 15 : 7, qa = query_dtype == TCOW_DTYPE_F32 ? 15 : 7;
  if ((reinterpret_cast<uintptr_t>(frames) & fa) || (reinterpret_cast<uintptr_t>(query) & qa) || (Wf % 8))
    return set_error(TCOW_ERR_ARG, "patch_gather: inputs must be 16-byte (fp32) / 8-byte (uint8) aligned, contiguous, width %% 8 == 0");
  const long long total = static_cast<long long>(B) * (Hf / patch) * (Wf / patch) * T * (4 * patch * patch / 8);
  const int grid = grid_for(total, 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(P);
#define TCOW_PG(FT, QT)                                                                                              \
  patch_gather_kernel<FT, QT><<<grid, 256, 0, st>>>(static_cast<const FT*>(frames), static_cast<const QT*>(query), out, B, \
                                                    T, Hf, Wf, patch, normalize, frame_scale, queries_per_video, sample0)
  if (frames_dtype == TCOW_DTYPE_F32 && query_dtype == TCOW_DTYPE_F32) TCOW_PG(float, float);
  else if (frames_dtype == TCOW_DTYPE_F32) TCOW_PG(float, uint8_t);
  else if (query_dtype == TCOW_DTYPE_F32) TCOW_PG(uint8_t, float);
  else TCOW_PG(uint8_t, uint8_t);
#undef TCOW_PG
  return check_launch("patch_gather_kernel");
}

extern "C" int tcow_patch_gather(const float* frames, const float* query, void* P, int B, int T, int Hf, int Wf,
                                 int patch, int normalize, int queries_per_video, int sample0, void* stream) {
  return tcow_patch_gather_typed(frames, TCOW_DTYPE_F32, query, TCOW_DTYPE_F32, P, B, T, Hf, Wf, patch, normalize, 1.0f,
                                 queries_per_video, sample0, stream);
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
