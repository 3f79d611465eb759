// Mask-decoding head tail: assemble the avg-pooled per-token patches into the low-resolution map and upsample
// it (bilinear align_corners=True, or nearest) into fp32 logits; plus the per-frame flag mean.
// Reference: model/mask_tracker.py:114-115 (pixel-shuffle rearrange), :118-132 (avg_pool2d + interpolate),
// :135-137 (flag linear + spatial mean).  The avg-pool is folded into the head GEMM weights on the host
// (tcow_b200/engine.py), so `low` already holds the pooled C x pp x pp values per token.
// Write-bound: 4*C*Hf*Wf bytes per (clip, frame); one float4 store per thread, reads hit L2.
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

// One CTA per (clip, channel, frame) image: the (Ho*pp) x (Wo*pp) low-resolution map is first assembled in shared
// memory (pp contiguous floats per gather), then every thread produces float4 runs of the full-resolution row.
__global__ void __launch_bounds__(256) mask_upsample_kernel(const float* __restrict__ low, int64_t ld_low,
                                                            float* __restrict__ out, int B, int T, int Ho, int Wo,
                                                            int C, int pp, int stride, int mode) {
  extern __shared__ float s_img[];
  const int Hl = Ho * pp, Wl = Wo * pp;
  const int Hf = Hl * stride, Wf = Wl * stride;
  const int N = Ho * Wo;
  const int img = blockIdx.x;  // ((b*C + c)*T + t)
  const int t = img % T, c = (img / T) % C, b = img / (T * C);
  for (int i = threadIdx.x; i < Hl * Wo; i += blockDim.x) {
    const int ly = i / Wo, pw = i % Wo;
    const int n = (ly / pp) * Wo + pw;
    const float* src = low + ((static_cast<int64_t>(b) * N + n) * T + t) * ld_low + (c * pp + (ly % pp)) * pp;
    float* dst = s_img + ly * Wl + pw * pp;
    if (pp == 4) {
      *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      for (int j = 0; j < pp; ++j) dst[j] = __ldg(src + j);
    }
  }
  __syncthreads();
  // PyTorch area_pixel_compute_scale(align_corners=True): (in-1)/(out-1), 0 when out == 1.
  const float sy = (Hf > 1) ? static_cast<float>(Hl - 1) / static_cast<float>(Hf - 1) : 0.f;
  const float sx = (Wf > 1) ? static_cast<float>(Wl - 1) / static_cast<float>(Wf - 1) : 0.f;
  const int WV = Wf / 4;
  float4* dst_img = reinterpret_cast<float4*>(out + static_cast<int64_t>(img) * Hf * Wf);
  for (int i = threadIdx.x; i < Hf * WV; i += blockDim.x) {
    const int y = i / WV, xv = i % WV;
    float o[4];
    if (mode == 1 || stride == 1) {  // nearest (scale_factor = stride): src = dst / stride
      const float* r = s_img + (y / stride) * Wl;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = r[(xv * 4 + e) / stride];
    } else {
      const float fy = sy * static_cast<float>(y);
      const int y0 = static_cast<int>(fy);
      const int y1 = y0 + ((y0 < Hl - 1) ? 1 : 0);
      const float wy1 = fy - static_cast<float>(y0), wy0 = 1.f - wy1;
      const float* r0 = s_img + y0 * Wl;
      const float* r1 = s_img + y1 * Wl;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float fx = sx * static_cast<float>(xv * 4 + e);
        const int x0 = static_cast<int>(fx);
        const int x1 = x0 + ((x0 < Wl - 1) ? 1 : 0);
        const float wx1 = fx - static_cast<float>(x0), wx0 = 1.f - wx1;
        o[e] = wy0 * (wx0 * r0[x0] + wx1 * r0[x1]) + wy1 * (wx0 * r1[x0] + wx1 * r1[x1]);
      }
    }
    __stcs(dst_img + i, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// flags[b,t,f] = mean_n low[(b*N+n)*T+t, col0+f]; one warp per (b,t).
__global__ void __launch_bounds__(128) flag_mean_kernel(const float* __restrict__ low, int64_t ld_low, float* __restrict__ flags,
                                                        int B, int N, int T, int F, int col0) {
  const int lane = threadIdx.x & 31;
  const int bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bt >= B * T) return;
  const int b = bt / T, t = bt % T;
  for (int f = 0; f < F; ++f) {
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s += __ldg(low + ((static_cast<int64_t>(b) * N + n) * T + t) * ld_low + col0 + f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) flags[static_cast<int64_t>(bt) * F + f] = s / static_cast<float>(N);
  }
}

// Per-image mask areas for the IoU metrics (eval/metrics.py:18-41): with pred = logit > 0 and gt = target > 0.5,
// areas[img] = (|gt|, |pred & gt|, |pred | gt|); one CTA per (b, c, t) image, one pass over logits and targets.
__global__ void __launch_bounds__(256) mask_iou_areas_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                             float* __restrict__ areas, int hw) {
  const int img = blockIdx.x;
  const float4* l4 = reinterpret_cast<const float4*>(logits + static_cast<int64_t>(img) * hw);
  const float4* t4 = reinterpret_cast<const float4*>(target + static_cast<int64_t>(img) * hw);
  int gt = 0, inter = 0, uni = 0;
  for (int i = threadIdx.x; i < hw / 4; i += blockDim.x) {
    const float4 l = __ldcs(l4 + i), t = __ldcs(t4 + i);
    const float lv[4] = {l.x, l.y, l.z, l.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool p = lv[e] > 0.f, g = tv[e] > 0.5f;
      gt += g; inter += (p && g); uni += (p || g);
    }
  }
  __shared__ int s_red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gt += __shfl_xor_sync(0xffffffffu, gt, o);
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    uni += __shfl_xor_sync(0xffffffffu, uni, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = gt; s_red[1][threadIdx.x >> 5] = inter; s_red[2][threadIdx.x >> 5] = uni;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += s_red[threadIdx.x][w];
    areas[static_cast<int64_t>(img) * 3 + threadIdx.x] = static_cast<float>(s);
  }
}

}  // namespace tcow

extern "C" int tcow_mask_iou_areas(const float* logits, const float* target, float* areas, int images, int hw,
                                   void* stream) {
  using namespace tcow;
  if (!logits || !target || !areas || images <= 0 || hw <= 0 || (hw % 4)) return set_error(TCOW_ERR_ARG, "mask_iou_areas: bad argument");
  mask_iou_areas_kernel<<<images, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, target, areas, hw);
  return check_launch("mask_iou_areas_kernel");
}

extern "C" int tcow_mask_upsample(const float* low, int64_t ld_low, float* out, int B, int T, int Ho, int Wo, int C,
                                  int pp, int stride, int mode, void* stream) {
  using namespace tcow;
  if (!low || !out || B <= 0 || T <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || pp <= 0 || stride <= 0)
    return set_error(TCOW_ERR_ARG, "mask_upsample: bad argument");
  if ((Wo * pp * stride) % 4) return set_error(TCOW_ERR_ARG, "mask_upsample: frame width must be a multiple of 4");
  if (mode != 0 && mode != 1) return set_error(TCOW_ERR_ARG, "mask_upsample: mode must be 0 (bilinear) or 1 (nearest)");
  if ((pp == 4) && ((ld_low % 4) || (reinterpret_cast<uintptr_t>(low) & 15)))
    return set_error(TCOW_ERR_ARG, "mask_upsample: low must be 16-byte aligned with a row pitch multiple of 4");
  const size_t smem = static_cast<size_t>(Ho) * pp * Wo * pp * sizeof(float);
  if (smem > 200 * 1024) return set_error(TCOW_ERR_ARG, "mask_upsample: low-resolution map of %zu bytes exceeds shared memory", smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mask_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  mask_upsample_kernel<<<B * C * T, 256, smem, static_cast<cudaStream_t>(stream)>>>(low, ld_low, out, B, T, Ho, Wo, C, pp,
                                                                               stride, mode);
  return check_launch("mask_upsample_kernel");
}

extern "C" int tcow_flag_mean(const float* low, int64_t ld_low, float* flags, int B, int N, int T, int F, int col0,
                              void* stream) {
  using namespace tcow;
  if (!low || !flags || B <= 0 || N <= 0 || T <= 0 || F <= 0) return set_error(TCOW_ERR_ARG, "flag_mean: bad argument");
  const int bt = B * T;
  flag_mean_kernel<<<(bt + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(low, ld_low, flags, B, N, T, F, col0);
  return check_launch("flag_mean_kernel");
}
