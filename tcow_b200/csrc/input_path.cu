// Input path on the device: decoder-layout uint8 video -> the clip tensors the reference's loader
// builds on the host.  Reference: data/data_plugin.py:152-176 (frame selection, `rgb / 255.0`), :178-200 (query / target
// masks, 'T H W C -> C T H W'), data/augs.py:166-210 (centre crop to the aspect ratio, optional flip / crop rectangle,
// torchvision Resize: antialiased bilinear for RGB, nearest for masks).
//   video [F, H, W, C] uint8 (C = 3 RGB as cv2 / imageio decode it, or C = 1 for a mask)
//   out   [C, T, Hf, Wf]   fp32 in [0,1] (bilinear)  or  uint8 (nearest), frame t = video frame  frame_start + t*frame_stride
// The resize reproduces ATen's separable anti-aliased filter (aten/src/ATen/native/cpu/UpSampleKernel.cpp,
// _compute_indices_min_size_weights_aa with the triangle filter, align_corners=False): horizontal pass first, then
// vertical, weights normalised per output index.  Bandwidth-trivial (a clip is 7-27 MB); one thread per output pixel.
#include "tcow_internal.h"

namespace tcow {

constexpr int AA_MAX_TAPS = 35;  // 2*ceil(support)+1 with support = scale <= 17

struct AaAxis {
  float scale, support, invscale;
  int in_size;
};

__host__ __device__ inline AaAxis aa_axis(int in_size, int out_size) {
  AaAxis a;
  a.scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  a.support = a.scale >= 1.0f ? a.scale : 1.0f;  // interp_size (2) * 0.5 * scale
  a.invscale = a.scale >= 1.0f ? 1.0f / a.scale : 1.0f;
  a.in_size = in_size;
  return a;
}

// Taps of output index i along one axis: first source index, count, and normalised weights into w[].
__device__ __forceinline__ void aa_taps(const AaAxis& a, int i, int& xmin, int& xsize, float* w) {
  const float center = static_cast<float>(static_cast<double>(a.scale) * (i + 0.5));
  const long long lo = static_cast<long long>(static_cast<double>(center - a.support) + 0.5);
  const long long hi = static_cast<long long>(static_cast<double>(center + a.support) + 0.5);
  xmin = lo > 0 ? static_cast<int>(lo) : 0;
  xsize = static_cast<int>(hi < a.in_size ? hi : a.in_size) - xmin;
  xsize = xsize < 0 ? 0 : (xsize > AA_MAX_TAPS ? AA_MAX_TAPS : xsize);
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    const float x = static_cast<float>((static_cast<double>(static_cast<float>(j + xmin) - center) + 0.5) * a.invscale);
    const float ax = fabsf(x);
    const float v = ax < 1.0f ? 1.0f - ax : 0.0f;
    w[j] = v;
    total += v;
  }
  if (total != 0.f)
    for (int j = 0; j < xsize; ++j) w[j] /= total;
}

// rect (y0, x0, h, w): the source window after centre crop / crop rectangle; flip mirrors it horizontally first.
template <int C>
__global__ void __launch_bounds__(256) clip_bilinear_kernel(const uint8_t* __restrict__ video, float* __restrict__ out,
                                                            int H, int W, int frame_start, int frame_stride, int T, int y0,
                                                            int x0, int h, int w, int flip, int Hf, int Wf) {
  const AaAxis ay = aa_axis(h, Hf), ax = aa_axis(w, Wf);
  const long long total = static_cast<long long>(T) * Hf * Wf;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % Wf), oy = static_cast<int>((i / Wf) % Hf), t = static_cast<int>(i / (static_cast<long long>(Wf) * Hf));
    float wy[AA_MAX_TAPS], wx[AA_MAX_TAPS];
    int ymin, ysize, xmin, xsize;
    aa_taps(ay, oy, ymin, ysize, wy);
    aa_taps(ax, ox, xmin, xsize, wx);
    const uint8_t* frame = video + static_cast<long long>(frame_start + t * frame_stride) * H * W * C;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int jy = 0; jy < ysize; ++jy) {
      const uint8_t* row = frame + static_cast<long long>(y0 + ymin + jy) * W * C;
      float hsum[C];
#pragma unroll
      for (int c = 0; c < C; ++c) hsum[c] = 0.f;
      for (int jx = 0; jx < xsize; ++jx) {
        const int sx = xmin + jx;
        const uint8_t* px = row + static_cast<long long>(x0 + (flip ? (w - 1 - sx) : sx)) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) hsum[c] = fmaf(wx[jx], __fdiv_rn(static_cast<float>(__ldg(px + c)), 255.0f), hsum[c]);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wy[jy], hsum[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[((static_cast<long long>(c) * T + t) * Hf + oy) * Wf + ox] = acc[c];
  }
}

// torch 'nearest' (legacy): src = min(floor(dst * (float)in/out), in - 1).
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size == out_size) return dst;
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const int s = static_cast<int>(floorf(static_cast<float>(dst) * scale));
  return s < in_size - 1 ? s : in_size - 1;
}

__global__ void __launch_bounds__(256) clip_nearest_kernel(const uint8_t* __restrict__ video, uint8_t* __restrict__ out, int C,
                                                           int H, int W, int frame_start, int frame_stride, int T, int y0,
                                                           int x0, int h, int w, int flip, int Hf, int Wf) {
  const long long total = static_cast<long long>(C) * T * Hf * Wf;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % Wf), oy = static_cast<int>((i / Wf) % Hf);
    const int t = static_cast<int>((i / (static_cast<long long>(Wf) * Hf)) % T), c = static_cast<int>(i / (static_cast<long long>(Wf) * Hf * T));
    const int sy = nearest_src(oy, h, Hf), sx0 = nearest_src(ox, w, Wf);
    const int sx = flip ? (w - 1 - sx0) : sx0;
    out[i] = __ldg(video + ((static_cast<long long>(frame_start + t * frame_stride) * H + (y0 + sy)) * W + (x0 + sx)) * C + c);
  }
}

static int check_clip_args(const void* video, const void* out, int F, int H, int W, int C, int frame_start, int frame_stride,
                           int T, int y0, int x0, int h, int w, int Hf, int Wf) {
  if (!video || !out || F <= 0 || H <= 0 || W <= 0 || T <= 0 || Hf <= 0 || Wf <= 0 || C < 1 || C > 4)
    return set_error(TCOW_ERR_ARG, "clip_from_video: bad argument");
  const long long last = static_cast<long long>(frame_start) + static_cast<long long>(T - 1) * frame_stride;
  if (frame_start < 0 || frame_start >= F || last < 0 || last >= F)
    return set_error(TCOW_ERR_ARG, "clip_from_video: frames %d + t*%d, t < %d leave the video (%d frames)", frame_start,
                     frame_stride, T, F);
  if (y0 < 0 || x0 < 0 || h <= 0 || w <= 0 || y0 + h > H || x0 + w > W)
    return set_error(TCOW_ERR_ARG, "clip_from_video: window (%d,%d,%d,%d) outside the %dx%d frame", y0, x0, h, w, H, W);
  return 0;
}

static int clip_grid(long long total) {
  const long long blocks = (total + 255) / 256, cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace tcow

extern "C" int tcow_clip_from_video_u8(const uint8_t* video, int F, int H, int W, int C, int frame_start, int frame_stride,
                                       int T, int y0, int x0, int h, int w, int flip, int Hf, int Wf, float* out,
                                       void* stream) {
  using namespace tcow;
  if (int rc = check_clip_args(video, out, F, H, W, C, frame_start, frame_stride, T, y0, x0, h, w, Hf, Wf)) return rc;
  if (static_cast<float>(h) / Hf > 17.f || static_cast<float>(w) / Wf > 17.f)
    return set_error(TCOW_ERR_ARG, "clip_from_video: downscale factor above 17 is not supported");
  const int grid = clip_grid(static_cast<long long>(T) * Hf * Wf);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TCOW_CLIP(CC)                                                                                                  \
  clip_bilinear_kernel<CC><<<grid, 256, 0, st>>>(video, out, H, W, frame_start, frame_stride, T, y0, x0, h, w, flip, Hf, Wf)
  switch (C) {
    case 1: TCOW_CLIP(1); break;
    case 2: TCOW_CLIP(2); break;
    case 3: TCOW_CLIP(3); break;
    default: TCOW_CLIP(4); break;
  }
#undef TCOW_CLIP
  return check_launch("clip_bilinear_kernel");
}

extern "C" int tcow_mask_clip_from_video_u8(const uint8_t* video, int F, int H, int W, int C, int frame_start,
                                            int frame_stride, int T, int y0, int x0, int h, int w, int flip, int Hf, int Wf,
                                            uint8_t* out, void* stream) {
  using namespace tcow;
  if (int rc = check_clip_args(video, out, F, H, W, C, frame_start, frame_stride, T, y0, x0, h, w, Hf, Wf)) return rc;
  clip_nearest_kernel<<<clip_grid(static_cast<long long>(C) * T * Hf * Wf), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      video, out, C, H, W, frame_start, frame_stride, T, y0, x0, h, w, flip, Hf, Wf);
  return check_launch("clip_nearest_kernel");
}
