// The reference's complete mask loss on the device.  Reference: loss.py:12-16 (bootstrapped top-k BCE),
// :19-31 (soft Jaccard / Tversky), :50-53 (BCE or focal), :83-148 (per-pixel weights: class balancing, occluded snitch
// pixels, hard negatives around the target), :164-225 (my_mask_loss: frame selection, weighted BCE, "AOT" mix, scaling).
//
// Layout: a (B,Q,T,H,W) slice of the (B,Q,3,T,H,W) logits / targets is addressed in place as G = B*Q groups of T frames of
// HW pixels: element (g,t,p) at base + g*group_stride + t*HW + p — no .contiguous() copies of the channel slices.
// Everything is a bandwidth pass over (x, y, w) = 12 B/element with cheap math; the top-k mean is an exact 3-level radix
// select on the fp32 bit pattern of the (non-negative) per-pixel loss, so no sort and no host synchronisation:
//   frame_select : sel[f] = any(w != 0 in frame f)                      (which_frames, loss.py:176-185)
//   pass 1       : sums { w*l, p*y, p*(1-y), (1-p)*y, y } over selected frames + histogram of bits [30:19] of l'
//   pass 2 / 3   : histogram of bits [18:7] / [6:0] inside the bin that holds the k-th largest value
//   pass 4       : count / sum of l' > tau and count of l' == tau   (tau = the exact k-th largest value)
//   finalize     : the scalar loss and the gradient coefficients;  grad pass: dL/dx in one sweep.
// l = BCE-with-logits or sigmoid focal loss (alpha 0.25, gamma 2, torchvision.ops.sigmoid_focal_loss); l' = l*w when the
// AOT terms are weighted (occluder / container channels, loss.py:283,301), else l.
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {
namespace {

constexpr int LF_BLOCKS = 592;        // 4 x 148
constexpr int LF_THREADS = 256;
constexpr int LF_BINS1 = 4096, LF_BINS2 = 4096, LF_BINS3 = 128;

// Device-resident state of one loss evaluation (doubles first; see tcow_mask_loss_state_bytes()).
struct LossState {
  double sums[5];          // sum w*l, sum p*y, sum p*(1-y), sum (1-p)*y, sum y       (selected frames)
  double sum_w_all;        // sum of all weights (final_weights.mean() >= 1e-4 test)
  double sum_gt;           // sum of l' > tau
  double loss;             // the scalar result
  double coef[8];          // gradient coefficients written by finalize
  unsigned long long n_sel_frames, N, k, k_rem1, k_rem2, k_rem3, c_gt, c_eq, above1, above2, above3;
  unsigned int bin1, bin2, bin3, tau_bits;
  int active;              // 0: the `else` branch of loss.py:221 (loss = 0, no gradient)
  unsigned long long hist1[LF_BINS1], hist2[LF_BINS2], hist3[LF_BINS3];
};

struct Geo {
  int64_t gs_x, gs_y, gs_w;   // group strides (elements) of logits / target / weights
  int T, hw4;                 // frames per group, float4 quads per frame
  int64_t quads;              // G*T*hw4
};

__device__ __forceinline__ void elem(float x, float y, bool focal, float& l, float& dl, float& p) {
  const float e = __expf(-fabsf(x));
  const float bce = fmaxf(x, 0.f) - x * y + log1pf(e);   // BCEWithLogits, the stable form torch uses
  p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  if (!focal) {
    l = bce;
    dl = p - y;
  } else {  // torchvision sigmoid_focal_loss(alpha=0.25, gamma=2, reduction='none')
    const float pt = p * y + (1.f - p) * (1.f - y);
    const float at = 0.25f * y + 0.75f * (1.f - y);
    const float om = 1.f - pt;
    l = at * om * om * bce;
    const float dpt = p * (1.f - p) * (2.f * y - 1.f);
    dl = at * (om * om * (p - y) - 2.f * om * dpt * bce);
  }
}

template <typename F>
__device__ __forceinline__ void for_each_quad(const Geo& g, const float* x, const float* y, const float* w,
                                              const uint8_t* sel, F&& f) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < g.quads;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t frame = i / g.hw4;
    if (sel && !sel[frame]) continue;
    const int64_t grp = frame / g.T, in_grp = (i - grp * g.T * g.hw4) * 4;
    const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + grp * g.gs_x + in_grp));
    const float4 yv = __ldcs(reinterpret_cast<const float4*>(y + grp * g.gs_y + in_grp));
    const float4 wv = w ? __ldcs(reinterpret_cast<const float4*>(w + grp * g.gs_w + in_grp)) : make_float4(1.f, 1.f, 1.f, 1.f);
    f(i, grp, in_grp, xv, yv, wv);
  }
}

// sel[f] = any(w != 0); accumulates the selected-frame count and the sum of all weights.
__global__ void __launch_bounds__(LF_THREADS) frame_select_kernel(const float* __restrict__ w, int64_t gs_w, int T, int hw4,
                                                                 uint8_t* __restrict__ sel, LossState* st) {
  const int64_t frame = blockIdx.x;
  const int64_t grp = frame / T;
  const float4* p = reinterpret_cast<const float4*>(w + grp * gs_w + (frame - grp * T) * hw4 * 4);
  float s = 0.f;
  int any = 0;
  for (int i = threadIdx.x; i < hw4; i += blockDim.x) {
    const float4 v = __ldg(p + i);
    s += (v.x + v.y) + (v.z + v.w);
    any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
  }
  __shared__ float s_sum[LF_THREADS / 32];
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(&s_any, 1);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < LF_THREADS / 32; ++i) tot += s_sum[i];
    sel[frame] = static_cast<uint8_t>(s_any);
    atomicAdd(&st->sum_w_all, tot);
    if (s_any) atomicAdd(&st->n_sel_frames, 1ULL);
  }
}

template <int LEVEL>
__global__ void __launch_bounds__(LF_THREADS) hist_kernel(Geo g, const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ w, const uint8_t* __restrict__ sel,
                                                         int focal, int weighted_aot, LossState* st) {
  constexpr int BINS = LEVEL == 1 ? LF_BINS1 : (LEVEL == 2 ? LF_BINS2 : LF_BINS3);
  __shared__ unsigned int s_hist[BINS];
  __shared__ double s_red[LF_THREADS / 32][5];
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const unsigned int pre1 = st->bin1, pre2 = st->bin2;   // written by the previous level's select kernel
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for_each_quad(g, x, y, w, sel, [&](int64_t, int64_t, int64_t, float4 xv, float4 yv, float4 wv) {
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float l, dl, p;
      elem(xs[e], ys[e], focal != 0, l, dl, p);
      const unsigned int bits = __float_as_uint(weighted_aot ? l * ws[e] : l) & 0x7fffffffu;
      if (LEVEL == 1) {
        acc[0] += ws[e] * l;
        acc[1] += p * ys[e];
        acc[2] += p * (1.f - ys[e]);
        acc[3] += (1.f - p) * ys[e];
        acc[4] += ys[e];
        atomicAdd(&s_hist[bits >> 19], 1u);
      } else if (LEVEL == 2) {
        if ((bits >> 19) == pre1) atomicAdd(&s_hist[(bits >> 7) & 0xfffu], 1u);
      } else {
        if ((bits >> 19) == pre1 && ((bits >> 7) & 0xfffu) == pre2) atomicAdd(&s_hist[bits & 0x7fu], 1u);
      }
    }
  });
  __syncthreads();
  unsigned long long* gh = LEVEL == 1 ? st->hist1 : (LEVEL == 2 ? st->hist2 : st->hist3);
  for (int i = threadIdx.x; i < BINS; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&gh[i], static_cast<unsigned long long>(s_hist[i]));
  if (LEVEL == 1) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      double v = 0.0;
      for (int i = 0; i < LF_THREADS / 32; ++i) v += s_red[i][threadIdx.x];
      atomicAdd(&st->sums[threadIdx.x], v);
    }
  }
}

// One CTA: walk the histogram of this level from the top bin down to the bin that holds the k-th largest element.
template <int LEVEL>
__global__ void select_kernel(LossState* st, int64_t hw, double topk_frac) {
  if (threadIdx.x != 0) return;
  constexpr int BINS = LEVEL == 1 ? LF_BINS1 : (LEVEL == 2 ? LF_BINS2 : LF_BINS3);
  const unsigned long long* h = LEVEL == 1 ? st->hist1 : (LEVEL == 2 ? st->hist2 : st->hist3);
  unsigned long long want;
  if (LEVEL == 1) {
    st->N = st->n_sel_frames * static_cast<unsigned long long>(hw);
    st->k = static_cast<unsigned long long>(topk_frac * static_cast<double>(st->N));   // int(topk_frac * numel), loss.py:13
    if (st->k < 1) st->k = 1;
    want = st->k;
  } else {
    want = LEVEL == 2 ? st->k_rem1 : st->k_rem2;
  }
  unsigned long long above = 0;
  int b = BINS - 1;
  for (; b > 0; --b) {
    if (above + h[b] >= want) break;
    above += h[b];
  }
  if (LEVEL == 1) { st->bin1 = b; st->above1 = above; st->k_rem1 = want - above; }
  if (LEVEL == 2) { st->bin2 = b; st->above2 = above; st->k_rem2 = want - above; }
  if (LEVEL == 3) {
    st->bin3 = b; st->above3 = above; st->k_rem3 = want - above;
    st->tau_bits = (st->bin1 << 19) | (st->bin2 << 7) | static_cast<unsigned int>(b);
  }
}

// count / sum of l' > tau, count of l' == tau
__global__ void __launch_bounds__(LF_THREADS) above_kernel(Geo g, const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ w, const uint8_t* __restrict__ sel,
                                                          int focal, int weighted_aot, LossState* st) {
  const unsigned int tau = st->tau_bits;
  double s = 0.0;
  unsigned int cg = 0, ce = 0;
  for_each_quad(g, x, y, w, sel, [&](int64_t, int64_t, int64_t, float4 xv, float4 yv, float4 wv) {
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float l, dl, p;
      elem(xs[e], ys[e], focal != 0, l, dl, p);
      const float la = weighted_aot ? l * ws[e] : l;
      const unsigned int bits = __float_as_uint(la) & 0x7fffffffu;
      if (bits > tau) { s += la; ++cg; }
      ce += bits == tau;
    }
  });
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    cg += __shfl_xor_sync(0xffffffffu, cg, o);
    ce += __shfl_xor_sync(0xffffffffu, ce, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&st->sum_gt, s);
    atomicAdd(&st->c_gt, static_cast<unsigned long long>(cg));
    atomicAdd(&st->c_eq, static_cast<unsigned long long>(ce));
  }
}

// loss.py:187-223 in scalars.  coef: [0] custom weight / N, [1] bootstrap weight / k, [2] tie fraction,
// [3],[4] Tversky c1, c2 (dL/dx += p(1-p)(c1*y + c2)), all still to be multiplied by the upstream gradient.
__global__ void finalize_kernel(LossState* st, int64_t total_frames, int64_t hw, double aot, int weighted_aot, double alpha,
                                double beta, double eps, float* loss_out) {
  if (threadIdx.x != 0) return;
  const double N = static_cast<double>(st->N);
  const double n_all = static_cast<double>(total_frames) * static_cast<double>(hw);
  st->active = (st->n_sel_frames > 0 && st->sum_w_all / n_all >= 1e-4) ? 1 : 0;     // loss.py:187
  for (int i = 0; i < 8; ++i) st->coef[i] = 0.0;
  if (!st->active) {
    st->loss = 0.0;
    *loss_out = 0.f;
    return;
  }
  const double custom = st->sums[0] / N;
  const double scale = sqrt(static_cast<double>(st->n_sel_frames) / static_cast<double>(total_frames));   // loss.py:219
  double loss = custom;
  st->coef[0] = scale / N;
  if (aot > 0.0) {
    const double k = static_cast<double>(st->k);
    const float tau = __uint_as_float(st->tau_bits);
    const double boot = (st->sum_gt + (k - static_cast<double>(st->c_gt)) * static_cast<double>(tau)) / k;
    double jac = boot, boot_share = 1.0;                       // weighted AOT: loss_jaccard = loss_bootstrap (loss.py:205)
    if (!weighted_aot) {
      boot_share = 0.5;
      const double num = st->sums[1];
      const double den = num + alpha * st->sums[2] + beta * st->sums[3] + eps;
      const bool has_target = st->sums[4] / N >= 1e-6;         // loss.py:20
      jac = has_target ? 1.0 - num / den : 0.0;
      if (has_target) {
        // L = 1 - num/den; d num = p(1-p) y; d den = p(1-p) (y + alpha (1-y) - beta y)
        st->coef[3] = scale * aot * 0.5 * (-1.0 / den + num / (den * den) * (1.0 - alpha - beta));
        st->coef[4] = scale * aot * 0.5 * (num / (den * den) * alpha);
      }
    }
    loss = ((boot + jac) / 2.0) * aot + custom * (1.0 - aot);
    st->coef[0] = scale * (1.0 - aot) / N;
    st->coef[1] = scale * aot * boot_share / k;
    st->coef[2] = st->c_eq ? (k - static_cast<double>(st->c_gt)) / static_cast<double>(st->c_eq) : 0.0;
  }
  st->loss = loss * scale;
  *loss_out = static_cast<float>(st->loss);
}

__global__ void __launch_bounds__(LF_THREADS) grad_kernel(Geo g, const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ w, const uint8_t* __restrict__ sel,
                                                         int focal, int weighted_aot, const LossState* __restrict__ st,
                                                         const float* __restrict__ upstream, float* __restrict__ grad,
                                                         int64_t gs_g, int accumulate) {
  const float up = upstream ? __ldg(upstream) : 1.f;
  const float c0 = static_cast<float>(st->coef[0]) * up, c1 = static_cast<float>(st->coef[1]) * up;
  const float tie = static_cast<float>(st->coef[2]);
  const float t1 = static_cast<float>(st->coef[3]) * up, t2 = static_cast<float>(st->coef[4]) * up;
  const unsigned int tau = st->tau_bits;
  const int active = st->active;
  // unselected frames get an explicit zero gradient (they are skipped by the forward passes), so no `sel` filter here
  for_each_quad(g, x, y, w, nullptr, [&](int64_t i, int64_t grp, int64_t in_grp, float4 xv, float4 yv, float4 wv) {
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
    float o[4];
    const bool on = active && sel[i / g.hw4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float l, dl, p;
      elem(xs[e], ys[e], focal != 0, l, dl, p);
      const float la = weighted_aot ? l * ws[e] : l;
      const unsigned int bits = __float_as_uint(la) & 0x7fffffffu;
      const float pick = bits > tau ? 1.f : (bits == tau ? tie : 0.f);
      const float v = c0 * ws[e] * dl + c1 * pick * (weighted_aot ? ws[e] : 1.f) * dl + p * (1.f - p) * fmaf(t1, ys[e], t2);
      o[e] = on ? v : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(grad + grp * gs_g + in_grp);
    if (accumulate) {
      const float4 old = *dst;
      *dst = make_float4(old.x + o[0], old.y + o[1], old.z + o[2], old.w + o[3]);
    } else {
      *dst = make_float4(o[0], o[1], o[2], o[3]);
    }
  });
}

// ------------------------------------------------------------------------------------------------ per-pixel weights
__device__ __forceinline__ int reflect(int i, int n) {   // torch 'reflect' padding (no edge repeat)
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// tmp[f,y,x] = any(target[f,y,x'] > 0 for |x' - x| <= r, reflected)
__global__ void __launch_bounds__(256) dilate_rows_kernel(const float* __restrict__ tgt, int64_t gs_t, int T, int H, int W, int r,
                                                         uint8_t* __restrict__ tmp, int64_t total) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int64_t row = i / W;
    const int64_t frame = row / H;
    const int64_t grp = frame / T;
    const float* src = tgt + grp * gs_t + (row - grp * T * H) * W;
    int any = 0;
    for (int d = -r; d <= r && !any; ++d) any = __ldg(src + reflect(x + d, W)) > 0.f;
    tmp[i] = static_cast<uint8_t>(any);
  }
}

// w = frame_w[f] * (class-balance factor) * (2 if the snitch pixel is occluded) * (hard-negative factor near the target)
__global__ void __launch_bounds__(256) pixel_weights_kernel(const float* __restrict__ tgt, int64_t gs_t, const uint8_t* __restrict__ occl,
                                                           int64_t gs_o, const float* __restrict__ frame_w,
                                                           const float* __restrict__ corr, const uint8_t* __restrict__ tmp,
                                                           int T, int H, int W, int r, float hn_factor,
                                                           float* __restrict__ out, int64_t total) {
  const float pos_corr = corr ? __ldg(corr) : 1.f, neg_corr = corr ? __ldg(corr + 1) : 1.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int yy = static_cast<int>((i / W) % H);
    const int64_t frame = i / (static_cast<int64_t>(W) * H);
    const int64_t grp = frame / T;
    const int64_t in_grp = i - grp * static_cast<int64_t>(T) * H * W;
    const float t = __ldg(tgt + grp * gs_t + in_grp);
    float wgt = 1.f;
    if (t == 0.f) wgt *= neg_corr;       // loss.py:124-125 (exact comparisons, as the reference's boolean masks)
    if (t == 1.f) wgt *= pos_corr;
    if (occl && __ldg(occl + grp * gs_o + in_grp) != 0) wgt *= 2.f;      // loss.py:128-129
    if (hn_factor > 1.f && t < 0.5f) {                                   // loss.py:133-146
      const uint8_t* col = tmp + frame * static_cast<int64_t>(H) * W + x;
      int any = 0;
      for (int d = -r; d <= r && !any; ++d) any = col[static_cast<int64_t>(reflect(yy + d, H)) * W];
      if (any) wgt *= hn_factor;
    }
    out[i] = wgt * (frame_w ? __ldg(frame_w + frame) : 1.f);
  }
}

// counts[0] = #(target == 1), counts[1] = #(target == 0)   (class balancing, loss.py:103-107)
__global__ void __launch_bounds__(256) class_count_kernel(const float* __restrict__ tgt, int64_t gs_t, int T, int hw4, int64_t quads,
                                                         unsigned long long* counts) {
  unsigned int pos = 0, neg = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < quads;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t frame = i / hw4, grp = frame / T;
    const float4 v = __ldg(reinterpret_cast<const float4*>(tgt + grp * gs_t + (i - grp * T * hw4) * 4));
    pos += (v.x == 1.f) + (v.y == 1.f) + (v.z == 1.f) + (v.w == 1.f);
    neg += (v.x == 0.f) + (v.y == 0.f) + (v.z == 0.f) + (v.w == 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o);
    neg += __shfl_xor_sync(0xffffffffu, neg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[0], static_cast<unsigned long long>(pos));
    atomicAdd(&counts[1], static_cast<unsigned long long>(neg));
  }
}

int grid_for_quads(int64_t quads) {
  const int64_t b = (quads + LF_THREADS - 1) / LF_THREADS;
  return static_cast<int>(b < LF_BLOCKS ? b : LF_BLOCKS);
}

}  // namespace
}  // namespace tcow

extern "C" int64_t tcow_mask_loss_state_bytes(void) { return static_cast<int64_t>(sizeof(tcow::LossState)); }

extern "C" int tcow_mask_loss_forward(const float* logits, int64_t gs_x, const float* target, int64_t gs_y, const float* weights,
                                      int64_t gs_w, int G, int T, int64_t hw, int focal, int weighted_aot, double aot_loss,
                                      double topk_frac, double alpha, double beta, double eps, uint8_t* frame_sel, void* state,
                                      float* loss_out, void* stream) {
  using namespace tcow;
  if (!logits || !target || !weights || !frame_sel || !state || !loss_out || G <= 0 || T <= 0 || hw <= 0 || (hw % 4))
    return set_error(TCOW_ERR_ARG, "mask_loss_forward: bad argument (hw must be a multiple of 4)");
  if ((gs_x % 4) || (gs_y % 4) || (gs_w % 4) || ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target) |
                                                  reinterpret_cast<uintptr_t>(weights)) & 15))
    return set_error(TCOW_ERR_ARG, "mask_loss_forward: tensors must be 16-byte aligned with group strides %% 4 == 0");
  if (!(topk_frac > 0.0 && topk_frac <= 1.0) || aot_loss < 0.0 || aot_loss > 1.0)
    return set_error(TCOW_ERR_ARG, "mask_loss_forward: topk_frac in (0,1], aot_loss in [0,1]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LossState* st = static_cast<LossState*>(state);
  const int64_t frames = static_cast<int64_t>(G) * T;
  const Geo g{gs_x, gs_y, gs_w, T, static_cast<int>(hw / 4), frames * (hw / 4)};
  const int grid = grid_for_quads(g.quads);
  cudaMemsetAsync(st, 0, sizeof(LossState), s);
  frame_select_kernel<<<static_cast<unsigned>(frames), LF_THREADS, 0, s>>>(weights, gs_w, T, g.hw4, frame_sel, st);
  hist_kernel<1><<<grid, LF_THREADS, 0, s>>>(g, logits, target, weights, frame_sel, focal, weighted_aot, st);
  if (aot_loss > 0.0) {
    select_kernel<1><<<1, 32, 0, s>>>(st, hw, topk_frac);
    hist_kernel<2><<<grid, LF_THREADS, 0, s>>>(g, logits, target, weights, frame_sel, focal, weighted_aot, st);
    select_kernel<2><<<1, 32, 0, s>>>(st, hw, topk_frac);
    hist_kernel<3><<<grid, LF_THREADS, 0, s>>>(g, logits, target, weights, frame_sel, focal, weighted_aot, st);
    select_kernel<3><<<1, 32, 0, s>>>(st, hw, topk_frac);
    above_kernel<<<grid, LF_THREADS, 0, s>>>(g, logits, target, weights, frame_sel, focal, weighted_aot, st);
  } else {
    select_kernel<1><<<1, 32, 0, s>>>(st, hw, 1.0);   // N only
  }
  finalize_kernel<<<1, 32, 0, s>>>(st, frames, hw, aot_loss, weighted_aot, alpha, beta, eps, loss_out);
  return check_launch("mask_loss_forward");
}

extern "C" int tcow_mask_loss_backward(const float* logits, int64_t gs_x, const float* target, int64_t gs_y, const float* weights,
                                       int64_t gs_w, int G, int T, int64_t hw, int focal, int weighted_aot,
                                       const uint8_t* frame_sel, const void* state, const float* upstream, float* grad,
                                       int64_t gs_g, int accumulate, void* stream) {
  using namespace tcow;
  if (!logits || !target || !weights || !frame_sel || !state || !grad || G <= 0 || T <= 0 || hw <= 0 || (hw % 4) || (gs_g % 4))
    return set_error(TCOW_ERR_ARG, "mask_loss_backward: bad argument");
  const int64_t frames = static_cast<int64_t>(G) * T;
  const Geo g{gs_x, gs_y, gs_w, T, static_cast<int>(hw / 4), frames * (hw / 4)};
  grad_kernel<<<grid_for_quads(g.quads) * 2, LF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      g, logits, target, weights, frame_sel, focal, weighted_aot, static_cast<const LossState*>(state), upstream, grad, gs_g,
      accumulate);
  return check_launch("mask_loss_grad_kernel");
}

extern "C" int tcow_loss_class_counts(const float* target, int64_t gs_t, int G, int T, int64_t hw, uint64_t* counts, void* stream) {
  using namespace tcow;
  if (!target || !counts || G <= 0 || T <= 0 || hw <= 0 || (hw % 4) || (gs_t % 4))
    return set_error(TCOW_ERR_ARG, "loss_class_counts: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t quads = static_cast<int64_t>(G) * T * (hw / 4);
  cudaMemsetAsync(counts, 0, 2 * sizeof(uint64_t), s);
  class_count_kernel<<<grid_for_quads(quads), 256, 0, s>>>(target, gs_t, T, static_cast<int>(hw / 4), quads,
                                                           reinterpret_cast<unsigned long long*>(counts));
  return check_launch("class_count_kernel");
}

extern "C" int tcow_loss_pixel_weights(const float* target, int64_t gs_t, const uint8_t* occl, int64_t gs_o, const float* frame_w,
                                       const float* corr, int G, int T, int H, int W, float hard_negative_factor, int band,
                                       uint8_t* tmp, float* out, void* stream) {
  using namespace tcow;
  if (!target || !out || G <= 0 || T <= 0 || H <= 0 || W <= 0 || (hard_negative_factor > 1.f && (!tmp || band < 1 || !(band & 1))))
    return set_error(TCOW_ERR_ARG, "loss_pixel_weights: bad argument (band must be odd, tmp needed for hard negatives)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(G) * T * H * W;
  const int64_t blocks = (total + 255) / 256, cap = static_cast<int64_t>(sm_count()) * 16;
  const int grid = static_cast<int>(blocks < cap ? blocks : cap);
  const int r = band / 2;
  if (hard_negative_factor > 1.f) dilate_rows_kernel<<<grid, 256, 0, s>>>(target, gs_t, T, H, W, r, tmp, total);
  pixel_weights_kernel<<<grid, 256, 0, s>>>(target, gs_t, occl, gs_o, frame_w, corr, tmp, T, H, W, r, hard_negative_factor,
                                            out, total);
  return check_launch("pixel_weights_kernel");
}
