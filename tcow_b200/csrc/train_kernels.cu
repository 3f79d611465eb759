// Bandwidth kernels of the training step (forward-with-save and backward of the non-GEMM operators).
//   LayerNorm fwd that also saves xhat (bf16) and rstd   — nn.LayerNorm at vit.py:135,142,150 in training mode
//   LayerNorm bwd fused with the residual-gradient accumulate and its bf16 copy (the next GEMM's operand)
//   column sums (bias gradients of every nn.Linear), embedding gradients (vision_tf.py:99-138),
//   the adjoint of the mask head tail (mask_tracker.py:114-137: pixel shuffle + avg-pool + bilinear upsample + flags)
// All reductions are deterministic: per-block partial sums in a caller-provided workspace, then a finalize pass.
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {
constexpr int LNB_MAX_BLOCKS = 592;   // 4 x 148
constexpr int CS_MAX_CHUNKS = 256;

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
}  // namespace

// ---------------------------------------------------------------------------------------------- LayerNorm, training
template <int NV>  // D = NV * 128
__global__ void __launch_bounds__(256) layernorm_train_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                              __nv_bfloat16* __restrict__ xhat, float* __restrict__ rstd_out,
                                                              int rows, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps_per_grid) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldcs(xr + i * 32 + lane);
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
    if (lane == 0) rstd_out[row] = rstd;
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
    uint2* hr = reinterpret_cast<uint2*>(xhat + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
      const float h0 = (v[i].x - mean) * rstd, h1 = (v[i].y - mean) * rstd, h2 = (v[i].z - mean) * rstd,
                  h3 = (v[i].w - mean) * rstd;
      hr[i * 32 + lane] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
      yr[i * 32 + lane] = make_uint2(pack_bf16(h0 * g.x + b.x, h1 * g.y + b.y), pack_bf16(h2 * g.z + b.z, h3 * g.w + b.w));
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   G (+)= dx;  Gb = bf16(G)
// per-block partial sums of dgamma = sum_rows dy*xhat and dbeta = sum_rows dy -> partial[block][2][D]
// One warp per row; lane l owns the 8-column groups (v*32+l)*8 .. +8 (16-byte bf16 vectors, 2 x float4 of G).  dy / xhat
// stay packed in registers and the G row is requested together with them, so that all of a row's HBM latency is paid
// once; the kernel is limited to 128 registers (two 8-warp blocks per SM).
template <int NV>  // D = NV * 256
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ xhat,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               float* __restrict__ G, __nv_bfloat16* __restrict__ Gb,
                                                               float* __restrict__ partial, int rows, int accumulate,
                                                               const float* __restrict__ next_scale,
                                                               __nv_bfloat16* __restrict__ Gs) {
  constexpr int D = NV * 256;
  __shared__ float s_red[8][2][256];  // one 256-column group (vector index) at a time
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_grid = gridDim.x * 8;
  float dg[NV][8], db[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[v][e] = db[v][e] = 0.f;
  const float4* gm4 = reinterpret_cast<const float4*>(gamma);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += warps_per_grid) {
    const uint4* dr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * D);
    const uint4* hr = reinterpret_cast<const uint4*>(xhat + static_cast<size_t>(row) * D);
    float4* gr = reinterpret_cast<float4*>(G + static_cast<size_t>(row) * D);
    uint4 a[NV], b[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      a[v] = __ldcs(dr + v * 32 + lane);
      b[v] = __ldcs(hr + v * 32 + lane);
    }
    if (accumulate) {  // pull the G row towards L2 now; it is read after the row statistics
#pragma unroll
      for (int v = 0; v < NV; ++v)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gr + (v * 32 + lane) * 2));
    }
    const float rs = __ldg(rstd + row);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint32_t aw[4] = {a[v].x, a[v].y, a[v].z, a[v].w}, bw[4] = {b[v].x, b[v].y, b[v].z, b[v].w};
      const float4 g0 = __ldg(gm4 + (v * 32 + lane) * 2), g1 = __ldg(gm4 + (v * 32 + lane) * 2 + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = bf_lo(aw[e]), d1 = bf_hi(aw[e]), h0 = bf_lo(bw[e]), h1 = bf_hi(bw[e]);
        dg[v][2 * e] = fmaf(d0, h0, dg[v][2 * e]);
        dg[v][2 * e + 1] = fmaf(d1, h1, dg[v][2 * e + 1]);
        db[v][2 * e] += d0;
        db[v][2 * e + 1] += d1;
        const float q0 = d0 * gg[2 * e], q1 = d1 * gg[2 * e + 1];  // g = dy * gamma
        s1 += q0 + q1;
        s2 = fmaf(q0, h0, fmaf(q1, h1, s2));
      }
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o2);
    }
    const float c1 = s1 * (1.0f / D), c2 = s2 * (1.0f / D);
    uint4* br = reinterpret_cast<uint4*>(Gb + static_cast<size_t>(row) * D);
    // stochastic depth: the next branch consumes scale[row] * G as well (its dY operand) — written here instead of by a
    // separate pass over Gb
    uint4* sr = Gs ? reinterpret_cast<uint4*>(Gs + static_cast<size_t>(row) * D) : nullptr;
    const float nsc = Gs ? __ldg(next_scale + row) : 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint32_t aw[4] = {a[v].x, a[v].y, a[v].z, a[v].w}, bw[4] = {b[v].x, b[v].y, b[v].z, b[v].w};
      const float4 g0 = __ldg(gm4 + (v * 32 + lane) * 2), g1 = __ldg(gm4 + (v * 32 + lane) * 2 + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float r[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h0 = bf_lo(bw[e]), h1 = bf_hi(bw[e]);
        r[2 * e] = rs * (bf_lo(aw[e]) * gg[2 * e] - c1 - h0 * c2);
        r[2 * e + 1] = rs * (bf_hi(aw[e]) * gg[2 * e + 1] - c1 - h1 * c2);
      }
      float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
      if (accumulate) {
        o0 = gr[(v * 32 + lane) * 2];
        o1 = gr[(v * 32 + lane) * 2 + 1];
      }
      o0.x += r[0]; o0.y += r[1]; o0.z += r[2]; o0.w += r[3];
      o1.x += r[4]; o1.y += r[5]; o1.z += r[6]; o1.w += r[7];
      gr[(v * 32 + lane) * 2] = o0;
      gr[(v * 32 + lane) * 2 + 1] = o1;
      br[v * 32 + lane] = make_uint4(pack_bf16(o0.x, o0.y), pack_bf16(o0.z, o0.w), pack_bf16(o1.x, o1.y), pack_bf16(o1.z, o1.w));
      if (sr)
        sr[v * 32 + lane] = make_uint4(pack_bf16(nsc * o0.x, nsc * o0.y), pack_bf16(nsc * o0.z, nsc * o0.w),
                                       pack_bf16(nsc * o1.x, nsc * o1.y), pack_bf16(nsc * o1.z, nsc * o1.w));
    }
  }
  // block reduction of the per-warp column sums, one 256-column group at a time
  float* pg = partial + static_cast<size_t>(blockIdx.x) * 2 * D;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s_red[warp][0][lane * 8 + e] = dg[v][e];
      s_red[warp][1][lane * 8 + e] = db[v][e];
    }
    __syncthreads();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += s_red[w][which][threadIdx.x];
      pg[which * D + v * 256 + threadIdx.x] = s;
    }
  }
}

// out[n] (+)= sum_p partial[p*stride + n].  Block = 32 columns x 8 warps; the warps split the P partials (P is up to a
// few hundred), then combine through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) finalize_sum_kernel(const float* __restrict__ partial, int P, int64_t stride,
                                                           float* __restrict__ out, int n, int accumulate) {
  __shared__ float s_red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (i < n) {
    int p = warp;
    for (; p + 8 < P; p += 16) {
      s0 += partial[p * stride + i];
      s1 += partial[(p + 8) * stride + i];
    }
    if (p < P) s0 += partial[p * stride + i];
  }
  s_red[warp][lane] = s0 + s1;
  __syncthreads();
  if (warp == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][lane];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// ---------------------------------------------------------------------------------------------- column sums
// partial[chunk][n] = sum over the chunk's rows of x[r, n]   (x bf16, 8 columns per thread)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int rows, int N,
                                                          float* __restrict__ partial) {
  __shared__ float s_red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 256 + lane * 8;
  const int chunks = gridDim.y;
  const int per = (rows + chunks - 1) / chunks;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col0 < N) {
    const __nv_bfloat16* xc = x + col0;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {  // four independent 16-byte loads in flight per lane
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(xc + static_cast<int64_t>(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[0] += bf_lo(v[u].x); acc[1] += bf_hi(v[u].x); acc[2] += bf_lo(v[u].y); acc[3] += bf_hi(v[u].y);
        acc[4] += bf_lo(v[u].z); acc[5] += bf_hi(v[u].z); acc[6] += bf_lo(v[u].w); acc[7] += bf_hi(v[u].w);
      }
    }
    for (; r < r1; r += 8) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(xc + static_cast<int64_t>(r) * ld));
      acc[0] += bf_lo(v.x); acc[1] += bf_hi(v.x); acc[2] += bf_lo(v.y); acc[3] += bf_hi(v.y);
      acc[4] += bf_lo(v.z); acc[5] += bf_hi(v.z); acc[6] += bf_lo(v.w); acc[7] += bf_hi(v.w);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_red[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.y) * N + c] = s;
  }
}

// out[r, :] = scale[r] * x[r, :]  (bf16, 8 columns per thread)
__global__ void scale_rows_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ scale,
                                       __nv_bfloat16* __restrict__ out, int64_t ldo, int rows, int nvec) {
  const int64_t total = static_cast<int64_t>(rows) * nvec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nvec), c = static_cast<int>(i % nvec);
    const float s = __ldg(scale + r);
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(x + r * ldx) + c);
    reinterpret_cast<uint4*>(out + r * ldo)[c] =
        make_uint4(pack_bf16(s * bf_lo(v.x), s * bf_hi(v.x)), pack_bf16(s * bf_lo(v.y), s * bf_hi(v.y)),
                   pack_bf16(s * bf_lo(v.z), s * bf_hi(v.z)), pack_bf16(s * bf_lo(v.w), s * bf_hi(v.w)));
  }
}

// ---------------------------------------------------------------------------------------------- embedding gradients
// dpos[1+n, :] (+)= sum_{b,t} G[(b*N+n)*T+t, :]   (block per n, float4 per thread)
__global__ void embed_bwd_pos_kernel(const float* __restrict__ G, float* __restrict__ dpos, int B, int N, int T, int D,
                                     int accumulate) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4* g = reinterpret_cast<const float4*>(G + (static_cast<int64_t>(b) * N + n) * T * D) + c;
      for (int t = 0; t < T; ++t) {
        const float4 v = __ldcs(g + static_cast<int64_t>(t) * (D / 4));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    float4* o = reinterpret_cast<float4*>(dpos + static_cast<int64_t>(1 + n) * D) + c;
    if (accumulate) { const float4 p = *o; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    *o = s;
  }
}

// partial[chunk][t][:] = sum over the chunk's (b,n) sequences of G[(b*N+n)*T+t, :]   (grid: T x chunks)
__global__ void embed_bwd_time_kernel(const float* __restrict__ G, float* __restrict__ partial, int BN, int T, int D) {
  const int t = blockIdx.x, chunks = gridDim.y;
  const int per = (BN + chunks - 1) / chunks;
  const int s0 = blockIdx.y * per, s1 = min(BN, s0 + per);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = s0; q < s1; ++q) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(G + (static_cast<int64_t>(q) * T + t) * D) + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(partial + (static_cast<int64_t>(blockIdx.y) * T + t) * D)[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------- mask head adjoint
// Adjoint of tcow_mask_upsample: one CTA per (b, c, t) image.  Bilinear (align_corners=True) weights are the hat
// function w(l, y) = max(0, 1 - |s*y - l|): d_low[ly,lx] = sum_y sum_x w(ly,y) w(lx,x) d_out[y,x], evaluated as two
// separable gathers (no atomics, deterministic).  Nearest: the sum over the stride x stride block.
// Writes bf16 into d_low[(b*N+n)*T+t, c*pp*pp + i*pp + j] (the dY operand of the head weight-gradient GEMM).
__global__ void __launch_bounds__(256) mask_upsample_bwd_kernel(const float* __restrict__ d_out, __nv_bfloat16* __restrict__ d_low,
                                                                int64_t ld_low, int B, int T, int Ho, int Wo, int C, int pp,
                                                                int stride, int mode) {
  extern __shared__ float s_tmp[];  // [Hf][Wl]
  const int Hl = Ho * pp, Wl = Wo * pp;
  const int Hf = Hl * stride, Wf = Wl * stride;
  const int N = Ho * Wo;
  const int img = blockIdx.x;  // ((b*C + c)*T + t)
  const int t = img % T, c = (img / T) % C, b = img / (T * C);
  const float* src = d_out + static_cast<int64_t>(img) * Hf * Wf;
  const bool nearest = (mode == 1 || stride == 1);
  const float sy = (Hf > 1) ? static_cast<float>(Hl - 1) / static_cast<float>(Hf - 1) : 0.f;
  const float sx = (Wf > 1) ? static_cast<float>(Wl - 1) / static_cast<float>(Wf - 1) : 0.f;
  // pass 1: along x
  for (int i = threadIdx.x; i < Hf * Wl; i += blockDim.x) {
    const int y = i / Wl, lx = i % Wl;
    const float* r = src + static_cast<int64_t>(y) * Wf;
    float s = 0.f;
    if (nearest) {
      for (int e = 0; e < stride; ++e) s += __ldg(r + lx * stride + e);
    } else {
      // x with |sx*x - lx| < 1
      int x_lo = (lx == 0 || sx == 0.f) ? 0 : static_cast<int>(floorf(static_cast<float>(lx - 1) / sx));
      int x_hi = sx == 0.f ? Wf - 1 : min(Wf - 1, static_cast<int>(ceilf(static_cast<float>(lx + 1) / sx)));
      x_lo = max(0, x_lo);
      for (int x = x_lo; x <= x_hi; ++x) {
        // the forward's weights: x0 = int(fx), w(x0) = 1 - (fx - x0), w(x0+1) = fx - x0
        const float fx = sx * static_cast<float>(x);
        const int x0 = static_cast<int>(fx);
        const int x1 = x0 + ((x0 < Wl - 1) ? 1 : 0);
        const float w1 = fx - static_cast<float>(x0), w0 = 1.f - w1;
        float w = 0.f;
        if (x0 == lx) w += w0;
        if (x1 == lx) w += w1;
        if (w != 0.f) s = fmaf(w, __ldg(r + x), s);
      }
    }
    s_tmp[i] = s;
  }
  __syncthreads();
  // pass 2: along y, then scatter into the token-major gradient
  for (int i = threadIdx.x; i < Hl * Wl; i += blockDim.x) {
    const int ly = i / Wl, lx = i % Wl;
    float s = 0.f;
    if (nearest) {
      for (int e = 0; e < stride; ++e) s += s_tmp[(ly * stride + e) * Wl + lx];
    } else {
      int y_lo = (ly == 0 || sy == 0.f) ? 0 : static_cast<int>(floorf(static_cast<float>(ly - 1) / sy));
      int y_hi = sy == 0.f ? Hf - 1 : min(Hf - 1, static_cast<int>(ceilf(static_cast<float>(ly + 1) / sy)));
      y_lo = max(0, y_lo);
      for (int y = y_lo; y <= y_hi; ++y) {
        const float fy = sy * static_cast<float>(y);
        const int y0 = static_cast<int>(fy);
        const int y1 = y0 + ((y0 < Hl - 1) ? 1 : 0);
        const float w1 = fy - static_cast<float>(y0), w0 = 1.f - w1;
        float w = 0.f;
        if (y0 == ly) w += w0;
        if (y1 == ly) w += w1;
        if (w != 0.f) s = fmaf(w, s_tmp[y * Wl + lx], s);
      }
    }
    const int n = (ly / pp) * Wo + lx / pp;
    d_low[((static_cast<int64_t>(b) * N + n) * T + t) * ld_low + (c * pp + (ly % pp)) * pp + (lx % pp)] = __float2bfloat16(s);
  }
}

// Columns [col0, ncols) of d_low: the flag gradient d_flags[b,t,f] / N (adjoint of the spatial mean,
// mask_tracker.py:137), zero in the padding columns.
__global__ void flag_mean_bwd_kernel(const float* __restrict__ d_flags, __nv_bfloat16* __restrict__ d_low, int64_t ld_low,
                                     int B, int N, int T, int F, int col0, int ncols) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.y + threadIdx.y;
  if (row >= static_cast<int64_t>(B) * N * T) return;
  const int t = static_cast<int>(row % T), b = static_cast<int>(row / (static_cast<int64_t>(N) * T));
  const float inv = 1.0f / static_cast<float>(N);
  for (int c = col0 + threadIdx.x; c < ncols; c += blockDim.x) {
    const int f = c - col0;
    const float v = (d_flags != nullptr && f < F) ? d_flags[(static_cast<int64_t>(b) * T + t) * F + f] * inv : 0.f;
    d_low[row * ld_low + c] = __float2bfloat16(v);
  }
}

// cls gradient fan-out (adjoint of tcow_cls_merge / the in-kernel frame-0 write of the spatial attention):
// d_out_cls[b,t,:] = d_out[cls_row0+b,:] * (mode 0: 1/T for every t (mean, vit.py:195); mode 1: t==0 only (vit.py:198))
__global__ void cls_merge_bwd_kernel(const __nv_bfloat16* __restrict__ d_out, int64_t ld, float* __restrict__ d_out_cls, int B,
                                     int T, int D, int64_t cls_row0, int mode) {
  const int bt = blockIdx.x, b = bt / T, t = bt % T;
  const float sc = mode == 0 ? 1.0f / static_cast<float>(T) : (t == 0 ? 1.f : 0.f);
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    d_out_cls[static_cast<int64_t>(bt) * D + c] = sc * __bfloat162float(d_out[(cls_row0 + b) * ld + c]);
}

template <int NV>
static int launch_ln_train(const float* x, const float* g, const float* b, void* y, void* xhat, float* rstd, int rows,
                           float eps, cudaStream_t s) {
  long long blocks = (static_cast<long long>(rows) + 7) / 8;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  layernorm_train_kernel<NV><<<static_cast<int>(blocks), 256, 0, s>>>(x, g, b, static_cast<__nv_bfloat16*>(y),
                                                                     static_cast<__nv_bfloat16*>(xhat), rstd, rows, eps);
  return check_launch("layernorm_train_kernel");
}

template <int NV>
static int launch_ln_bwd(const void* dy, const void* xhat, const float* rstd, const float* gamma, float* G, void* Gb,
                         float* dgamma, float* dbeta, float* ws, int rows, int accumulate, cudaStream_t s,
                         const float* next_scale = nullptr, void* Gs = nullptr) {
  constexpr int D = NV * 256;
  int blocks = (rows + 7) / 8;
  if (blocks > LNB_MAX_BLOCKS) blocks = LNB_MAX_BLOCKS;
  layernorm_bwd_kernel<NV><<<blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(xhat),
                                                 rstd, gamma, G, static_cast<__nv_bfloat16*>(Gb), ws, rows, accumulate,
                                                 next_scale, static_cast<__nv_bfloat16*>(Gs));
  int rc = check_launch("layernorm_bwd_kernel");
  if (rc) return rc;
  finalize_sum_kernel<<<(D + 31) / 32, 256, 0, s>>>(ws, blocks, 2 * D, dgamma, D, 1);
  finalize_sum_kernel<<<(D + 31) / 32, 256, 0, s>>>(ws + D, blocks, 2 * D, dbeta, D, 1);
  return check_launch("finalize_sum_kernel");
}

}  // namespace tcow

extern "C" int64_t tcow_train_workspace_floats(int max_cols) {
  // the largest of: LayerNorm bwd partials (blocks x 2 x D), column-sum partials (chunks x N), time-embed partials
  const int64_t a = static_cast<int64_t>(tcow::LNB_MAX_BLOCKS) * 2 * 1024;
  const int64_t b = static_cast<int64_t>(tcow::CS_MAX_CHUNKS) * (max_cols > 0 ? max_cols : 4096);  // colsum: chunks x N
  return a > b ? a : b;
}

extern "C" int tcow_layernorm_bf16_train(const float* x, const float* gamma, const float* beta, void* y, void* xhat,
                                         float* rstd, int rows, int D, float eps, void* stream) {
  using namespace tcow;
  if (!x || !y || !xhat || !rstd || !gamma || !beta || rows <= 0) return set_error(TCOW_ERR_ARG, "layernorm_train: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 768: return launch_ln_train<6>(x, gamma, beta, y, xhat, rstd, rows, eps, s);
    case 1024: return launch_ln_train<8>(x, gamma, beta, y, xhat, rstd, rows, eps, s);
  }
  return set_error(TCOW_ERR_ARG, "layernorm_train: unsupported width %d (768 or 1024)", D);
}

extern "C" int tcow_layernorm_bwd(const void* dy, const void* xhat, const float* rstd, const float* gamma, float* G,
                                  void* Gb, float* dgamma, float* dbeta, float* workspace, int rows, int D,
                                  int accumulate, void* stream) {
  using namespace tcow;
  if (!dy || !xhat || !rstd || !gamma || !G || !Gb || !dgamma || !dbeta || !workspace || rows <= 0)
    return set_error(TCOW_ERR_ARG, "layernorm_bwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 768: return launch_ln_bwd<3>(dy, xhat, rstd, gamma, G, Gb, dgamma, dbeta, workspace, rows, accumulate, s);
    case 1024: return launch_ln_bwd<4>(dy, xhat, rstd, gamma, G, Gb, dgamma, dbeta, workspace, rows, accumulate, s);
  }
  return set_error(TCOW_ERR_ARG, "layernorm_bwd: unsupported width %d (768 or 1024)", D);
}

extern "C" int tcow_layernorm_bwd_scaled(const void* dy, const void* xhat, const float* rstd, const float* gamma, float* G,
                                         void* Gb, float* dgamma, float* dbeta, float* workspace, int rows, int D,
                                         int accumulate, const float* next_scale, void* Gs, void* stream) {
  using namespace tcow;
  if (!dy || !xhat || !rstd || !gamma || !G || !Gb || !dgamma || !dbeta || !workspace || !next_scale || !Gs || rows <= 0)
    return set_error(TCOW_ERR_ARG, "layernorm_bwd_scaled: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (D) {
    case 768: return launch_ln_bwd<3>(dy, xhat, rstd, gamma, G, Gb, dgamma, dbeta, workspace, rows, accumulate, s, next_scale, Gs);
    case 1024: return launch_ln_bwd<4>(dy, xhat, rstd, gamma, G, Gb, dgamma, dbeta, workspace, rows, accumulate, s, next_scale, Gs);
  }
  return set_error(TCOW_ERR_ARG, "layernorm_bwd_scaled: unsupported width %d (768 or 1024)", D);
}

extern "C" int tcow_colsum_bf16(const void* x, int64_t ldx, int rows, int N, float* out, float* workspace,
                                int accumulate, void* stream) {
  using namespace tcow;
  if (!x || !out || !workspace || rows <= 0 || N <= 0) return set_error(TCOW_ERR_ARG, "colsum: bad argument");
  if ((N % 8) || (ldx % 8) || (reinterpret_cast<uintptr_t>(x) & 15)) return set_error(TCOW_ERR_ARG, "colsum: N and pitch must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // ~8 blocks per SM over (column groups x row chunks), at least 64 rows per chunk
  const int colgroups = (N + 255) / 256;
  int chunks = (8 * sm_count() + colgroups - 1) / colgroups;
  if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
  if (chunks > CS_MAX_CHUNKS) chunks = CS_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  if (static_cast<int64_t>(chunks) * N > tcow_train_workspace_floats(N)) return set_error(TCOW_ERR_ARG, "colsum: workspace too small");
  colsum_bf16_kernel<<<dim3((N + 255) / 256, chunks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, rows, N, workspace);
  int rc = check_launch("colsum_bf16_kernel");
  if (rc) return rc;
  finalize_sum_kernel<<<(N + 31) / 32, 256, 0, s>>>(workspace, chunks, N, out, N, accumulate);
  return check_launch("finalize_sum_kernel");
}

extern "C" int tcow_embed_bwd(const float* G, float* dpos, float* dtime, float* dcls_pos0, float* workspace, int B, int N,
                              int T, int D, int accumulate, void* stream) {
  using namespace tcow;
  if (!G || !dpos || !dtime || !workspace || B <= 0 || N <= 0 || T <= 0 || D % 4) return set_error(TCOW_ERR_ARG, "embed_bwd: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  embed_bwd_pos_kernel<<<N, 192, 0, s>>>(G, dpos, B, N, T, D, accumulate);
  int chunks = (B * N + 63) / 64;
  if (chunks > 16) chunks = 16;
  if (static_cast<int64_t>(chunks) * T * D > tcow_train_workspace_floats(0)) return set_error(TCOW_ERR_ARG, "embed_bwd: workspace too small");
  embed_bwd_time_kernel<<<dim3(T, chunks), 192, 0, s>>>(G, workspace, B * N, T, D);
  finalize_sum_kernel<<<(T * D + 31) / 32, 256, 0, s>>>(workspace, chunks, static_cast<int64_t>(T) * D, dtime, T * D, accumulate);
  if (dcls_pos0)  // d(cls_token) = d(pos_embed[0]) = sum_b G[M + b]
    finalize_sum_kernel<<<(D + 31) / 32, 256, 0, s>>>(G + static_cast<int64_t>(B) * N * T * D, B, D, dcls_pos0, D, accumulate);
  return check_launch("embed_bwd");
}

extern "C" int tcow_mask_head_bwd(const float* d_out, const float* d_flags, void* d_low, int64_t ld_low, int B, int T,
                                  int Ho, int Wo, int C, int pp, int stride, int mode, int F, int col0, int ncols,
                                  void* stream) {
  using namespace tcow;
  if (!d_out || !d_low || B <= 0 || T <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || pp <= 0 || stride <= 0)
    return set_error(TCOW_ERR_ARG, "mask_head_bwd: bad argument");
  if (col0 != C * pp * pp || ncols < col0 + F || ld_low < ncols) return set_error(TCOW_ERR_ARG, "mask_head_bwd: column layout mismatch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(Ho) * pp * stride * Wo * pp * sizeof(float);
  if (smem > 200 * 1024) return set_error(TCOW_ERR_ARG, "mask_head_bwd: %zu bytes of shared memory needed", smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mask_upsample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  mask_upsample_bwd_kernel<<<B * C * T, 256, smem, s>>>(d_out, static_cast<__nv_bfloat16*>(d_low), ld_low, B, T, Ho, Wo, C, pp,
                                                       stride, mode);
  int rc = check_launch("mask_upsample_bwd_kernel");
  if (rc) return rc;
  if (ncols > col0) {
    const int64_t rows = static_cast<int64_t>(B) * Ho * Wo * T;
    flag_mean_bwd_kernel<<<static_cast<unsigned>((rows + 15) / 16), dim3(16, 16), 0, s>>>(d_flags, static_cast<__nv_bfloat16*>(d_low), ld_low,
                                                                                     B, Ho * Wo, T, F, col0, ncols);
    rc = check_launch("flag_mean_bwd_kernel");
  }
  return rc;
}

extern "C" int tcow_cls_merge_bwd(const void* d_out, int64_t ld_out, float* d_out_cls, int B, int T, int D,
                                  int64_t cls_row0, int mode, void* stream) {
  using namespace tcow;
  if (!d_out || !d_out_cls || B <= 0 || T <= 0 || D <= 0) return set_error(TCOW_ERR_ARG, "cls_merge_bwd: bad argument");
  cls_merge_bwd_kernel<<<B * T, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(d_out), ld_out, d_out_cls, B, T, D,
                                                                         cls_row0, mode);
  return check_launch("cls_merge_bwd_kernel");
}

extern "C" int tcow_scale_rows_bf16(const void* x, int64_t ldx, const float* scale, void* out, int64_t ld_out, int rows,
                                    int N, void* stream) {
  using namespace tcow;
  if (!x || !scale || !out || rows <= 0 || N <= 0) return set_error(TCOW_ERR_ARG, "scale_rows: bad argument");
  if ((N % 8) || (ldx % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "scale_rows: N and pitches must be multiples of 8");
  const int64_t total = static_cast<int64_t>(rows) * (N / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  scale_rows_bf16_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, scale, static_cast<__nv_bfloat16*>(out), ld_out, rows, N / 8);
  return check_launch("scale_rows_bf16_kernel");
}
