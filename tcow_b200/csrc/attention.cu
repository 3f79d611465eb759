// Fused attention kernels for the divided space-time block (vit.py:78-123 Attention.forward as called at
// vit.py:172 (temporal, causal) and vit.py:186 (spatial, full)).  Both read q/k/v straight out of the QKV GEMM
// output in the canonical token-row layout — no rearranges (vit.py:170,173,181-185,210 are eliminated) — keep
// the softmax state in registers (fp32, exp2 with pre-scaled logits, warp-shuffle row reductions), apply the
// mask in-kernel and never materialise the score matrix.
//
// Warp-level mma.sync m16n8k16 bf16 with ldmatrix from XOR-swizzled shared memory.  The temporal kernel (30x30
// problems, bandwidth-shaped) is the product path; the spatial kernel here is the independent mma.sync implementation
// kept for A/B checks (TCOW_SPATIAL_IMPL=mma) — the product path is tcgen05/TMEM (attn_spatial_tc.cu).
#include <math.h>
#include <stdlib.h>

#include "attn_frag.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

// ============================================================================================ temporal
// One warp per (sequence, head).  T <= T_PAD in {32, 64}; all keys in one pass (no online rescale).
template <int T_PAD>
__global__ void __launch_bounds__(128) attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld_qkv,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                            int num_seq, int T, int heads, int causal_diag,
                                                            float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_att[];
  constexpr int TILE = T_PAD * ROW_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long work = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (work >= static_cast<long long>(num_seq) * heads) return;
  const int seq = static_cast<int>(work / heads), head = static_cast<int>(work % heads);
  const uint32_t sQ = smem_u32(smem_att) + warp * 3 * TILE;
  const uint32_t sK = sQ + TILE, sV = sK + TILE;
  const int D = heads * HD;

  // ---- load q, k, v rows of this (sequence, head): T rows x 128 B each, 16 B per cp.async
  const __nv_bfloat16* src = qkv + static_cast<int64_t>(seq) * T * ld_qkv + head * HD;
  for (int idx = lane; idx < T * 8 * 3; idx += 32) {
    const int which = idx / (T * 8), rem = idx % (T * 8);
    const int row = rem >> 3, chunk = rem & 7;
    cp_async_16(sw_addr(sQ + which * TILE, row, chunk), src + static_cast<int64_t>(row) * ld_qkv + which * D + chunk * 8);
  }
  cp_async_commit();
  for (int idx = lane; idx < (T_PAD - T) * 8 * 3; idx += 32) {  // zero the padding rows (V must be finite)
    const int which = idx / ((T_PAD - T) * 8), rem = idx % ((T_PAD - T) * 8);
    const int row = T + (rem >> 3), chunk = rem & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sw_addr(sQ + which * TILE, row, chunk)), "r"(0) : "memory");
  }
  cp_async_wait<0>();
  __syncwarp();

  temporal_attend_seq<T_PAD>(sQ, sK, sV, 0, T, causal_diag, scale_log2);
  __syncwarp();
  __nv_bfloat16* dst = out + static_cast<int64_t>(seq) * T * ld_out + head * HD;
  for (int idx = lane; idx < T * 8; idx += 32) {
    const int row = idx >> 3, chunk = idx & 7;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw_addr(sQ, row, chunk)));
    *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(row) * ld_out + chunk * 8) = v;
  }
}

// ============================================================================================ spatial
// CTA = (query block of 32*NW tokens, (clip b, frame t, head)); flash loop over 64-key blocks, double buffered.
// Token i of frame (b,t): use_cls ? (i == 0 ? cls row of clip b : patch i-1) : patch i; patch n lives at
// canonical row (b*N+n)*T+t, so consecutive tokens are T rows apart — gathered with 128-byte cp.async rows.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
attn_spatial_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld_qkv, __nv_bfloat16* __restrict__ out,
                    int64_t ld_out, float* __restrict__ out_cls, int B, int N, int T, int heads, int use_cls,
                    int64_t cls_row0, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_att[];
  constexpr int BQ = NW * 32;
  constexpr int KB = 64;
  constexpr int KV_TILE = KB * ROW_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = N + use_cls;
  const int q0 = blockIdx.x * BQ;
  const int bth = blockIdx.y;
  const int head = bth % heads;
  const int t = (bth / heads) % T;
  const int b = bth / (heads * T);
  const int D = heads * HD;
  const uint32_t sQ = smem_u32(smem_att);
  const uint32_t sKV = sQ + BQ * ROW_BYTES;  // [2 stages][K | V]

  auto token_row = [&](int i) -> int64_t {
    if (use_cls) {
      if (i == 0) return cls_row0 + b;
      --i;
    }
    return (static_cast<int64_t>(b) * N + i) * T + t;
  };
  auto load_kv = [&](int kb, int stage) {
    const uint32_t dK = sKV + stage * 2 * KV_TILE, dV = dK + KV_TILE;
    for (int idx = threadIdx.x; idx < KB * 8 * 2; idx += BQ) {
      const int which = idx / (KB * 8), rem = idx % (KB * 8);
      const int row = rem >> 3, chunk = rem & 7;
      const int tok = kb * KB + row;
      const uint32_t d = sw_addr(which ? dV : dK, row, chunk);
      if (tok < S)
        cp_async_16(d, qkv + token_row(tok) * ld_qkv + (1 + which) * D + head * HD + chunk * 8);
      else
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(d), "r"(0) : "memory");
    }
  };

  // ---- Q block + first K/V block
  for (int idx = threadIdx.x; idx < BQ * 8; idx += BQ) {
    const int row = idx >> 3, chunk = idx & 7;
    const int tok = q0 + row;
    const uint32_t d = sw_addr(sQ, row, chunk);
    if (tok < S)
      cp_async_16(d, qkv + token_row(tok) * ld_qkv + head * HD + chunk * 8);
    else
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(d), "r"(0) : "memory");
  }
  load_kv(0, 0);
  cp_async_commit();

  const int nkb = (S + KB - 1) / KB;
  const int g = lane >> 2, tq = lane & 3;
  const int wrow0 = warp * 32;                    // this warp's first query row inside the block
  const bool active = (q0 + wrow0) < S;           // warps whose 32 queries are all padding only help loading
  uint32_t qa[2][4][4];
  float o[2][8][4];
  float m_run[2][2], l_run[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[mt][nd][c] = 0.f;
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
  }

#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    const int stage = kb & 1;
    if (kb + 1 < nkb) load_kv(kb + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kb == 0 && active) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) load_a_frag(sQ, wrow0 + mt * 16, ks, qa[mt][ks]);
    }
    if (active) {
      const uint32_t sK = sKV + stage * 2 * KV_TILE, sV = sK + KV_TILE;
      float s[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) s[mt][nt][c] = 0.f;
#pragma unroll
      for (int np = 0; np < 4; ++np)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t bf[4];
          load_bk_frag(sK, np * 16, ks, bf);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_bf16_16816(s[mt][2 * np], qa[mt][ks], bf[0], bf[1]);
            mma_bf16_16816(s[mt][2 * np + 1], qa[mt][ks], bf[2], bf[3]);
          }
        }
      const int kbase = kb * KB;
      const bool tail = (kbase + KB > S);
      uint32_t pa[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float mx[2] = {m_run[mt][0], m_run[mt][1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v = s[mt][nt][c] * scale_log2;
            if (tail && (kbase + nt * 8 + tq * 2 + (c & 1)) >= S) v = -INFINITY;
            s[mt][nt][c] = v;
            mx[c >> 1] = fmaxf(mx[c >> 1], v);
          }
        float alpha[2], sum[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
          mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
          alpha[h] = exp2f(m_run[mt][h] - mx[h]);  // first block: exp2(-inf) = 0
          m_run[mt][h] = mx[h];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float p = exp2f(s[mt][nt][c] - mx[c >> 1]);
            s[mt][nt][c] = p;
            sum[c >> 1] += p;
          }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
          sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
          l_run[mt][h] = l_run[mt][h] * alpha[h] + sum[h];
        }
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          o[mt][nd][0] *= alpha[0];
          o[mt][nd][1] *= alpha[0];
          o[mt][nd][2] *= alpha[1];
          o[mt][nd][3] *= alpha[1];
        }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          pa[mt][kt][0] = pack_bf16(s[mt][2 * kt][0], s[mt][2 * kt][1]);
          pa[mt][kt][1] = pack_bf16(s[mt][2 * kt][2], s[mt][2 * kt][3]);
          pa[mt][kt][2] = pack_bf16(s[mt][2 * kt + 1][0], s[mt][2 * kt + 1][1]);
          pa[mt][kt][3] = pack_bf16(s[mt][2 * kt + 1][2], s[mt][2 * kt + 1][3]);
        }
      }
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bf[4];
          load_bv_frag(sV, kt * 16, np * 2, bf);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_bf16_16816(o[mt][2 * np], pa[mt][kt], bf[0], bf[1]);
            mma_bf16_16816(o[mt][2 * np + 1], pa[mt][kt], bf[2], bf[3]);
          }
        }
    }
    __syncthreads();  // everyone is done with `stage` before it is refilled
  }

  // ---- normalise; cls query -> out_cls (fp32); everything else staged as bf16 in this warp's Q rows
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float inv = 1.0f / l_run[mt][h];
        const int row = wrow0 + mt * 16 + g + h * 8;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          const float v0 = o[mt][nd][2 * h] * inv, v1 = o[mt][nd][2 * h + 1] * inv;
          if (use_cls && (q0 + row) == 0) {
            float* dc = out_cls + (static_cast<int64_t>(b) * T + t) * D + head * HD + nd * 8 + tq * 2;
            dc[0] = v0;
            dc[1] = v1;
            if (t == 0)  // frame-0 cls output doubles as the cls input row of the projection (vit.py:198)
              *reinterpret_cast<uint32_t*>(out + (cls_row0 + b) * ld_out + head * HD + nd * 8 + tq * 2) = pack_bf16(v0, v1);
          }
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw_addr(sQ, row, nd) + tq * 4), "r"(pack_bf16(v0, v1)) : "memory");
        }
      }
    __syncwarp();
    for (int idx = lane; idx < 32 * 8; idx += 32) {
      const int row = wrow0 + (idx >> 3), chunk = idx & 7;
      const int tok = q0 + row;
      if (tok >= S || (use_cls && tok == 0)) continue;
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw_addr(sQ, row, chunk)));
      *reinterpret_cast<uint4*>(out + token_row(tok) * ld_out + head * HD + chunk * 8) = v;
    }
  }
}

// cls residual input: out[cls_row0+b,:] = out_cls[b,0,:] (mode 1) or mean_t out_cls[b,t,:] (mode 0).
__global__ void cls_merge_kernel(const float* __restrict__ out_cls, __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                 int B, int T, int D, int64_t cls_row0, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx % D;
  float v;
  if (mode == 1) {
    v = out_cls[static_cast<int64_t>(b) * T * D + d];
  } else {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += out_cls[(static_cast<int64_t>(b) * T + t) * D + d];
    v = s / static_cast<float>(T);
  }
  out[(cls_row0 + b) * ld_out + d] = __float2bfloat16_rn(v);
}

template <int T_PAD>
static int launch_temporal(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int num_seq, int T, int heads,
                           int causal_diag, cudaStream_t s) {
  auto kern = attn_temporal_kernel<T_PAD>;
  constexpr int smem = 4 * 3 * T_PAD * ROW_BYTES;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const long long work = static_cast<long long>(num_seq) * heads;
  const long long blocks = (work + 3) / 4;
  if (blocks > 0x7fffffffLL) return set_error(TCOW_ERR_ARG, "attn_temporal: too many sequences");
  kern<<<static_cast<unsigned>(blocks), 128, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), ld_qkv,
                                                        static_cast<__nv_bfloat16*>(out), ld_out, num_seq, T, heads,
                                                        causal_diag, 0.125f * 1.4426950408889634f);
  return check_launch("attn_temporal_kernel");
}

template <int NW>
static int launch_spatial(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N,
                          int T, int heads, int use_cls, int64_t cls_row0, cudaStream_t s) {
  auto kern = attn_spatial_kernel<NW>;
  constexpr int smem = NW * 32 * ROW_BYTES + 2 * 2 * 64 * ROW_BYTES;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const int S = N + use_cls;
  dim3 grid((S + NW * 32 - 1) / (NW * 32), B * T * heads);
  if (grid.y > 65535u) return set_error(TCOW_ERR_ARG, "attn_spatial: B*T*heads = %u exceeds 65535; split the batch", grid.y);
  kern<<<grid, NW * 32, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out),
                                   ld_out, out_cls, B, N, T, heads, use_cls, cls_row0,
                                   0.125f * 1.4426950408889634f);
  return check_launch("attn_spatial_kernel");
}

}  // namespace tcow

extern "C" int tcow_attn_temporal(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int num_seq, int T,
                                  int heads, int causal_diag, void* stream) {
  using namespace tcow;
  if (!qkv || !out || num_seq <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_temporal: bad argument");
  if (T < 1 || T > 64) return set_error(TCOW_ERR_ARG, "attn_temporal: T=%d unsupported (1..64)", T);
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_temporal: row pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T <= 32) return launch_temporal<32>(qkv, ld_qkv, out, ld_out, num_seq, T, heads, causal_diag, s);
  return launch_temporal<64>(qkv, ld_qkv, out, ld_out, num_seq, T, heads, causal_diag, s);
}

extern "C" int tcow_attn_spatial(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B,
                                 int N, int T, int heads, int use_cls, int64_t cls_row0, void* stream) {
  using namespace tcow;
  if (!qkv || !out || B <= 0 || N <= 0 || T <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_spatial: bad argument");
  if (use_cls && !out_cls) return set_error(TCOW_ERR_ARG, "attn_spatial: out_cls required when use_cls");
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_spatial: row pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int S = N + (use_cls ? 1 : 0);
  // Tensor-core (tcgen05/TMEM) kernels: K/V resident in shared memory whenever one frame's keys fit (S <= 304), else
  // streamed in 128-key blocks (e.g. 480x640 frames, S = 1201).  TCOW_SPATIAL_IMPL=mma forces the mma.sync flash kernel
  // below (kept as an independent implementation for A/B checks), =stream forces the streamed kernel for every S.
  static const char impl = [] { const char* e = getenv("TCOW_SPATIAL_IMPL"); return e ? e[0] : '\0'; }();
  const bool force_mma = impl == 'm';
  if (!force_mma && (ld_qkv % 8) == 0) {
    if (S <= 304 && impl == 'r')   // round-1 resident kernel (two CTAs per SM), kept for A/B timing
      return launch_spatial_tc(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
    if (S <= 304 && impl != 's')
      return launch_spatial_pp(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
    return launch_spatial_stream(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
  }
  // pick the query-block width that wastes the fewest padded query rows (ties -> wider block)
  const int pad10 = ((S + 319) / 320) * 320, pad4 = ((S + 127) / 128) * 128;
  if (pad10 <= pad4) return launch_spatial<10>(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
  return launch_spatial<4>(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
}

extern "C" int tcow_attn_spatial_train(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls,
                                       float* lse, int B, int N, int T, int heads, int use_cls, int64_t cls_row0,
                                       void* stream) {
  using namespace tcow;
  if (!qkv || !out || !lse || B <= 0 || N <= 0 || T <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_spatial_train: bad argument");
  if (use_cls && !out_cls) return set_error(TCOW_ERR_ARG, "attn_spatial_train: out_cls required when use_cls");
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_spatial_train: row pitches must be multiples of 8");
  if (N + (use_cls ? 1 : 0) > 304)
    return set_error(TCOW_ERR_ARG, "attn_spatial_train: %d tokens per frame > 304 not supported in training", N + (use_cls ? 1 : 0));
  static const char impl = [] { const char* e = getenv("TCOW_SPATIAL_IMPL"); return e ? e[0] : '\0'; }();
  if (impl == 'r')
    return launch_spatial_tc(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0,
                             static_cast<cudaStream_t>(stream), lse);
  return launch_spatial_pp(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0,
                           static_cast<cudaStream_t>(stream), lse);
}

extern "C" int tcow_cls_merge(const float* out_cls, void* out, int64_t ld_out, int B, int T, int D, int64_t cls_row0,
                              int mode, void* stream) {
  using namespace tcow;
  if (!out_cls || !out || B <= 0 || T <= 0 || D <= 0) return set_error(TCOW_ERR_ARG, "cls_merge: bad argument");
  const int n = B * D;
  cls_merge_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out_cls, static_cast<__nv_bfloat16*>(out), ld_out, B, T, D, cls_row0, mode);
  return check_launch("cls_merge_kernel");
}
