// Backward of the two attention operators of the divided space-time block (vit.py:78-111 as called at vit.py:172
// (temporal, causal) and vit.py:186 (spatial, full)); flash-style: nothing of size S x S ever leaves the SM.
//
// With P = softmax(Q K^T * hd^-0.5 + mask), O = P V and the upstream gradient dO:
//   D_i = sum_d dO[i,d] O[i,d]          dP = dO V^T           dS = P o (dP - D) * hd^-0.5
//   dQ = dS K                           dK = dS^T Q           dV = P^T dO
// This file holds the TEMPORAL operator (30-long sequences, recomputes its own log-sum-exp) and the entry point of the
// spatial one, whose kernels are tcgen05/TMEM (attn_spatial_bwd_tc.cu, P recomputed from the forward's saved log-sum-exp).  Each 16-row tile of queries (dQ) and of keys (dK, dV) is owned by one warp, which
// walks the other dimension in 16-wide steps with mma.sync m16n8k16 (bf16 in, fp32 accumulate) out of XOR-swizzled
// shared memory; the key-owner pass recomputes S^T = K Q^T so that P^T and dS^T come out of the tensor cores already
// in A-fragment layout — no transposes, no atomics, deterministic.  Gradients are staged per warp and written as
// 128-byte rows into the canonical token-row layout of the qkv gradient (the dY operand of the qkv weight GEMMs).
#include <math.h>
#include <stdlib.h>

#include "attn_frag.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// dQ tile of queries [m0, m0+16): dq (C layout, 16 x 64 fp32) = sum_j dS_ij K_j.  s_lse / s_D: per-query log2-sum-exp
// and D in shared memory.  key j visible to query i iff j < S and (causal_diag < 0 or j <= i + causal_diag).
__device__ __forceinline__ void attn_bwd_dq_tile(uint32_t sQ, uint32_t sK, uint32_t sV, uint32_t sdO, const float* s_lse,
                                                 const float* s_D, int m0, int S, int causal_diag, float scale_log2,
                                                 float scale, float (&dq)[8][4]) {
  const int lane = lane_id();
  const int g = lane >> 2, tq = lane & 3;
  uint32_t qa[4][4], da[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    load_a_frag(sQ, m0, ks, qa[ks]);
    load_a_frag(sdO, m0, ks, da[ks]);
  }
  const float lse[2] = {s_lse[m0 + g], s_lse[m0 + g + 8]};
  const float Dr[2] = {s_D[m0 + g], s_D[m0 + g + 8]};
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int c = 0; c < 4; ++c) dq[nd][c] = 0.f;
  int jend = S;
  if (causal_diag >= 0 && m0 + 16 + causal_diag < jend) jend = m0 + 16 + causal_diag;  // keys >= jend: invisible
#pragma unroll 1
  for (int j0 = 0; j0 < jend; j0 += 16) {
    float s[2][4], dp[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[nt][c] = dp[nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      load_bk_frag(sK, j0, ks, b);
      mma_bf16_16816(s[0], qa[ks], b[0], b[1]);
      mma_bf16_16816(s[1], qa[ks], b[2], b[3]);
      load_bk_frag(sV, j0, ks, b);
      mma_bf16_16816(dp[0], da[ks], b[0], b[1]);
      mma_bf16_16816(dp[1], da[ks], b[2], b[3]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = m0 + g + (c >> 1) * 8;
        const int j = j0 + nt * 8 + tq * 2 + (c & 1);
        const bool ok = (j < S) && (causal_diag < 0 || j <= i + causal_diag);
        const float p = ok ? ex2f(fmaf(s[nt][c], scale_log2, -lse[c >> 1])) : 0.f;
        s[nt][c] = p * (dp[nt][c] - Dr[c >> 1]) * scale;  // dS
      }
    uint32_t a[4];
    a[0] = pack_bf16(s[0][0], s[0][1]);
    a[1] = pack_bf16(s[0][2], s[0][3]);
    a[2] = pack_bf16(s[1][0], s[1][1]);
    a[3] = pack_bf16(s[1][2], s[1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_bv_frag(sK, j0, np * 2, b);
      mma_bf16_16816(dq[2 * np], a, b[0], b[1]);
      mma_bf16_16816(dq[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// dK, dV tiles of keys [j0, j0+16): dv = sum_i P^T_ji dO_i, dk = sum_i dS^T_ji Q_i, with S^T = K Q^T recomputed.
__device__ __forceinline__ void attn_bwd_dkv_tile(uint32_t sQ, uint32_t sK, uint32_t sV, uint32_t sdO, const float* s_lse,
                                                  const float* s_D, int j0, int S, int causal_diag, float scale_log2,
                                                  float scale, float (&dk)[8][4], float (&dv)[8][4]) {
  const int lane = lane_id();
  const int g = lane >> 2, tq = lane & 3;
  uint32_t ka[4][4], va[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    load_a_frag(sK, j0, ks, ka[ks]);
    load_a_frag(sV, j0, ks, va[ks]);
  }
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int c = 0; c < 4; ++c) dk[nd][c] = dv[nd][c] = 0.f;
  // queries i < j0 - causal_diag see none of these keys
  int mstart = 0;
  if (causal_diag >= 0 && j0 - causal_diag > 0) mstart = ((j0 - causal_diag) >> 4) << 4;
#pragma unroll 1
  for (int m0 = mstart; m0 < S; m0 += 16) {
    float st[2][4], dpt[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) st[nt][c] = dpt[nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      load_bk_frag(sQ, m0, ks, b);
      mma_bf16_16816(st[0], ka[ks], b[0], b[1]);
      mma_bf16_16816(st[1], ka[ks], b[2], b[3]);
      load_bk_frag(sdO, m0, ks, b);
      mma_bf16_16816(dpt[0], va[ks], b[0], b[1]);
      mma_bf16_16816(dpt[1], va[ks], b[2], b[3]);
    }
    float p[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int i0 = m0 + nt * 8 + tq * 2;  // the two query columns of this thread in n-tile nt
      const float2 l2 = *reinterpret_cast<const float2*>(s_lse + i0);
      const float2 d2 = *reinterpret_cast<const float2*>(s_D + i0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + g + (c >> 1) * 8;
        const int i = i0 + (c & 1);
        const bool ok = (i < S) && (j < S) && (causal_diag < 0 || j <= i + causal_diag);
        const float pv = ok ? ex2f(fmaf(st[nt][c], scale_log2, -((c & 1) ? l2.y : l2.x))) : 0.f;
        p[nt][c] = pv;
        st[nt][c] = pv * (dpt[nt][c] - ((c & 1) ? d2.y : d2.x)) * scale;  // dS^T
      }
    }
    uint32_t pa[4], dsa[4];
    pa[0] = pack_bf16(p[0][0], p[0][1]);
    pa[1] = pack_bf16(p[0][2], p[0][3]);
    pa[2] = pack_bf16(p[1][0], p[1][1]);
    pa[3] = pack_bf16(p[1][2], p[1][3]);
    dsa[0] = pack_bf16(st[0][0], st[0][1]);
    dsa[1] = pack_bf16(st[0][2], st[0][3]);
    dsa[2] = pack_bf16(st[1][0], st[1][1]);
    dsa[3] = pack_bf16(st[1][2], st[1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_bv_frag(sdO, m0, np * 2, b);
      mma_bf16_16816(dv[2 * np], pa, b[0], b[1]);
      mma_bf16_16816(dv[2 * np + 1], pa, b[2], b[3]);
      load_bv_frag(sQ, m0, np * 2, b);
      mma_bf16_16816(dk[2 * np], dsa, b[0], b[1]);
      mma_bf16_16816(dk[2 * np + 1], dsa, b[2], b[3]);
    }
  }
}

// C-layout 16 x 64 fp32 tile -> bf16 rows in a per-warp swizzled staging tile (16 rows x 128 bytes)
__device__ __forceinline__ void stage_tile(uint32_t stage, const float (&t)[8][4]) {
  const int lane = lane_id();
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw_addr(stage, g + h * 8, nd) + tq * 4),
                   "r"(pack_bf16(t[nd][2 * h], t[nd][2 * h + 1]))
                   : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bfl(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfh(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  return bfl(a.x) * bfl(b.x) + bfh(a.x) * bfh(b.x) + bfl(a.y) * bfl(b.y) + bfh(a.y) * bfh(b.y) + bfl(a.z) * bfl(b.z) +
         bfh(a.z) * bfh(b.z) + bfl(a.w) * bfl(b.w) + bfh(a.w) * bfh(b.w);
}

}  // namespace

// ============================================================================================ temporal
// One warp per (sequence, head); q, k, v, dO, O of the T rows in shared memory (T <= T_PAD in {32, 64}).
template <int T_PAD>
__global__ void __launch_bounds__(128) attn_temporal_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld_qkv,
                                                                const __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                                const __nv_bfloat16* __restrict__ d_out, int64_t ld_do,
                                                                __nv_bfloat16* __restrict__ d_qkv, int64_t ld_dqkv,
                                                                int num_seq, int T, int heads, int causal_diag,
                                                                float scale_log2, float scale) {
  extern __shared__ __align__(1024) uint8_t smem_tb[];
  constexpr int TILE = T_PAD * ROW_BYTES;
  constexpr int WARP_BYTES = 5 * TILE + 2 * T_PAD * 4;  // q, k, v, dO, staging(/O) + lse + D
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long work = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (work >= static_cast<long long>(num_seq) * heads) return;
  const int seq = static_cast<int>(work / heads), head = static_cast<int>(work % heads);
  uint8_t* wbase = smem_tb + warp * WARP_BYTES;
  const uint32_t sQ = smem_u32(wbase);
  const uint32_t sK = sQ + TILE, sV = sK + TILE, sdO = sV + TILE, sO = sdO + TILE;
  float* s_lse = reinterpret_cast<float*>(wbase + 5 * TILE);
  float* s_D = s_lse + T_PAD;
  const int D = heads * HD;
  const int g = lane >> 2, tq = lane & 3;

  // loads without index divisions: lane = (row mod 4, 16-byte chunk)
  const int lrow = lane >> 3, lchunk = lane & 7;
  const __nv_bfloat16* src = qkv + static_cast<int64_t>(seq) * T * ld_qkv + head * HD + lchunk * 8;
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const __nv_bfloat16* p = src + which * D + static_cast<int64_t>(lrow) * ld_qkv;
#pragma unroll 4
    for (int row = lrow; row < T; row += 4, p += 4 * ld_qkv) cp_async_16(sw_addr(sQ + which * TILE, row, lchunk), p);
  }
  {
    const __nv_bfloat16* pd = d_out + (static_cast<int64_t>(seq) * T + lrow) * ld_do + head * HD + lchunk * 8;
    const __nv_bfloat16* po = out + (static_cast<int64_t>(seq) * T + lrow) * ld_out + head * HD + lchunk * 8;
#pragma unroll 4
    for (int row = lrow; row < T; row += 4, pd += 4 * ld_do, po += 4 * ld_out) {
      cp_async_16(sw_addr(sdO, row, lchunk), pd);
      cp_async_16(sw_addr(sO, row, lchunk), po);
    }
  }
  cp_async_commit();
#pragma unroll
  for (int which = 0; which < 5; ++which)   // zero the padding rows of all five tiles
    for (int row = T + lrow; row < T_PAD; row += 4)
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sw_addr(sQ + which * TILE, row, lchunk)), "r"(0) : "memory");
  cp_async_wait<0>();
  __syncwarp();

  // ---- D_i = dO_i . O_i  (8 lanes per row)
  for (int r0 = 0; r0 < T_PAD; r0 += 4) {
    const int row = r0 + (lane >> 3), chunk = lane & 7;
    float s = dot8(lds128(sw_addr(sdO, row, chunk)), lds128(sw_addr(sO, row, chunk)));
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (chunk == 0) s_D[row] = s;
  }
  // ---- lse_i (base 2) of the masked, scaled scores: one pass over all key tiles per 16-query tile
#pragma unroll 1
  for (int m0 = 0; m0 < T_PAD; m0 += 16) {
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) load_a_frag(sQ, m0, ks, qa[ks]);
    float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
    float s[T_PAD / 8][4];
#pragma unroll
    for (int np = 0; np < T_PAD / 16; ++np) {
#pragma unroll
      for (int c = 0; c < 4; ++c) s[2 * np][c] = s[2 * np + 1][c] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b[4];
        load_bk_frag(sK, np * 16, ks, b);
        mma_bf16_16816(s[2 * np], qa[ks], b[0], b[1]);
        mma_bf16_16816(s[2 * np + 1], qa[ks], b[2], b[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < T_PAD / 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = m0 + g + (c >> 1) * 8;
        const int j = nt * 8 + tq * 2 + (c & 1);
        const bool ok = (j < T) && (causal_diag < 0 || j <= i + causal_diag);
        const float v = ok ? s[nt][c] * scale_log2 : -INFINITY;
        s[nt][c] = v;
        mx[c >> 1] = fmaxf(mx[c >> 1], v);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
#pragma unroll
    for (int nt = 0; nt < T_PAD / 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) sum[c >> 1] += ex2f(s[nt][c] - mx[c >> 1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
      if (tq == 0) s_lse[m0 + g + h * 8] = mx[h] + log2f(sum[h]);
    }
  }
  __syncwarp();

  // ---- dQ, then dK / dV, 16 rows at a time; staged in the (consumed) O tile, written as 128-byte rows
  __nv_bfloat16* dst = d_qkv + static_cast<int64_t>(seq) * T * ld_dqkv + head * HD;
  auto flush = [&](int r0, int which) {
    __syncwarp();
    for (int idx = lane; idx < 16 * 8; idx += 32) {
      const int row = idx >> 3, chunk = idx & 7;
      if (r0 + row < T)
        *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(r0 + row) * ld_dqkv + which * D + chunk * 8) =
            lds128(sw_addr(sO, row, chunk));
    }
    __syncwarp();
  };
#pragma unroll 1
  for (int m0 = 0; m0 < T; m0 += 16) {
    float dq[8][4];
    attn_bwd_dq_tile(sQ, sK, sV, sdO, s_lse, s_D, m0, T, causal_diag, scale_log2, scale, dq);
    stage_tile(sO, dq);
    flush(m0, 0);
  }
#pragma unroll 1
  for (int j0 = 0; j0 < T; j0 += 16) {
    float dk[8][4], dv[8][4];
    attn_bwd_dkv_tile(sQ, sK, sV, sdO, s_lse, s_D, j0, T, causal_diag, scale_log2, scale, dk, dv);
    stage_tile(sO, dk);
    flush(j0, 1);
    stage_tile(sO, dv);
    flush(j0, 2);
  }
}

}  // namespace tcow

extern "C" int tcow_attn_temporal_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const void* d_out,
                                      int64_t ld_do, void* d_qkv, int64_t ld_dqkv, int num_seq, int T, int heads,
                                      int causal_diag, void* stream) {
  using namespace tcow;
  if (!qkv || !out || !d_out || !d_qkv || num_seq <= 0 || T <= 0 || heads <= 0)
    return set_error(TCOW_ERR_ARG, "attn_temporal_bwd: bad argument");
  if (T > 64) return set_error(TCOW_ERR_ARG, "attn_temporal_bwd: T = %d > 64 not supported", T);
  if ((ld_qkv % 8) || (ld_out % 8) || (ld_do % 8) || (ld_dqkv % 8)) return set_error(TCOW_ERR_ARG, "attn_temporal_bwd: pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(num_seq) * heads;
  const int blocks = static_cast<int>((work + 3) / 4);
  const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
  auto go = [&](auto kern, int t_pad) -> int {
    const int smem = 4 * (5 * t_pad * ROW_BYTES + 2 * t_pad * 4);
    static bool configured[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[t_pad == 64][dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured[t_pad == 64][dev & 63] = true;
    }
    kern<<<blocks, 128, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<const __nv_bfloat16*>(out), ld_out,
                                   static_cast<const __nv_bfloat16*>(d_out), ld_do, static_cast<__nv_bfloat16*>(d_qkv), ld_dqkv,
                                   num_seq, T, heads, causal_diag, scale_log2, scale);
    return check_launch("attn_temporal_bwd_kernel");
  };
  return T <= 32 ? go(attn_temporal_bwd_kernel<32>, 32) : go(attn_temporal_bwd_kernel<64>, 64);
}

extern "C" int tcow_attn_spatial_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* out_cls,
                                     const void* d_out, int64_t ld_do, const float* d_out_cls, const float* lse, void* d_qkv,
                                     int64_t ld_dqkv, float* d_cls, int B, int N, int T, int heads, int use_cls,
                                     int64_t cls_row0, void* stream) {
  using namespace tcow;
  if (!qkv || !out || !d_out || !lse || !d_qkv || B <= 0 || N <= 0 || T <= 0 || heads <= 0)
    return set_error(TCOW_ERR_ARG, "attn_spatial_bwd: bad argument");
  if (use_cls && (!out_cls || !d_out_cls)) return set_error(TCOW_ERR_ARG, "attn_spatial_bwd: cls buffers missing");
  if (!d_cls) return set_error(TCOW_ERR_ARG, "attn_spatial_bwd: scratch missing");
  if (N + (use_cls ? 1 : 0) > 304)
    return set_error(TCOW_ERR_ARG, "attn_spatial_bwd: %d tokens per frame > 304 not supported in training", N + (use_cls ? 1 : 0));
  if ((ld_qkv % 8) || (ld_out % 8) || (ld_do % 8) || (ld_dqkv % 8)) return set_error(TCOW_ERR_ARG, "attn_spatial_bwd: pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tcgen05/TMEM kernels (attn_spatial_bwd_tc.cu): pass Q with the queries on the TMEM lanes, pass KV with the keys
  return launch_spatial_bwd_tc(qkv, ld_qkv, out, ld_out, out_cls, d_out, ld_do, d_out_cls, lse, d_qkv, ld_dqkv, d_cls,
                               d_cls + static_cast<int64_t>(B) * T * 3 * heads * 64, B, N, T, heads, use_cls ? 1 : 0,
                               cls_row0, s);
}
