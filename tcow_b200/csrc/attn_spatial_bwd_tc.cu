// Backward of the spatial attention on the 5th-gen tensor cores (tcgen05 + TMEM); two persistent kernels shaped like the
// round-1 forward kernel (one CTA = one chain of MMA -> softmax -> MMA), two CTAs per SM each.  With P = softmax(Q K^T/8), dP = dO V^T,
// dS = P o (dP - D)/8, D_i = dO_i . O_i  (autograd of vit.py:88-109 as called at vit.py:186):
//
//  pass Q  (queries on the TMEM lanes; work item (b,t,head), K and V resident in shared memory, 128-query tiles):
//      per 96-key block:  S = Q K_j^T and dP = dO V_j^T (two SS MMAs into TMEM columns [0,96) / [96,192))
//                         -> 128 threads, one query row each: p = 2^(s c - lse), dS -> bf16 over the S columns
//                         -> dQ += dS K_j (TS MMA, A = dS from TMEM, K as an MN-major operand), columns [192,256)
//  pass KV (keys on the lanes; Q and dO of the frame resident, 128-key tiles, 64-query blocks):
//      S^T = K Q_i^T, dP^T = V dO_i^T (columns [0,64) / [64,128)) -> one key row per thread, lse / D of the queries
//      broadcast from shared memory: P^T and dS^T (bf16) written back over the consumed columns
//      -> dV += P^T dO_i, dK += dS^T Q_i (TS MMAs, dO / Q as MN-major operands), columns [128,192) / [192,256)
// Nothing is transposed in memory, nothing of size S x S leaves the SM, no atomics (the cls token's per-frame gradients
// go to d_cls and are summed over the frames by cls_grad_reduce_tc_kernel).  Tested against the fp32 autograd of the same
// attention (tests/gpu_checks_train.py).
#include <math.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {

constexpr int BT_ROWS = 304;                  // tokens per frame, rounded up to 16
constexpr int BT_FULL_BYTES = BT_ROWS * 128;  // one resident [304][64] bf16 operand
constexpr int BT_TILE_BYTES = 128 * 128;      // one 128-row tile

struct BwdArgs {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  const __nv_bfloat16* out;
  int64_t ld_out;
  const float* out_cls;
  const __nv_bfloat16* d_out;
  int64_t ld_do;
  const float* d_out_cls;
  const float* lse;
  const float* dsum;   // D = dO . O per (item, token): written by attn_row_dot_kernel, laid out like lse
  __nv_bfloat16* d_qkv;
  int64_t ld_dqkv;
  float* d_cls;
  int B, N, T, heads, use_cls;
  int64_t cls_row0;
  float scale_log2, scale;
};

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bflo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bfhi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float dot8b(const uint4& a, const uint4& b) {
  return bflo(a.x) * bflo(b.x) + bfhi(a.x) * bfhi(b.x) + bflo(a.y) * bflo(b.y) + bfhi(a.y) * bfhi(b.y) +
         bflo(a.z) * bflo(b.z) + bfhi(a.z) * bfhi(b.z) + bflo(a.w) * bflo(b.w) + bfhi(a.w) * bfhi(b.w);
}
__device__ __forceinline__ void st_smem_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// D = dO . O and the base-2 lse of token `tok` of work item `item` (0 / 0 past S): two 4-byte loads (attn_row_dot_kernel
// has already reduced dO . O for every token and head).
__device__ __forceinline__ void row_stats(const BwdArgs& a, int item, int tok, int S, float& lse, float& Dv) {
  lse = 0.f;
  Dv = 0.f;
  if (tok >= S) return;
  lse = __ldg(a.lse + static_cast<int64_t>(item) * BT_ROWS + tok);
  Dv = __ldg(a.dsum + static_cast<int64_t>(item) * BT_ROWS + tok);
}

// cls token rows (token N of the frame) into row `r` of swizzled [rows][64] bf16 tiles: lanes 0-7 copy 16-byte chunks.
__device__ __forceinline__ void put_cls_qkv_row(const BwdArgs& a, int b, int h, int which, uint32_t tile, int r, int lane8) {
  const int D = a.heads * 64;
  const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + which * D + h * 64 + lane8 * 8);
  st_smem_v4(tile + r * 128 + ((lane8 ^ (r & 7)) << 4), v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void put_cls_do_row(const BwdArgs& a, int b, int t, int h, uint32_t tile, int r, int lane8) {
  const int D = a.heads * 64;
  const float4* d = reinterpret_cast<const float4*>(a.d_out_cls + (static_cast<int64_t>(b) * a.T + t) * D + h * 64 + lane8 * 8);
  const float4 x0 = __ldg(d), x1 = __ldg(d + 1);
  st_smem_v4(tile + r * 128 + ((lane8 ^ (r & 7)) << 4), pack_bf16(x0.x, x0.y), pack_bf16(x0.z, x0.w), pack_bf16(x1.x, x1.y),
             pack_bf16(x1.z, x1.w));
}

// 64 fp32 accumulator columns of this thread's TMEM lane -> bf16 row in global memory, or fp32 (cls token).
__device__ __forceinline__ void store_row64(uint32_t taddr, __nv_bfloat16* dst_bf16, float* dst_f32) {
  uint32_t o0[32], o1[32];
  tmem_ld_32x32(taddr, o0);
  tmem_ld_32x32(taddr + 32, o1);
  tmem_ld_wait();
  if (dst_f32) {
    float4* d = reinterpret_cast<float4*>(dst_f32);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] = make_float4(__uint_as_float(o0[4 * e]), __uint_as_float(o0[4 * e + 1]), __uint_as_float(o0[4 * e + 2]),
                         __uint_as_float(o0[4 * e + 3]));
      d[8 + e] = make_float4(__uint_as_float(o1[4 * e]), __uint_as_float(o1[4 * e + 1]), __uint_as_float(o1[4 * e + 2]),
                             __uint_as_float(o1[4 * e + 3]));
    }
  } else if (dst_bf16) {
    uint4* d = reinterpret_cast<uint4*>(dst_bf16);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[e] = make_uint4(pack_bf16(__uint_as_float(o0[8 * e]), __uint_as_float(o0[8 * e + 1])),
                        pack_bf16(__uint_as_float(o0[8 * e + 2]), __uint_as_float(o0[8 * e + 3])),
                        pack_bf16(__uint_as_float(o0[8 * e + 4]), __uint_as_float(o0[8 * e + 5])),
                        pack_bf16(__uint_as_float(o0[8 * e + 6]), __uint_as_float(o0[8 * e + 7])));
      d[4 + e] = make_uint4(pack_bf16(__uint_as_float(o1[8 * e]), __uint_as_float(o1[8 * e + 1])),
                            pack_bf16(__uint_as_float(o1[8 * e + 2]), __uint_as_float(o1[8 * e + 3])),
                            pack_bf16(__uint_as_float(o1[8 * e + 4]), __uint_as_float(o1[8 * e + 5])),
                            pack_bf16(__uint_as_float(o1[8 * e + 6]), __uint_as_float(o1[8 * e + 7])));
    }
  }
}

}  // namespace

// ============================================================================================ pass Q: dQ
constexpr int BQ_KB = 96;  // keys per block: S (96) + dP (96) + dQ (64) = 256 TMEM columns
constexpr int BQ_SMEM = 2 * BT_FULL_BYTES + 2 * BT_TILE_BYTES + 256 + 1024;

__global__ void __launch_bounds__(256, 2)
attn_spatial_bwd_q_kernel(const __grid_constant__ CUtensorMap tmQf, const __grid_constant__ CUtensorMap tmQt,
                          const __grid_constant__ CUtensorMap tmKVf, const __grid_constant__ CUtensorMap tmKVt,
                          const __grid_constant__ CUtensorMap tmDf, const __grid_constant__ CUtensorMap tmDt, const BwdArgs a) {
  extern __shared__ uint8_t smem_bq[];
  const uint32_t raw = smem_u32(smem_bq);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t k_buf = base, v_buf = base + BT_FULL_BYTES;
  const uint32_t q_buf = base + 2 * BT_FULL_BYTES, do_buf = q_buf + BT_TILE_BYTES;
  const uint32_t bars = do_buf + BT_TILE_BYTES;
  const uint32_t kv_full = bars, kv_empty = bars + 8, q_full = bars + 16, q_empty = bars + 24, s_full = bars + 32,
                 p_full = bars + 40, o_full = bars + 48, tmem_slot = bars + 56;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_bq + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int nblk = (S16 + BQ_KB - 1) / BQ_KB;
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int kv_full_rows = N < 256 ? N : 256, kv_tail_rows = N - kv_full_rows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQf); prefetch_tmap(&tmQt); prefetch_tmap(&tmKVf); prefetch_tmap(&tmKVt); prefetch_tmap(&tmDf);
    prefetch_tmap(&tmDt);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(q_full, 1); mbar_init(q_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // rows that TMA never writes must hold finite values: zero all operand buffers once
  for (int idx = threadIdx.x; idx < (2 * BT_FULL_BYTES + 2 * BT_TILE_BYTES) / 16; idx += blockDim.x)
    st_smem_v4(base + idx * 16, 0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t kv_it = 0, q_it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++kv_it) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      mbar_wait(kv_empty, (kv_it & 1) ^ 1);
      if (a.use_cls && lane < 16) put_cls_qkv_row(a, b, h, 1 + (lane >> 3), (lane >> 3) ? v_buf : k_buf, N, lane & 7);
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2u * N * 128u);
        tma_load_4d(k_buf, &tmKVf, D + h * 64, t, 0, b, kv_full);
        tma_load_4d(v_buf, &tmKVf, 2 * D + h * 64, t, 0, b, kv_full);
        if (kv_tail_rows > 0) {
          tma_load_4d(k_buf + 256 * 128, &tmKVt, D + h * 64, t, 256, b, kv_full);
          tma_load_4d(v_buf + 256 * 128, &tmKVt, 2 * D + h * 64, t, 256, b, kv_full);
        }
      }
      __syncwarp();
      for (int j = 0; j < nq; ++j, ++q_it) {
        mbar_wait(q_empty, (q_it & 1) ^ 1);
        const int rows = (N - 128 * j) < 128 ? (N - 128 * j) : 128;
        if (a.use_cls && (N >> 7) == j) {  // the cls query (token N) and its dO row
          if (lane < 8) put_cls_qkv_row(a, b, h, 0, q_buf, N - 128 * j, lane);
          else if (lane < 16) put_cls_do_row(a, b, t, h, do_buf, N - 128 * j, lane - 8);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (rows > 0) {
            mbar_expect_tx(q_full, 2u * static_cast<uint32_t>(rows) * 128u);
            tma_load_4d(q_buf, rows == 128 ? &tmQf : &tmQt, h * 64, t, 128 * j, b, q_full);
            tma_load_4d(do_buf, rows == 128 ? &tmDf : &tmDt, h * 64, t, 128 * j, b, q_full);
          } else {
            mbar_arrive(q_full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, 1);
    uint32_t kv_it = 0, q_it = 0, p_ct = 0;
    auto issue_s_dp = [&](int blk) {  // S = Q K_blk^T -> columns [0,96); dP = dO V_blk^T -> columns [96,192)
      const int nkb = (S16 - blk * BQ_KB) < BQ_KB ? (S16 - blk * BQ_KB) : BQ_KB;
      const uint32_t idesc = umma_idesc_bf16(128, nkb);
      const uint64_t qd = umma_desc_k_sw128(q_buf), dd = umma_desc_k_sw128(do_buf);
      const uint64_t kd = umma_desc_k_sw128(k_buf + blk * BQ_KB * 128), vd = umma_desc_k_sw128(v_buf + blk * BQ_KB * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd + 2u * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + BQ_KB, dd + 2u * k, vd + 2u * k, idesc, k > 0 ? 1u : 0u);
      umma_commit(s_full);
    };
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++kv_it) {
      mbar_wait(kv_full, kv_it & 1);
      for (int j = 0; j < nq; ++j, ++q_it) {
        mbar_wait(q_full, q_it & 1);
        tc_fence_after();
        if (elect_one()) issue_s_dp(0);
        __syncwarp();
        for (int blk = 0; blk < nblk; ++blk) {
          const int nkb = (S16 - blk * BQ_KB) < BQ_KB ? (S16 - blk * BQ_KB) : BQ_KB;
          mbar_wait(p_full, p_ct & 1);
          ++p_ct;
          tc_fence_after();
          if (elect_one()) {
            // dQ (+)= dS_blk K_blk : 16 keys per MMA (8 TMEM columns of dS, 16 rows = 2048 B of K, MN-major)
            const uint64_t kd = umma_desc_mn_sw128(k_buf + blk * BQ_KB * 128, 1024);
            for (int kk = 0; kk < (nkb >> 4); ++kk)
              umma_bf16_ts(tmem_base + 2 * BQ_KB, tmem_base + 8u * kk, kd + 128u * kk, idesc_dq, (blk | kk) != 0 ? 1u : 0u);
            if (blk + 1 < nblk) {
              issue_s_dp(blk + 1);
            } else {
              umma_commit(q_empty);
              umma_commit(o_full);
              if (j == nq - 1) umma_commit(kv_empty);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ dS + output (one query row per thread)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = a.scale_log2, log2_scale = log2f(a.scale);
    uint32_t s_ct = 0, o_ct = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      for (int j = 0; j < nq; ++j, ++o_ct) {
        const int tok = 128 * j + row;
        float lse, Dv;
        row_stats(a, item, tok, S, lse, Dv);
        const float lse_s = lse - log2_scale;  // folds the 1/sqrt(hd) factor of dS into the exponent
        const uint64_t sc2 = f2_pack(sc, sc), nl2 = f2_pack(-lse_s, -lse_s), nD2 = f2_pack(-Dv, -Dv);
        for (int blk = 0; blk < nblk; ++blk, ++s_ct) {
          const int nkb = (S16 - blk * BQ_KB) < BQ_KB ? (S16 - blk * BQ_KB) : BQ_KB;
          mbar_wait(s_full, s_ct & 1);
          tc_fence_after();
          for (int c = 0; c * 32 < nkb; ++c) {
            uint32_t vs[32], vp[32], pk[16];
            tmem_ld_32x32(t_lane + 32 * c, vs);
            tmem_ld_32x32(t_lane + BQ_KB + 32 * c, vp);
            tmem_ld_wait();
            const int key0 = blk * BQ_KB + 32 * c;
            if (key0 + 32 <= S) {
              // packed fp32x2: dS = 2^(s c - lse + log2(scale)) * (dP - D): 3 issue slots per element pair + 2 MUFU
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                float x0, x1;
                f2_unpack(f2_fma(f2_pack_u(vs[e], vs[e + 1]), sc2, nl2), x0, x1);
                const uint64_t p2 = f2_pack(ex2a(x0), ex2a(x1));
                float d0, d1;
                f2_unpack(f2_mul(p2, f2_add(f2_pack_u(vp[e], vp[e + 1]), nD2)), d0, d1);
                pk[e >> 1] = pack_bf16(d0, d1);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                const bool ok0 = key0 + e < S, ok1 = key0 + e + 1 < S;
                const float p0 = ex2a(fmaf(__uint_as_float(vs[e]), sc, -lse_s));
                const float p1 = ex2a(fmaf(__uint_as_float(vs[e + 1]), sc, -lse_s));
                const float d0 = ok0 ? p0 * (__uint_as_float(vp[e]) - Dv) : 0.f;
                const float d1 = ok1 ? p1 * (__uint_as_float(vp[e + 1]) - Dv) : 0.f;
                pk[e >> 1] = pack_bf16(d0, d1);
              }
            }
            tmem_st_32x16(t_lane + 16 * c, pk);  // dS over S columns already consumed
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full);
        }
        mbar_wait(o_full, o_ct & 1);
        tc_fence_after();
        __nv_bfloat16* dst = nullptr;
        float* dst_cls = nullptr;
        if (tok < N) dst = a.d_qkv + ((static_cast<int64_t>(b) * N + tok) * T + t) * a.ld_dqkv + h * 64;
        else if (tok == N && a.use_cls) dst_cls = a.d_cls + ((static_cast<int64_t>(b) * T + t) * 3 + 0) * D + h * 64;
        store_row64(t_lane + 2 * BQ_KB, dst, dst_cls);
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ============================================================================================ pass KV: dK, dV
constexpr int BK_QB = 64;  // queries per block: S^T (64) + dP^T (64) + dV (64) + dK (64) = 256 TMEM columns
constexpr int BK_STAT = 320;  // per-query lse / D slots (S16 rounded up to the 64-query block)
constexpr int BK_SMEM = 2 * BT_FULL_BYTES + 2 * BT_TILE_BYTES + 2 * BK_STAT * 4 + 256 + 1024;

__global__ void __launch_bounds__(256, 2)
attn_spatial_bwd_kv_kernel(const __grid_constant__ CUtensorMap tmQf, const __grid_constant__ CUtensorMap tmQt,
                           const __grid_constant__ CUtensorMap tmKVf, const __grid_constant__ CUtensorMap tmKVt,
                           const __grid_constant__ CUtensorMap tmDf, const __grid_constant__ CUtensorMap tmDt, const BwdArgs a) {
  // tmQf/tmQt: qkv, 128-row / (N % 128)-row boxes (K and V tiles); tmKVf/tmKVt: qkv, 256-row / (N-256)-row boxes (resident Q);
  // tmDf/tmDt: d_out, 256-row / (N-256)-row boxes (resident dO)
  extern __shared__ uint8_t smem_bk[];
  const uint32_t raw = smem_u32(smem_bk);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t q_buf = base, do_buf = base + BT_FULL_BYTES;
  const uint32_t k_buf = base + 2 * BT_FULL_BYTES, v_buf = k_buf + BT_TILE_BYTES;
  const uint32_t stats = v_buf + BT_TILE_BYTES;
  float* s_lse = reinterpret_cast<float*>(smem_bk + (stats - raw));
  float* s_D = s_lse + BK_STAT;
  const uint32_t bars = stats + 2 * BK_STAT * 4;
  const uint32_t qdo_full = bars, qdo_empty = bars + 8, kv_full = bars + 16, kv_empty = bars + 24, s_full = bars + 32,
                 p_full = bars + 40, acc_full = bars + 48, acc_empty = bars + 56, tmem_slot = bars + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_bk + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nkt = (S + 127) >> 7;               // 128-key tiles
  const int nqb = (S16 + BK_QB - 1) / BK_QB;    // 64-query blocks
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int res_full_rows = N < 256 ? N : 256, res_tail_rows = N - res_full_rows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQf); prefetch_tmap(&tmQt); prefetch_tmap(&tmKVf); prefetch_tmap(&tmKVt); prefetch_tmap(&tmDf);
    prefetch_tmap(&tmDt);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1); mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(acc_full, 1); mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  for (int idx = threadIdx.x; idx < (2 * BT_FULL_BYTES + 2 * BT_TILE_BYTES) / 16; idx += blockDim.x)
    st_smem_v4(base + idx * 16, 0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t it = 0, kv_it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      mbar_wait(qdo_empty, (it & 1) ^ 1);
      if (a.use_cls) {  // token N of the resident Q / dO
        if (lane < 8) put_cls_qkv_row(a, b, h, 0, q_buf, N, lane);
        else if (lane < 16) put_cls_do_row(a, b, t, h, do_buf, N, lane - 8);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(qdo_full, 2u * N * 128u);
        tma_load_4d(q_buf, &tmKVf, h * 64, t, 0, b, qdo_full);
        tma_load_4d(do_buf, &tmDf, h * 64, t, 0, b, qdo_full);
        if (res_tail_rows > 0) {
          tma_load_4d(q_buf + 256 * 128, &tmKVt, h * 64, t, 256, b, qdo_full);
          tma_load_4d(do_buf + 256 * 128, &tmDt, h * 64, t, 256, b, qdo_full);
        }
      }
      __syncwarp();
      for (int kt = 0; kt < nkt; ++kt, ++kv_it) {
        mbar_wait(kv_empty, (kv_it & 1) ^ 1);
        const int rows = (N - 128 * kt) < 128 ? (N - 128 * kt) : 128;
        if (a.use_cls && (N >> 7) == kt && lane < 16)
          put_cls_qkv_row(a, b, h, 1 + (lane >> 3), (lane >> 3) ? v_buf : k_buf, N - 128 * kt, lane & 7);
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (rows > 0) {
            mbar_expect_tx(kv_full, 2u * static_cast<uint32_t>(rows) * 128u);
            const CUtensorMap* m = rows == 128 ? &tmQf : &tmQt;
            tma_load_4d(k_buf, m, D + h * 64, t, 128 * kt, b, kv_full);
            tma_load_4d(v_buf, m, 2 * D + h * 64, t, 128 * kt, b, kv_full);
          } else {
            mbar_arrive(kv_full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 64, 1);
    uint32_t it = 0, kv_it = 0, p_ct = 0;
    auto issue_st = [&](int qb) {  // S^T = K Q_qb^T -> columns [0,64); dP^T = V dO_qb^T -> columns [64,128)
      const int n = (S16 - qb * BK_QB) < BK_QB ? (S16 - qb * BK_QB) : BK_QB;
      const uint32_t idesc = umma_idesc_bf16(128, n);
      const uint64_t kd = umma_desc_k_sw128(k_buf), vd = umma_desc_k_sw128(v_buf);
      const uint64_t qd = umma_desc_k_sw128(q_buf + qb * BK_QB * 128), dd = umma_desc_k_sw128(do_buf + qb * BK_QB * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, kd + 2u * k, qd + 2u * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + BK_QB, vd + 2u * k, dd + 2u * k, idesc, k > 0 ? 1u : 0u);
      umma_commit(s_full);
    };
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      mbar_wait(qdo_full, it & 1);
      for (int kt = 0; kt < nkt; ++kt, ++kv_it) {
        mbar_wait(kv_full, kv_it & 1);
        tc_fence_after();
        if (elect_one()) issue_st(0);
        __syncwarp();
        mbar_wait(acc_empty, (kv_it & 1) ^ 1);  // the previous tile's dK / dV have been read out of TMEM
        for (int qb = 0; qb < nqb; ++qb) {
          const int n = (S16 - qb * BK_QB) < BK_QB ? (S16 - qb * BK_QB) : BK_QB;
          mbar_wait(p_full, p_ct & 1);
          ++p_ct;
          tc_fence_after();
          if (elect_one()) {
            // dV (+)= P^T dO_qb, dK (+)= dS^T Q_qb : 16 queries per MMA (8 TMEM columns, 16 rows = 2048 B of dO / Q)
            const uint64_t dd = umma_desc_mn_sw128(do_buf + qb * BK_QB * 128, 1024);
            const uint64_t qd = umma_desc_mn_sw128(q_buf + qb * BK_QB * 128, 1024);
            for (int kk = 0; kk < (n >> 4); ++kk)
              umma_bf16_ts(tmem_base + 128, tmem_base + 8u * kk, dd + 128u * kk, idesc_acc, (qb | kk) != 0 ? 1u : 0u);
            for (int kk = 0; kk < (n >> 4); ++kk)
              umma_bf16_ts(tmem_base + 192, tmem_base + BK_QB + 8u * kk, qd + 128u * kk, idesc_acc, (qb | kk) != 0 ? 1u : 0u);
            if (qb + 1 < nqb) {
              issue_st(qb + 1);
            } else {
              umma_commit(kv_empty);
              umma_commit(acc_full);
              if (kt == nkt - 1) umma_commit(qdo_empty);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ P^T / dS^T + output (one key row per thread)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = a.scale_log2, scale = a.scale;
    const uint64_t sc2 = f2_pack(sc, sc), scale2 = f2_pack(scale, scale);
    uint32_t s_ct = 0, acc_ct = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      // per-query statistics of this frame -> shared memory (queries past S: -lse = -inf -> p = 0)
      for (int q = row; q < BK_STAT; q += 128) {
        float lse, Dv;
        row_stats(a, item, q, S, lse, Dv);
        s_lse[q] = q < S ? -lse : -INFINITY;   // stored negated: the addend of the exponent FMA
        s_D[q] = -Dv * scale;                  // and of the (dP scale - D scale) FMA
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int kt = 0; kt < nkt; ++kt, ++acc_ct) {
        const int key = 128 * kt + row;
        const bool kvalid = key < S;
        for (int qb = 0; qb < nqb; ++qb, ++s_ct) {
          const int n = (S16 - qb * BK_QB) < BK_QB ? (S16 - qb * BK_QB) : BK_QB;
          mbar_wait(s_full, s_ct & 1);
          tc_fence_after();
          for (int c = 0; c * 32 < n; ++c) {
            uint32_t vs[32], vp[32], pp[16], pd[16];
            tmem_ld_32x32(t_lane + 32 * c, vs);
            tmem_ld_32x32(t_lane + BK_QB + 32 * c, vp);
            tmem_ld_wait();
            const int q0 = qb * BK_QB + 32 * c;
            if (kvalid) {
              // packed fp32x2: P^T = 2^(s c - lse_q), dS^T = P^T * (dP^T scale - D_q scale); -lse_q and -D_q*scale come
              // from shared memory (same address in every lane: broadcast)
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(s_lse + q0 + e);
                const float4 d4 = *reinterpret_cast<const float4*>(s_D + q0 + e);
                float x0, x1, x2, x3;
                f2_unpack(f2_fma(f2_pack_u(vs[e], vs[e + 1]), sc2, f2_pack(l4.x, l4.y)), x0, x1);
                f2_unpack(f2_fma(f2_pack_u(vs[e + 2], vs[e + 3]), sc2, f2_pack(l4.z, l4.w)), x2, x3);
                const float p0 = ex2a(x0), p1 = ex2a(x1), p2 = ex2a(x2), p3 = ex2a(x3);
                float r0, r1, r2, r3;
                f2_unpack(f2_mul(f2_pack(p0, p1), f2_fma(f2_pack_u(vp[e], vp[e + 1]), scale2, f2_pack(d4.x, d4.y))), r0, r1);
                f2_unpack(f2_mul(f2_pack(p2, p3), f2_fma(f2_pack_u(vp[e + 2], vp[e + 3]), scale2, f2_pack(d4.z, d4.w))), r2, r3);
                pp[e >> 1] = pack_bf16(p0, p1);
                pp[(e >> 1) + 1] = pack_bf16(p2, p3);
                pd[e >> 1] = pack_bf16(r0, r1);
                pd[(e >> 1) + 1] = pack_bf16(r2, r3);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) pp[e] = pd[e] = 0u;
            }
            tmem_st_32x16(t_lane + 16 * c, pp);            // P^T over consumed S^T columns
            tmem_st_32x16(t_lane + BK_QB + 16 * c, pd);    // dS^T over consumed dP^T columns
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full);
        }
        mbar_wait(acc_full, acc_ct & 1);
        tc_fence_after();
        __nv_bfloat16 *dv_dst = nullptr, *dk_dst = nullptr;
        float *dv_cls = nullptr, *dk_cls = nullptr;
        if (key < N) {
          __nv_bfloat16* r = a.d_qkv + ((static_cast<int64_t>(b) * N + key) * T + t) * a.ld_dqkv + h * 64;
          dk_dst = r + D;
          dv_dst = r + 2 * D;
        } else if (key == N && a.use_cls) {
          float* r = a.d_cls + (static_cast<int64_t>(b) * T + t) * 3 * D + h * 64;
          dk_cls = r + D;
          dv_cls = r + 2 * D;
        }
        store_row64(t_lane + 128, dv_dst, dv_cls);
        store_row64(t_lane + 192, dk_dst, dk_cls);
        tc_fence_before();
        mbar_arrive(acc_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// D[(b*T+t)*heads+h][tok] = sum_d dO[tok, h*64+d] * O[tok, h*64+d] for every patch token (one warp per token row: 96
// 16-byte vectors, 8 lanes per head) and, in the last blocks, for the cls token (fp32 O / dO, dO rounded to bf16 as the
// tensor cores see it).  Both backward passes then read D like the lse.
__global__ void __launch_bounds__(256) attn_row_dot_kernel(const BwdArgs a, float* __restrict__ dsum) {
  const int lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads, D = heads * 64;
  const int64_t M = static_cast<int64_t>(a.B) * N * T;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row < M) {
    const int t = static_cast<int>(row % T), n = static_cast<int>((row / T) % N), b = static_cast<int>(row / (static_cast<int64_t>(T) * N));
    const uint4* o = reinterpret_cast<const uint4*>(a.out + row * a.ld_out);
    const uint4* d = reinterpret_cast<const uint4*>(a.d_out + row * a.ld_do);
    for (int v = lane; v < D / 8; v += 32) {
      float s = dot8b(__ldg(d + v), __ldg(o + v));
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((lane & 7) == 0) dsum[((static_cast<int64_t>(b) * T + t) * heads + (v >> 3)) * BT_ROWS + n] = s;
    }
  } else if (a.use_cls) {
    const int64_t c = row - M;  // (b*T + t)*heads + h
    if (c < static_cast<int64_t>(a.B) * T * heads) {
      const int h = static_cast<int>(c % heads);
      const int64_t bt = c / heads;
      const float2 ov = __ldg(reinterpret_cast<const float2*>(a.out_cls + bt * D + h * 64) + lane);
      const float2 dv = __ldg(reinterpret_cast<const float2*>(a.d_out_cls + bt * D + h * 64) + lane);
      const uint32_t p = pack_bf16(dv.x, dv.y);
      float s = bflo(p) * ov.x + bfhi(p) * ov.y;
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
      if (lane == 0) dsum[c * BT_ROWS + N] = s;
    }
  }
}

// d_qkv[cls_row0+b, :] (bf16) = sum_t d_cls[b,t,:]   (3*D columns)
__global__ void cls_grad_reduce_tc_kernel(const float* __restrict__ d_cls, __nv_bfloat16* __restrict__ d_qkv, int64_t ld, int T,
                                          int cols, int64_t cls_row0) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += d_cls[(static_cast<int64_t>(b) * T + t) * cols + c];
    d_qkv[(cls_row0 + b) * ld + c] = __float2bfloat16(s);
  }
}

int launch_spatial_bwd_tc(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* out_cls,
                          const void* d_out, int64_t ld_do, const float* d_out_cls, const float* lse, void* d_qkv,
                          int64_t ld_dqkv, float* d_cls, float* dsum, int B, int N, int T, int heads, int use_cls,
                          int64_t cls_row0, cudaStream_t stream) {
  alignas(64) CUtensorMap q128, q128t, q256, q256t, d128, d128t, d256, d256t;
  const int cols = 3 * heads * 64, dcols = heads * 64;
  const int t128 = N % 128, f256 = N < 256 ? N : 256, t256 = N - f256;
  int rc;
  if ((rc = make_patch_tmap(&q128, qkv, ld_qkv, cols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&q128t, qkv, ld_qkv, cols, B, N, T, t128 > 0 ? t128 : 1))) return rc;
  if ((rc = make_patch_tmap(&q256, qkv, ld_qkv, cols, B, N, T, f256))) return rc;
  if ((rc = make_patch_tmap(&q256t, qkv, ld_qkv, cols, B, N, T, t256 > 0 ? t256 : 1))) return rc;
  if ((rc = make_patch_tmap(&d128, d_out, ld_do, dcols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&d128t, d_out, ld_do, dcols, B, N, T, t128 > 0 ? t128 : 1))) return rc;
  if ((rc = make_patch_tmap(&d256, d_out, ld_do, dcols, B, N, T, f256))) return rc;
  if ((rc = make_patch_tmap(&d256t, d_out, ld_do, dcols, B, N, T, t256 > 0 ? t256 : 1))) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_spatial_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const BwdArgs a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<const __nv_bfloat16*>(out), ld_out, out_cls,
                  static_cast<const __nv_bfloat16*>(d_out), ld_do, d_out_cls, lse, dsum, static_cast<__nv_bfloat16*>(d_qkv),
                  ld_dqkv, d_cls, B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, 0.125f};
  const int items = B * T * heads;
  {
    const int64_t rows = static_cast<int64_t>(B) * N * T + (use_cls ? static_cast<int64_t>(items) : 0);
    attn_row_dot_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(a, dsum);
    rc = check_launch("attn_row_dot_kernel");
    if (rc) return rc;
  }
  const int slots = 2 * sm_count();
  const int grid = items < slots ? items : slots;
  attn_spatial_bwd_q_kernel<<<grid, 256, BQ_SMEM, stream>>>(q128, q128t, q256, q256t, d128, d128t, a);
  rc = check_launch("attn_spatial_bwd_q_kernel");
  if (rc) return rc;
  attn_spatial_bwd_kv_kernel<<<grid, 256, BK_SMEM, stream>>>(q128, q128t, q256, q256t, d256, d256t, a);
  rc = check_launch("attn_spatial_bwd_kv_kernel");
  if (rc) return rc;
  if (use_cls) {
    cls_grad_reduce_tc_kernel<<<B, 256, 0, stream>>>(d_cls, static_cast<__nv_bfloat16*>(d_qkv), ld_dqkv, T, 3 * heads * 64, cls_row0);
    rc = check_launch("cls_grad_reduce_kernel");
  }
  return rc;
}

}  // namespace tcow
