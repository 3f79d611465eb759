// Spatial attention on the 5th-gen tensor cores (tcgen05 + TMEM), persistent, one CTA per SM.
//
// Work item = (clip b, frame t, head h): full softmax attention over S = N (+1 cls) tokens of head dim 64
// (vit.py:78-111 as called at vit.py:186 on the tokens assembled at vit.py:179-185).  Per item:
//   K, V [S,64] bf16 -> smem (TMA 4-D gather of the strided canonical rows, SWIZZLE_128B; cls row appended by
//                     the producer warp); double buffered across items
//   per 128-query tile:  Q tile -> smem ring (TMA)                      S = Q K^T  : tcgen05.mma SS, fp32 in TMEM
//                        softmax : 128 threads, one query row each, two passes over TMEM (max, then exp2/sum),
//                                  P (bf16) written back over S with tcgen05.st          — no shuffles needed
//                        O = P V : tcgen05.mma with A = P from TMEM, B = V as an MN-major smem operand
//                        O / l  -> bf16 -> canonical output rows (cls query -> out_cls fp32)
// Warps: 0 TMA producer (+cls rows), 1 MMA issuer, 2 TMEM allocator, 4-7 softmax/epilogue.
// TMEM: S/P columns [0,320), O columns [320,384).   Limits: S <= 304 (else the mma.sync kernel is used).
#include <math.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

constexpr int SP_ROWS = 304;                  // K/V rows per stage (S rounded up to 16)
constexpr int SP_KV_BYTES = SP_ROWS * 128;    // 38912 (multiple of 1024)
constexpr int SP_STAGE_BYTES = 2 * SP_KV_BYTES;
constexpr int SP_QTILE_BYTES = 128 * 128;
constexpr int SP_QSLOTS = 3;
constexpr int SP_SMEM = 2 * SP_STAGE_BYTES + SP_QSLOTS * SP_QTILE_BYTES + 256 + 1024;
constexpr int SP_TMEM_O = 320;

struct SpatialArgs {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  __nv_bfloat16* out;
  int64_t ld_out;
  float* out_cls;
  int B, N, T, heads, use_cls;
  int64_t cls_row0;
  float scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 1)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tmQfull, const __grid_constant__ CUtensorMap tmQtail,
                       const __grid_constant__ CUtensorMap tmKVfull, const __grid_constant__ CUtensorMap tmKVtail,
                       const SpatialArgs a) {
  extern __shared__ uint8_t smem_sp[];
  const uint32_t raw = smem_u32(smem_sp);
  const uint32_t base = (raw + 1023u) & ~1023u;
  auto k_buf = [&](int st) { return base + st * SP_STAGE_BYTES; };
  auto v_buf = [&](int st) { return base + st * SP_STAGE_BYTES + SP_KV_BYTES; };
  auto q_buf = [&](int slot) { return base + 2 * SP_STAGE_BYTES + slot * SP_QTILE_BYTES; };
  const uint32_t bars = base + 2 * SP_STAGE_BYTES + SP_QSLOTS * SP_QTILE_BYTES;
  auto kv_full = [&](int s) { return bars + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 8u * (2 + s); };
  auto q_full = [&](int s) { return bars + 8u * (4 + s); };
  auto q_empty = [&](int s) { return bars + 8u * (7 + s); };
  const uint32_t s_full = bars + 8u * 10, p_full = bars + 8u * 11, o_full = bars + 8u * 12;
  const uint32_t tmem_slot = bars + 8u * 13;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_sp + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int n1 = S16 < 256 ? S16 : 256, n2 = S16 - n1;
  const int kv_full_rows = N < 256 ? N : 256, kv_tail_rows = N - kv_full_rows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQfull);
    prefetch_tmap(&tmQtail);
    prefetch_tmap(&tmKVfull);
    prefetch_tmap(&tmKVtail);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < SP_QSLOTS; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // Rows [S, 304) of every K/V stage are never written by TMA: zero them once (P is 0 there, V must be finite).
  for (int idx = threadIdx.x; idx < (SP_ROWS - S) * 8 * 4; idx += blockDim.x) {
    const int bufi = idx / ((SP_ROWS - S) * 8), rem = idx % ((SP_ROWS - S) * 8);
    const int row = S + (rem >> 3), chunk = rem & 7;
    const uint32_t b0 = (bufi & 1) ? v_buf(bufi >> 1) : k_buf(bufi >> 1);
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(b0 + row * 128 + ((chunk ^ (row & 7)) << 4)), "r"(0) : "memory");
  }
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
      const int st = kv_it & 1;
      mbar_wait(kv_empty(st), ((kv_it >> 1) & 1) ^ 1);
      if (a.use_cls && lane < 16) {  // cls k / v rows -> row N of the K / V tiles
        const int which = lane >> 3, chunk = lane & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + (1 + which) * D + h * 64 + chunk * 8);
        const uint32_t dst = (which ? v_buf(st) : k_buf(st)) + N * 128 + ((chunk ^ (N & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(kv_full(st), 2u * N * 128u);
        tma_load_4d(k_buf(st), &tmKVfull, D + h * 64, t, 0, b, kv_full(st));
        tma_load_4d(v_buf(st), &tmKVfull, 2 * D + h * 64, t, 0, b, kv_full(st));
        if (kv_tail_rows > 0) {
          tma_load_4d(k_buf(st) + 256 * 128, &tmKVtail, D + h * 64, t, 256, b, kv_full(st));
          tma_load_4d(v_buf(st) + 256 * 128, &tmKVtail, 2 * D + h * 64, t, 256, b, kv_full(st));
        }
      }
      __syncwarp();
      for (int j = 0; j < nq; ++j, ++q_it) {
        const int slot = q_it % SP_QSLOTS;
        mbar_wait(q_empty(slot), ((q_it / SP_QSLOTS) & 1) ^ 1);
        const int rows = (N - 128 * j) < 128 ? (N - 128 * j) : 128;  // patch rows in this tile (may be <= 0)
        if (a.use_cls && (N >> 7) == j && lane < 8) {                 // the cls query is token N
          const int r = N - 128 * j;
          const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + h * 64 + lane * 8);
          const uint32_t dst = q_buf(slot) + r * 128 + ((lane ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (rows > 0) {
            mbar_expect_tx(q_full(slot), static_cast<uint32_t>(rows) * 128u);
            tma_load_4d(q_buf(slot), rows == 128 ? &tmQfull : &tmQtail, h * 64, t, 128 * j, b, q_full(slot));
          } else {
            mbar_arrive(q_full(slot));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s1 = umma_idesc_bf16(128, n1);
    const uint32_t idesc_s2 = umma_idesc_bf16(128, n2 > 0 ? n2 : 16);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    uint32_t kv_it = 0, q_it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++kv_it) {
      const int st = kv_it & 1;
      mbar_wait(kv_full(st), (kv_it >> 1) & 1);
      for (int j = 0; j < nq; ++j, ++q_it) {
        const int slot = q_it % SP_QSLOTS;
        mbar_wait(q_full(slot), (q_it / SP_QSLOTS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qd = umma_desc_k_sw128(q_buf(slot));
          const uint64_t kd = umma_desc_k_sw128(k_buf(st));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd + 2u * k, idesc_s1, k > 0 ? 1u : 0u);
          if (n2 > 0) {
            const uint64_t kd2 = umma_desc_k_sw128(k_buf(st) + 256 * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 256, qd + 2u * k, kd2 + 2u * k, idesc_s2, k > 0 ? 1u : 0u);
          }
          umma_commit(q_empty(slot));
          umma_commit(s_full);
        }
        __syncwarp();
        mbar_wait(p_full, q_it & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = umma_desc_mn_sw128(v_buf(st), 1024);
          for (int kk = 0; kk < (S16 >> 4); ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
            umma_bf16_ts(tmem_base + SP_TMEM_O, tmem_base + 8u * kk, vd + 128u * kk, idesc_o, kk > 0 ? 1u : 0u);
          umma_commit(o_full);
          if (j == nq - 1) umma_commit(kv_empty(st));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output (one query row per thread)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int nchunk = (S16 + 31) >> 5;
    const float sc = a.scale_log2;
    uint32_t tile_ct = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      for (int j = 0; j < nq; ++j, ++tile_ct) {
        const int tok = 128 * j + row;
        const bool valid = tok < S;
        mbar_wait(s_full, tile_ct & 1);
        tc_fence_after();
        // pass 1: row maximum
        float mx = -INFINITY;
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + 32 * c, v);
          tmem_ld_wait();
          if (valid) {
            if (32 * c + 32 <= S) {
#pragma unroll
              for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (32 * c + e < S) mx = fmaxf(mx, __uint_as_float(v[e]));
            }
          }
        }
        // pass 2: p = 2^(s*sc - mx*sc), row sum, P (bf16) over the S columns already consumed
        const float mxs = valid ? mx * sc : 0.f;
        float l = 0.f;
        for (int c = 0; c < nchunk; ++c) {
          uint32_t v[32], pk[16];
          tmem_ld_32x32(t_lane + 32 * c, v);
          tmem_ld_wait();
          if (valid) {
            const bool full = (32 * c + 32 <= S);
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), sc, -mxs));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc, -mxs));
              if (!full) {
                if (32 * c + e >= S) p0 = 0.f;
                if (32 * c + e + 1 >= S) p1 = 0.f;
              }
              l += p0 + p1;
              pk[e >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          }
          tmem_st_32x16(t_lane + 16 * c, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        // O
        mbar_wait(o_full, tile_ct & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + SP_TMEM_O, o0);
        tmem_ld_32x32(t_lane + SP_TMEM_O + 32, o1);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l;
          if (a.use_cls && tok == N) {
            float4* dst = reinterpret_cast<float4*>(a.out_cls + (static_cast<int64_t>(b) * T + t) * D + h * 64);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              dst[e] = make_float4(__uint_as_float(o0[4 * e]) * inv, __uint_as_float(o0[4 * e + 1]) * inv,
                                   __uint_as_float(o0[4 * e + 2]) * inv, __uint_as_float(o0[4 * e + 3]) * inv);
              dst[8 + e] = make_float4(__uint_as_float(o1[4 * e]) * inv, __uint_as_float(o1[4 * e + 1]) * inv,
                                       __uint_as_float(o1[4 * e + 2]) * inv, __uint_as_float(o1[4 * e + 3]) * inv);
            }
          } else {
            uint4* dst = reinterpret_cast<uint4*>(a.out + ((static_cast<int64_t>(b) * N + tok) * T + t) * a.ld_out + h * 64);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              dst[e] = make_uint4(pack_bf16(__uint_as_float(o0[8 * e]) * inv, __uint_as_float(o0[8 * e + 1]) * inv),
                                  pack_bf16(__uint_as_float(o0[8 * e + 2]) * inv, __uint_as_float(o0[8 * e + 3]) * inv),
                                  pack_bf16(__uint_as_float(o0[8 * e + 4]) * inv, __uint_as_float(o0[8 * e + 5]) * inv),
                                  pack_bf16(__uint_as_float(o0[8 * e + 6]) * inv, __uint_as_float(o0[8 * e + 7]) * inv));
              dst[4 + e] = make_uint4(pack_bf16(__uint_as_float(o1[8 * e]) * inv, __uint_as_float(o1[8 * e + 1]) * inv),
                                      pack_bf16(__uint_as_float(o1[8 * e + 2]) * inv, __uint_as_float(o1[8 * e + 3]) * inv),
                                      pack_bf16(__uint_as_float(o1[8 * e + 4]) * inv, __uint_as_float(o1[8 * e + 5]) * inv),
                                      pack_bf16(__uint_as_float(o1[8 * e + 6]) * inv, __uint_as_float(o1[8 * e + 7]) * inv));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// 4-D view of the patch rows of qkv: (column, t, n, b) -> ((b*N+n)*T+t)*ld + column; box = 64 columns x box_n tokens.
static int make_patch_tmap(CUtensorMap* m, const void* qkv, int64_t ld, int cols, int B, int N, int T, int box_n) {
  const uint64_t dims[4] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(T), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(T) * ld * 2,
                               static_cast<uint64_t>(N) * T * ld * 2};
  const uint32_t box[4] = {64, 1, static_cast<uint32_t>(box_n), 1};
  return make_tmap_nd(m, false, qkv, 4, dims, strides, box);
}

int launch_spatial_tc(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                      int heads, int use_cls, int64_t cls_row0, cudaStream_t stream) {
  alignas(64) CUtensorMap tmQf, tmQt, tmKVf, tmKVt;
  const int cols = 3 * heads * 64;
  const int q_tail = N % 128, kv_full = N < 256 ? N : 256, kv_tail = N - kv_full;
  int rc;
  if ((rc = make_patch_tmap(&tmQf, qkv, ld_qkv, cols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&tmQt, qkv, ld_qkv, cols, B, N, T, q_tail > 0 ? q_tail : 1))) return rc;
  if ((rc = make_patch_tmap(&tmKVf, qkv, ld_qkv, cols, B, N, T, kv_full))) return rc;
  if ((rc = make_patch_tmap(&tmKVt, qkv, ld_qkv, cols, B, N, T, kv_tail > 0 ? kv_tail : 1))) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  SpatialArgs a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out), ld_out, out_cls,
                B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f};
  const int items = B * T * heads;
  const int grid = items < sm_count() ? items : sm_count();
  attn_spatial_tc_kernel<<<grid, 256, SP_SMEM, stream>>>(tmQf, tmQt, tmKVf, tmKVt, a);
  return check_launch("attn_spatial_tc_kernel");
}

}  // namespace tcow
