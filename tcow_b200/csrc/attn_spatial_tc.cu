// Spatial attention on the 5th-gen tensor cores (tcgen05 + TMEM), persistent, TWO CTAs per SM.
//
// Work item = (clip b, frame t, head h): full softmax attention over S = N (+1 cls) tokens of head dim 64
// (vit.py:78-111 as called at vit.py:186 on the tokens assembled at vit.py:179-185).  Per item:
//   K, V [S,64] bf16 -> smem (TMA 4-D gather of the strided canonical rows, SWIZZLE_128B; the cls row is
//                     appended by the producer warp)
//   per 128-query tile:  Q tile -> smem ring (TMA); keys in (up to) two blocks A = [0,160), B = [160,S16):
//     S_x = Q K_x^T   : tcgen05.mma SS, fp32 accumulator in TMEM columns [0,160)
//     softmax         : 128 threads, one query row each, straight out of TMEM (pipelined tcgen05.ld); running
//                       max / sum in registers; P (bf16) written back over the consumed S columns (tcgen05.st);
//                       block B rescales O in TMEM by 2^(m_old-m_new) only when the max moved (online softmax)
//     O (+)= P_x V_x  : tcgen05.mma with A = P from TMEM, B = V as an MN-major smem operand, O in columns [160,224)
//     O / l -> bf16 -> canonical output rows (cls query -> out_cls fp32)
// The two CTAs resident on one SM interleave: while one runs its softmax (MUFU-bound: 16 exp2/clk/SM) the other's
// MMAs, TMA loads and output stores proceed.  256 TMEM columns and ~110 KB of shared memory per CTA.
// Warps: 0 TMA producer (+cls rows), 1 MMA issuer, 2 TMEM allocator, 4-7 softmax/epilogue.
// Limits: S <= 304 (else the mma.sync flash kernel in attention.cu is used).
#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

constexpr int SP_ROWS = 304;                  // K/V rows (S rounded up to 16)
constexpr int SP_KV_BYTES = SP_ROWS * 128;    // 38912 (multiple of 1024)
constexpr int SP_QTILE_BYTES = 128 * 128;
constexpr int SP_QSLOTS = 2;
constexpr int SP_SMEM = 2 * SP_KV_BYTES + SP_QSLOTS * SP_QTILE_BYTES + 256 + 1024;
constexpr int SP_BLOCK_A = 160;               // keys in the first block (TMEM S columns)
constexpr int SP_TMEM_O = 160;                // O accumulator columns [160, 224)
constexpr int SP_TMEM_COLS = 256;

struct SpatialArgs {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  __nv_bfloat16* out;
  int64_t ld_out;
  float* out_cls;
  int B, N, T, heads, use_cls;
  int64_t cls_row0;
  float scale_log2;
  float* lse;       // training: [B*T*heads][304] base-2 log-sum-exp of the scaled scores per query token (or nullptr)
  long long* prof;  // SP_PROFILE builds only
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// max over the valid entries of one 32-column chunk (keys key0 .. key0+31, valid if < S)
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], int key0, int S, float mx) {
  if (key0 + 32 <= S) {
    float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      m0 = fmaxf(m0, __uint_as_float(v[e]));
      m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
      m2 = fmaxf(m2, __uint_as_float(v[e + 2]));
      m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
#pragma unroll
  for (int e = 0; e < 32; ++e)
    if (key0 + e < S) mx = fmaxf(mx, __uint_as_float(v[e]));
  return mx;
}

// p = 2^(s*sc - mxs) for one chunk; packs bf16 pairs into pk; returns the chunk's sum.  Full chunks use packed
// fp32x2 FMAs/adds (half the issue slots); the exp2 itself is the MUFU unit (16/clk/SM), the kernel's real bound.
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[32], uint32_t (&pk)[16], int key0, int S, float sc,
                                           float mxs) {
  if (key0 + 32 <= S) {
    const uint64_t sc2 = f2_pack(sc, sc), nm2 = f2_pack(-mxs, -mxs);
    uint64_t la = f2_pack(0.f, 0.f), lb = la;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      float x0, x1, x2, x3;
      f2_unpack(f2_fma(f2_pack_u(v[e], v[e + 1]), sc2, nm2), x0, x1);
      f2_unpack(f2_fma(f2_pack_u(v[e + 2], v[e + 3]), sc2, nm2), x2, x3);
      const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
      la = f2_add(la, f2_pack(p0, p1));
      lb = f2_add(lb, f2_pack(p2, p3));
      pk[e >> 1] = pack_bf16(p0, p1);
      pk[(e >> 1) + 1] = pack_bf16(p2, p3);
    }
    float s0, s1;
    f2_unpack(f2_add(la, lb), s0, s1);
    return s0 + s1;
  }
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), sc, -mxs));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc, -mxs));
    if (key0 + e >= S) p0 = 0.f;
    if (key0 + e + 1 >= S) p1 = 0.f;
    l0 += p0;
    l1 += p1;
    pk[e >> 1] = pack_bf16(p0, p1);
  }
  return l0 + l1;
}

__global__ void __launch_bounds__(256, 2)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tmQfull, const __grid_constant__ CUtensorMap tmQtail,
                       const __grid_constant__ CUtensorMap tmKVfull, const __grid_constant__ CUtensorMap tmKVtail,
                       const SpatialArgs a) {
  extern __shared__ uint8_t smem_sp[];
  const uint32_t raw = smem_u32(smem_sp);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t k_buf = base, v_buf = base + SP_KV_BYTES;
  auto q_buf = [&](int slot) { return base + 2 * SP_KV_BYTES + slot * SP_QTILE_BYTES; };
  const uint32_t bars = base + 2 * SP_KV_BYTES + SP_QSLOTS * SP_QTILE_BYTES;
  const uint32_t kv_full = bars, kv_empty = bars + 8;
  auto q_full = [&](int s) { return bars + 16u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 32u + 8u * s; };
  const uint32_t s_full = bars + 48, p_full = bars + 56, o_full = bars + 64;
  const uint32_t tmem_slot = bars + 72;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_sp + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int n_a = S16 < SP_BLOCK_A ? S16 : SP_BLOCK_A, n_b = S16 - n_a;  // key blocks (multiples of 16)
  const int kv_full_rows = N < 256 ? N : 256, kv_tail_rows = N - kv_full_rows;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQfull);
    prefetch_tmap(&tmQtail);
    prefetch_tmap(&tmKVfull);
    prefetch_tmap(&tmKVtail);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
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
    tmem_alloc(tmem_slot, SP_TMEM_COLS);
    tmem_relinquish();
  }
  // Rows [S, 304) of K/V are never written by TMA: zero them once (P is 0 there, V must be finite).
  for (int idx = threadIdx.x; idx < (SP_ROWS - S) * 8 * 2; idx += blockDim.x) {
    const int which = idx / ((SP_ROWS - S) * 8), rem = idx % ((SP_ROWS - S) * 8);
    const int row = S + (rem >> 3), chunk = rem & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"((which ? v_buf : k_buf) + row * 128 + ((chunk ^ (row & 7)) << 4)),
                 "r"(0)
                 : "memory");
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
      mbar_wait(kv_empty, (kv_it & 1) ^ 1);
      if (a.use_cls && lane < 16) {  // cls k / v rows -> row N of the K / V tiles
        const int which = lane >> 3, chunk = lane & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + (1 + which) * D + h * 64 + chunk * 8);
        const uint32_t dst = (which ? v_buf : k_buf) + N * 128 + ((chunk ^ (N & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2u * N * 128u);
        tma_load_4d(k_buf, &tmKVfull, D + h * 64, t, 0, b, kv_full);
        tma_load_4d(v_buf, &tmKVfull, 2 * D + h * 64, t, 0, b, kv_full);
        if (kv_tail_rows > 0) {
          tma_load_4d(k_buf + 256 * 128, &tmKVtail, D + h * 64, t, 256, b, kv_full);
          tma_load_4d(v_buf + 256 * 128, &tmKVtail, 2 * D + h * 64, t, 256, b, kv_full);
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
    const uint32_t idesc_a = umma_idesc_bf16(128, n_a);
    const uint32_t idesc_b = umma_idesc_bf16(128, n_b > 0 ? n_b : 16);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    const uint64_t kd_a = umma_desc_k_sw128(k_buf), kd_b = umma_desc_k_sw128(k_buf + SP_BLOCK_A * 128);
    const uint64_t vd_a = umma_desc_mn_sw128(v_buf, 1024), vd_b = umma_desc_mn_sw128(v_buf + SP_BLOCK_A * 128, 1024);
    uint32_t kv_it = 0, q_it = 0, p_ct = 0;
#ifdef SP_PROFILE
    long long macc[3] = {0, 0, 0};
    long long mt;
#define SP_T0 mt = clock64();
#define SP_T1(i) macc[i] += clock64() - mt;
#else
#define SP_T0
#define SP_T1(i)
#endif
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++kv_it) {
      SP_T0 mbar_wait(kv_full, kv_it & 1); SP_T1(0)
      for (int j = 0; j < nq; ++j, ++q_it) {
        const int slot = q_it % SP_QSLOTS;
        SP_T0 mbar_wait(q_full(slot), (q_it / SP_QSLOTS) & 1); SP_T1(1)
        tc_fence_after();
        const uint64_t qd = umma_desc_k_sw128(q_buf(slot));
        if (elect_one()) {  // S_a = Q K_a^T
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd_a + 2u * k, idesc_a, k > 0 ? 1u : 0u);
          if (n_b == 0) umma_commit(q_empty(slot));
          umma_commit(s_full);
        }
        __syncwarp();
        SP_T0 mbar_wait(p_full, p_ct & 1); SP_T1(2)
        ++p_ct;
        tc_fence_after();
        if (elect_one()) {  // O = P_a V_a ; then S_b = Q K_b^T (in order behind it: P_a is consumed first)
          for (int kk = 0; kk < (n_a >> 4); ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
            umma_bf16_ts(tmem_base + SP_TMEM_O, tmem_base + 8u * kk, vd_a + 128u * kk, idesc_o, kk > 0 ? 1u : 0u);
          if (n_b > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd_b + 2u * k, idesc_b, k > 0 ? 1u : 0u);
            umma_commit(q_empty(slot));
            umma_commit(s_full);
          } else {
            umma_commit(o_full);
            if (j == nq - 1) umma_commit(kv_empty);
          }
        }
        __syncwarp();
        if (n_b > 0) {
          SP_T0 mbar_wait(p_full, p_ct & 1); SP_T1(2)
          ++p_ct;
          tc_fence_after();
          if (elect_one()) {  // O += P_b V_b
            for (int kk = 0; kk < (n_b >> 4); ++kk)
              umma_bf16_ts(tmem_base + SP_TMEM_O, tmem_base + 8u * kk, vd_b + 128u * kk, idesc_o, 1u);
            umma_commit(o_full);
            if (j == nq - 1) umma_commit(kv_empty);
          }
          __syncwarp();
        }
      }
    }
#ifdef SP_PROFILE
    if (lane == 0 && a.prof && blockIdx.x < 512) {
      long long* d = a.prof + blockIdx.x * 16;
      d[8] = macc[0]; d[9] = macc[1]; d[10] = macc[2];
    }
#endif
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output (one query row per thread)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = a.scale_log2;
    uint32_t s_ct = 0, o_ct = 0;
#ifdef SP_PROFILE
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
#endif
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      for (int j = 0; j < nq; ++j, ++o_ct) {
        const int tok = 128 * j + row;
        const bool valid = tok < S;
        float m_run = -INFINITY, l_run = 0.f;
        for (int blk = 0; blk < (n_b > 0 ? 2 : 1); ++blk, ++s_ct) {
          const int key_base = blk ? SP_BLOCK_A : 0;
          const int nchunk = ((blk ? n_b : n_a) + 31) >> 5;
#ifdef SP_PROFILE
          long long t0 = clock64();
#endif
          mbar_wait(s_full, s_ct & 1);
          tc_fence_after();
#ifdef SP_PROFILE
          long long t1 = clock64();
          pacc[0] += t1 - t0;
#endif
          uint32_t va[32], vb[32], pk[16];
          // ---- pass 1: block maximum (TMEM loads software-pipelined one chunk ahead)
          float mx = m_run;
          tmem_ld_32x32(t_lane, va);
          tmem_ld_wait();
          for (int c = 0; c < nchunk; c += 2) {
            if (c + 1 < nchunk) tmem_ld_32x32(t_lane + 32 * (c + 1), vb);
            mx = chunk_max(va, key_base + 32 * c, S, mx);
            tmem_ld_wait();
            if (c + 1 < nchunk) {
              if (c + 2 < nchunk) tmem_ld_32x32(t_lane + 32 * (c + 2), va);
              mx = chunk_max(vb, key_base + 32 * (c + 1), S, mx);
              tmem_ld_wait();
            }
          }
#ifdef SP_PROFILE
          long long tp1 = clock64();
          pacc[4] += tp1 - t1;
#endif
          // ---- online softmax: when block B raises a row's maximum, O (block A's partial result, in TMEM) is
          // rescaled by 2^(m_old - m_new).  The exact maximum is kept as the reference (not a lazy threshold): the
          // dominant probability is then exactly 1.0 in bf16, which measurably tightens the result.
          if (blk == 1) {
            const bool moved = valid && (mx > m_run);
            const float alpha = moved ? ex2_approx((m_run - mx) * sc) : 1.f;
            l_run *= alpha;
            if (__any_sync(0xffffffffu, moved)) {
              const uint64_t al2 = f2_pack(alpha, alpha);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                tmem_ld_32x32(t_lane + SP_TMEM_O + 32 * hh, va);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  float r0, r1;
                  f2_unpack(f2_mul(f2_pack_u(va[e], va[e + 1]), al2), r0, r1);
                  pk[e] = __float_as_uint(r0);
                  pk[e + 1] = __float_as_uint(r1);
                }
                tmem_st_32x16(t_lane + SP_TMEM_O + 32 * hh, pk);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  float r0, r1;
                  f2_unpack(f2_mul(f2_pack_u(va[16 + e], va[17 + e]), al2), r0, r1);
                  pk[e] = __float_as_uint(r0);
                  pk[e + 1] = __float_as_uint(r1);
                }
                tmem_st_32x16(t_lane + SP_TMEM_O + 32 * hh + 16, pk);
              }
            }
          }
#ifdef SP_PROFILE
          long long tp2 = clock64();
          pacc[5] += tp2 - tp1;
#endif
          const float mxs = mx * sc;  // rows past S compute on stale data; their results are never stored
          m_run = mx;
          // ---- pass 2: p = 2^(s*sc - mx*sc), row sum, P (bf16) over the S columns already consumed
          tmem_ld_32x32(t_lane, va);
          tmem_ld_wait();
          for (int c = 0; c < nchunk; c += 2) {
            if (c + 1 < nchunk) tmem_ld_32x32(t_lane + 32 * (c + 1), vb);
            l_run += chunk_exp(va, pk, key_base + 32 * c, S, sc, mxs);
            tmem_ld_wait();  // chunk c+1 is in registers before P chunk c overwrites columns [16c, 16c+16)
            tmem_st_32x16(t_lane + 16 * c, pk);
            if (c + 1 < nchunk) {
              if (c + 2 < nchunk) tmem_ld_32x32(t_lane + 32 * (c + 2), va);
              l_run += chunk_exp(vb, pk, key_base + 32 * (c + 1), S, sc, mxs);
              tmem_ld_wait();
              tmem_st_32x16(t_lane + 16 * (c + 1), pk);
            }
          }
#ifdef SP_PROFILE
          pacc[6] += clock64() - tp2;
#endif
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full);
#ifdef SP_PROFILE
          pacc[1] += clock64() - t1;
#endif
        }
        // ---- O / l -> global
#ifdef SP_PROFILE
        long long t2 = clock64();
#endif
        mbar_wait(o_full, o_ct & 1);
        tc_fence_after();
#ifdef SP_PROFILE
        long long t3 = clock64();
        pacc[2] += t3 - t2;
#endif
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + SP_TMEM_O, o0);
        tmem_ld_32x32(t_lane + SP_TMEM_O + 32, o1);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l_run;
          if (a.lse) a.lse[static_cast<int64_t>(item) * SP_ROWS + tok] = fmaf(m_run, sc, log2f(l_run));
          if (a.use_cls && tok == N) {
            float4* dst = reinterpret_cast<float4*>(a.out_cls + (static_cast<int64_t>(b) * T + t) * D + h * 64);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              dst[e] = make_float4(__uint_as_float(o0[4 * e]) * inv, __uint_as_float(o0[4 * e + 1]) * inv,
                                   __uint_as_float(o0[4 * e + 2]) * inv, __uint_as_float(o0[4 * e + 3]) * inv);
              dst[8 + e] = make_float4(__uint_as_float(o1[4 * e]) * inv, __uint_as_float(o1[4 * e + 1]) * inv,
                                       __uint_as_float(o1[4 * e + 2]) * inv, __uint_as_float(o1[4 * e + 3]) * inv);
            }
            if (t == 0) {  // frame-0 cls output doubles as the cls input row of the projection (vit.py:198)
              uint4* dc = reinterpret_cast<uint4*>(a.out + (a.cls_row0 + b) * a.ld_out + h * 64);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                dc[e] = make_uint4(pack_bf16(__uint_as_float(o0[8 * e]) * inv, __uint_as_float(o0[8 * e + 1]) * inv),
                                   pack_bf16(__uint_as_float(o0[8 * e + 2]) * inv, __uint_as_float(o0[8 * e + 3]) * inv),
                                   pack_bf16(__uint_as_float(o0[8 * e + 4]) * inv, __uint_as_float(o0[8 * e + 5]) * inv),
                                   pack_bf16(__uint_as_float(o0[8 * e + 6]) * inv, __uint_as_float(o0[8 * e + 7]) * inv));
                dc[4 + e] = make_uint4(pack_bf16(__uint_as_float(o1[8 * e]) * inv, __uint_as_float(o1[8 * e + 1]) * inv),
                                       pack_bf16(__uint_as_float(o1[8 * e + 2]) * inv, __uint_as_float(o1[8 * e + 3]) * inv),
                                       pack_bf16(__uint_as_float(o1[8 * e + 4]) * inv, __uint_as_float(o1[8 * e + 5]) * inv),
                                       pack_bf16(__uint_as_float(o1[8 * e + 6]) * inv, __uint_as_float(o1[8 * e + 7]) * inv));
              }
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
#ifdef SP_PROFILE
        pacc[3] += clock64() - t3;
#endif
      }
    }
#ifdef SP_PROFILE
    if (warp == 4 && lane == 0 && a.prof && blockIdx.x < 512) {
      long long* d = a.prof + blockIdx.x * 16;
      d[0] = pacc[0]; d[1] = pacc[1]; d[2] = pacc[2]; d[3] = pacc[3]; d[4] = clock64() - tstart;
      d[5] = pacc[4]; d[6] = pacc[5]; d[7] = pacc[6];
    }
#endif
  }

#ifdef SP_PROFILE
  if (warp == 4 && lane == 0 && a.prof && blockIdx.x < 512) {
    // filled below via sp_prof_* shared variables
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, SP_TMEM_COLS);
}

// ============================================================================================ streamed K/V
// Same attention for ANY number of tokens per frame (480x640 frames: S = 1201): K and V do not fit shared memory, so a
// work item is one 128-query tile of one (clip, frame, head) and the keys are streamed in blocks of 128 through a
// two-stage TMA ring (flash attention): per block  S = Q K_j^T (SS MMA, TMEM columns [0,128))  ->  online softmax by 128
// threads straight out of TMEM (running max / sum in registers, O in TMEM rescaled only when a row's max moved, P written
// back over the consumed S columns)  ->  O += P_j V_j (TS MMA, V MN-major) with S_{j+1} = Q K_{j+1}^T queued right behind
// it.  Consecutive items share (b,t,head), so the K/V blocks of one head are re-read from L2, not HBM.
// 192 TMEM columns, ~97 KB shared memory: two CTAs per SM, as in the resident-K/V kernel above.
constexpr int ST_KB = 128;                       // keys per block
constexpr int ST_STAGES = 2;
constexpr int ST_STAGE_BYTES = 2 * ST_KB * 128;  // K block + V block
constexpr int ST_QTILE_BYTES = 128 * 128;
constexpr int ST_QSLOTS = 2;
constexpr int ST_SMEM = ST_STAGES * ST_STAGE_BYTES + ST_QSLOTS * ST_QTILE_BYTES + 256 + 1024;
constexpr int ST_TMEM_O = 128;
constexpr int ST_TMEM_COLS = 256;

__global__ void __launch_bounds__(256, 2)
attn_spatial_stream_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmTail,
                           const SpatialArgs a) {
  extern __shared__ uint8_t smem_st[];
  const uint32_t raw = smem_u32(smem_st);
  const uint32_t base = (raw + 1023u) & ~1023u;
  auto k_buf = [&](int st) { return base + st * ST_STAGE_BYTES; };
  auto v_buf = [&](int st) { return base + st * ST_STAGE_BYTES + ST_KB * 128; };
  auto q_buf = [&](int slot) { return base + ST_STAGES * ST_STAGE_BYTES + slot * ST_QTILE_BYTES; };
  const uint32_t bars = base + ST_STAGES * ST_STAGE_BYTES + ST_QSLOTS * ST_QTILE_BYTES;
  auto kv_full = [&](int st) { return bars + 8u * st; };
  auto kv_empty = [&](int st) { return bars + 16u + 8u * st; };
  auto q_full = [&](int s) { return bars + 32u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 48u + 8u * s; };
  const uint32_t s_full = bars + 64, p_full = bars + 72, o_full = bars + 80;
  const uint32_t tmem_slot = bars + 88;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_st + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int nq = (S + 127) >> 7;
  const int nblk = (S + ST_KB - 1) / ST_KB;
  const int D = heads * 64;
  const int items = a.B * T * heads * nq;
  const int tail_rows = N % 128;  // patch rows in the last (partial) 128-row tile; 0: N is a multiple of 128

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmFull);
    prefetch_tmap(&tmTail);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < ST_QSLOTS; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, ST_TMEM_COLS);
    tmem_relinquish();
  }
  // Rows that TMA never writes (past the last key / query) keep whatever the stage held before; they are masked, but V
  // must be finite (0 * NaN) — zero everything once, afterwards the buffers only ever hold real data.
  for (int idx = threadIdx.x; idx < (ST_STAGES * ST_STAGE_BYTES + ST_QSLOTS * ST_QTILE_BYTES) / 16; idx += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + idx * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t kv_it = 0, q_it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++q_it) {
      const int qt = item % nq, h = (item / nq) % heads, t = (item / (nq * heads)) % T, b = item / (nq * heads * T);
      const int slot = q_it % ST_QSLOTS;
      mbar_wait(q_empty(slot), ((q_it / ST_QSLOTS) & 1) ^ 1);
      {
        const int rows = (N - 128 * qt) < 128 ? (N - 128 * qt) : 128;  // patch rows in this query tile (may be <= 0)
        if (a.use_cls && (N >> 7) == qt && lane < 8) {                  // the cls query is token N
          const int r = N - 128 * qt;
          const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + h * 64 + lane * 8);
          const uint32_t dst = q_buf(slot) + r * 128 + ((lane ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (rows > 0) {
            mbar_expect_tx(q_full(slot), static_cast<uint32_t>(rows) * 128u);
            tma_load_4d(q_buf(slot), rows == 128 ? &tmFull : &tmTail, h * 64, t, 128 * qt, b, q_full(slot));
          } else {
            mbar_arrive(q_full(slot));
          }
        }
        __syncwarp();
      }
      for (int j = 0; j < nblk; ++j, ++kv_it) {
        const int st = kv_it % ST_STAGES;
        mbar_wait(kv_empty(st), ((kv_it / ST_STAGES) & 1) ^ 1);
        const int rows = (N - ST_KB * j) < ST_KB ? (N - ST_KB * j) : ST_KB;
        if (a.use_cls && (N >> 7) == j && lane < 16) {  // cls k / v rows -> row N - 128 j of this block
          const int which = lane >> 3, chunk = lane & 7, r = N - ST_KB * j;
          const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + (1 + which) * D + h * 64 + chunk * 8);
          const uint32_t dst = (which ? v_buf(st) : k_buf(st)) + r * 128 + ((chunk ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (rows > 0) {
            mbar_expect_tx(kv_full(st), 2u * static_cast<uint32_t>(rows) * 128u);
            const CUtensorMap* m = rows == ST_KB ? &tmFull : &tmTail;
            tma_load_4d(k_buf(st), m, D + h * 64, t, ST_KB * j, b, kv_full(st));
            tma_load_4d(v_buf(st), m, 2 * D + h * 64, t, ST_KB * j, b, kv_full(st));
          } else {
            mbar_arrive(kv_full(st));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, ST_KB);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    uint32_t kv_it = 0, q_it = 0, p_ct = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++q_it) {
      const int slot = q_it % ST_QSLOTS;
      mbar_wait(q_full(slot), (q_it / ST_QSLOTS) & 1);
      const uint64_t qd = umma_desc_k_sw128(q_buf(slot));
      // S_0 = Q K_0^T
      mbar_wait(kv_full(kv_it % ST_STAGES), (kv_it / ST_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = umma_desc_k_sw128(k_buf(kv_it % ST_STAGES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd + 2u * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j, ++kv_it) {
        const int st = kv_it % ST_STAGES;
        mbar_wait(p_full, p_ct & 1);
        ++p_ct;
        if (j + 1 < nblk) mbar_wait(kv_full((kv_it + 1) % ST_STAGES), ((kv_it + 1) / ST_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          // O (+)= P_j V_j : 16 keys per MMA (8 TMEM columns of P, 16 rows = 2048 B of V); then S_{j+1} right behind it
          const uint64_t vd = umma_desc_mn_sw128(v_buf(st), 1024);
#pragma unroll
          for (int kk = 0; kk < ST_KB / 16; ++kk)
            umma_bf16_ts(tmem_base + ST_TMEM_O, tmem_base + 8u * kk, vd + 128u * kk, idesc_o, (j | kk) != 0 ? 1u : 0u);
          if (j + 1 < nblk) {
            const uint64_t kd = umma_desc_k_sw128(k_buf((kv_it + 1) % ST_STAGES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, qd + 2u * k, kd + 2u * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(kv_empty(st));
            umma_commit(s_full);
          } else {
            umma_commit(kv_empty(st));
            umma_commit(q_empty(slot));
            umma_commit(o_full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ online softmax + output (one query row per thread)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = a.scale_log2;
    uint32_t s_ct = 0, o_ct = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++o_ct) {
      const int qt = item % nq, h = (item / nq) % heads, t = (item / (nq * heads)) % T, b = item / (nq * heads * T);
      const int tok = 128 * qt + row;
      const bool valid = tok < S;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j, ++s_ct) {
        const int key_base = ST_KB * j;
        mbar_wait(s_full, s_ct & 1);
        tc_fence_after();
        uint32_t va[32], vb[32], pk[16];
        // ---- pass 1: block maximum
        float mx = m_run;
        tmem_ld_32x32(t_lane, va);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_32x32(t_lane + 32 * (c + 1), vb);
          mx = chunk_max(va, key_base + 32 * c, S, mx);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld_32x32(t_lane + 32 * (c + 2), va);
          mx = chunk_max(vb, key_base + 32 * (c + 1), S, mx);
          tmem_ld_wait();
        }
        // ---- the running maximum moved: rescale the partial O (exact maximum kept as the reference)
        if (j > 0) {
          const bool moved = valid && (mx > m_run);
          const float alpha = moved ? ex2_approx((m_run - mx) * sc) : 1.f;
          l_run *= alpha;
          if (__any_sync(0xffffffffu, moved)) {
            const uint64_t al2 = f2_pack(alpha, alpha);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              tmem_ld_32x32(t_lane + ST_TMEM_O + 32 * hh, va);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                float r0, r1;
                f2_unpack(f2_mul(f2_pack_u(va[e], va[e + 1]), al2), r0, r1);
                pk[e] = __float_as_uint(r0);
                pk[e + 1] = __float_as_uint(r1);
              }
              tmem_st_32x16(t_lane + ST_TMEM_O + 32 * hh, pk);
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                float r0, r1;
                f2_unpack(f2_mul(f2_pack_u(va[16 + e], va[17 + e]), al2), r0, r1);
                pk[e] = __float_as_uint(r0);
                pk[e + 1] = __float_as_uint(r1);
              }
              tmem_st_32x16(t_lane + ST_TMEM_O + 32 * hh + 16, pk);
            }
          }
        }
        const float mxs = mx * sc;
        m_run = mx;
        // ---- pass 2: p = 2^(s*sc - mx*sc), row sum, P (bf16) over the S columns already consumed
        tmem_ld_32x32(t_lane, va);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_32x32(t_lane + 32 * (c + 1), vb);
          l_run += chunk_exp(va, pk, key_base + 32 * c, S, sc, mxs);
          tmem_ld_wait();  // chunk c+1 is in registers before P chunk c overwrites columns [16c, 16c+16)
          tmem_st_32x16(t_lane + 16 * c, pk);
          if (c + 2 < 4) tmem_ld_32x32(t_lane + 32 * (c + 2), va);
          l_run += chunk_exp(vb, pk, key_base + 32 * (c + 1), S, sc, mxs);
          tmem_ld_wait();
          tmem_st_32x16(t_lane + 16 * (c + 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // ---- O / l -> global
      mbar_wait(o_full, o_ct & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_lane + ST_TMEM_O, o0);
      tmem_ld_32x32(t_lane + ST_TMEM_O + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      if (valid) {
        const float inv = 1.0f / l_run;
        if (a.lse) a.lse[(static_cast<int64_t>(item / nq)) * (nq * 128) + tok] = fmaf(m_run, sc, log2f(l_run));
        uint32_t ob[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          ob[e] = pack_bf16(__uint_as_float(o0[2 * e]) * inv, __uint_as_float(o0[2 * e + 1]) * inv);
          ob[16 + e] = pack_bf16(__uint_as_float(o1[2 * e]) * inv, __uint_as_float(o1[2 * e + 1]) * inv);
        }
        if (a.use_cls && tok == N) {
          float4* dst = reinterpret_cast<float4*>(a.out_cls + (static_cast<int64_t>(b) * T + t) * D + h * 64);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            dst[e] = make_float4(__uint_as_float(o0[4 * e]) * inv, __uint_as_float(o0[4 * e + 1]) * inv,
                                 __uint_as_float(o0[4 * e + 2]) * inv, __uint_as_float(o0[4 * e + 3]) * inv);
            dst[8 + e] = make_float4(__uint_as_float(o1[4 * e]) * inv, __uint_as_float(o1[4 * e + 1]) * inv,
                                     __uint_as_float(o1[4 * e + 2]) * inv, __uint_as_float(o1[4 * e + 3]) * inv);
          }
          if (t == 0) {  // frame-0 cls output doubles as the cls input row of the projection (vit.py:198)
            uint4* dc = reinterpret_cast<uint4*>(a.out + (a.cls_row0 + b) * a.ld_out + h * 64);
#pragma unroll
            for (int e = 0; e < 8; ++e) dc[e] = make_uint4(ob[4 * e], ob[4 * e + 1], ob[4 * e + 2], ob[4 * e + 3]);
          }
        } else {
          uint4* dst = reinterpret_cast<uint4*>(a.out + ((static_cast<int64_t>(b) * N + tok) * T + t) * a.ld_out + h * 64);
#pragma unroll
          for (int e = 0; e < 8; ++e) dst[e] = make_uint4(ob[4 * e], ob[4 * e + 1], ob[4 * e + 2], ob[4 * e + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, ST_TMEM_COLS);
}

// 4-D view of the patch rows of qkv: (column, t, n, b) -> ((b*N+n)*T+t)*ld + column; box = 64 columns x box_n tokens.
int make_patch_tmap(void* m, const void* qkv, int64_t ld, int cols, int B, int N, int T, int box_n) {
  const uint64_t dims[4] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(T), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(T) * ld * 2,
                               static_cast<uint64_t>(N) * T * ld * 2};
  const uint32_t box[4] = {64, 1, static_cast<uint32_t>(box_n), 1};
  return make_tmap_nd(m, false, qkv, 4, dims, strides, box);
}

int launch_spatial_tc(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                      int heads, int use_cls, int64_t cls_row0, cudaStream_t stream, float* lse) {
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
                B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, lse, nullptr};
#ifdef SP_PROFILE
  static long long* dprof = nullptr;
  if (!dprof) cudaMalloc(&dprof, 16 * 8 * 512);
  cudaMemsetAsync(dprof, 0, 16 * 8 * 512, stream);
  a.prof = dprof;
#endif
  const int items = B * T * heads;
  const int slots = 2 * sm_count();
  int grid = items < slots ? items : slots;
#ifdef SP_PROFILE
  if (const char* e = getenv("TCOW_SP_GRID")) grid = atoi(e);
#endif
  attn_spatial_tc_kernel<<<grid, 256, SP_SMEM, stream>>>(tmQf, tmQt, tmKVf, tmKVt, a);
#ifdef SP_PROFILE
  {
    static int calls = 0;
    if (++calls == 5) {
      cudaStreamSynchronize(stream);
      static long long h[16 * 512];
      cudaMemcpy(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost);
      double acc[16] = {0};
      for (int c = 0; c < grid && c < 512; ++c) for (int k = 0; k < 16; ++k) acc[k] += (double)h[c * 16 + k];
      const int n = grid < 512 ? grid : 512;
      printf("SP_PROFILE per CTA avg cycles: total %.0f | softmax warp: wait_s %.0f compute %.0f wait_o %.0f epilogue %.0f | mma warp: wait_kv %.0f wait_q %.0f wait_p %.0f | pass1 %.0f rescale %.0f pass2 %.0f\n",
             acc[4] / n, acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[8] / n, acc[9] / n, acc[10] / n, acc[5] / n, acc[6] / n, acc[7] / n);
    }
  }
#endif
  return check_launch("attn_spatial_tc_kernel");
}

int launch_spatial_stream(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                          int heads, int use_cls, int64_t cls_row0, cudaStream_t stream) {
  alignas(64) CUtensorMap tmFull, tmTail;
  const int cols = 3 * heads * 64;
  const int tail = N % 128;
  int rc;
  if ((rc = make_patch_tmap(&tmFull, qkv, ld_qkv, cols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&tmTail, qkv, ld_qkv, cols, B, N, T, tail > 0 ? tail : 1))) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  SpatialArgs a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out), ld_out, out_cls,
                B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, nullptr, nullptr};
  const int S = N + use_cls;
  const long long items = static_cast<long long>(B) * T * heads * ((S + 127) / 128);
  const int slots = 2 * sm_count();
  const int grid = items < slots ? static_cast<int>(items) : slots;
  attn_spatial_stream_kernel<<<grid, 256, ST_SMEM, stream>>>(tmFull, tmTail, a);
  return check_launch("attn_spatial_stream_kernel");
}

}  // namespace tcow
