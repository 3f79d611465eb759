// Spatial attention on the 5th-gen tensor cores (tcgen05 + TMEM) for frames whose keys do NOT fit shared memory:
// K and V streamed in 128-key blocks (flash attention), two CTAs per SM.  Frames with S = N (+1 cls) <= 304 tokens run the
// resident-K/V kernel of attn_spatial_r1.cu; this file serves S > 304 (480x640 frames: S = 1201) and holds the
// 4-D tensor-map helper both share.  Reference: vit.py:78-111 as called at vit.py:186 on the tokens of vit.py:179-185.
#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {


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

// Attention for ANY number of tokens per frame (480x640 frames: S = 1201): K and V do not fit shared memory, so a
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
                B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, nullptr};
  const int S = N + use_cls;
  const long long items = static_cast<long long>(B) * T * heads * ((S + 127) / 128);
  const int slots = 2 * sm_count();
  const int grid = items < slots ? static_cast<int>(items) : slots;
  attn_spatial_stream_kernel<<<grid, 256, ST_SMEM, stream>>>(tmFull, tmTail, a);
  return check_launch("attn_spatial_stream_kernel");
}

}  // namespace tcow
