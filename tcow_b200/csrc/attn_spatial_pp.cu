// Spatial attention on tcgen05 + TMEM, resident K/V, ONE persistent CTA per SM with TWO query tiles in flight ("ping-pong").
//
// Work item = (clip b, frame t, head h): full softmax attention over S = N (+1 cls) <= 304 tokens of head dim 64
// (vit.py:78-111 as called at vit.py:186 on the tokens assembled at vit.py:179-185).  The stream of 128-query tiles of
// this CTA's items alternates between two softmax groups (8 warps each); while one group runs its softmax the tensor core
// works on the other group's tile, so neither the MMA -> softmax -> MMA hand-offs nor the K/V and Q loads are exposed:
//   warp 0   TMA producer: K/V of item i into buffer i&1 (4-D gather of the strided canonical rows, SWIZZLE_128B; the cls
//            row is appended by hand), Q tiles into a 4-slot ring
//   warp 1   MMA issuer for both groups, fixed round-robin:  S = Q K^T (SS, keys in two blocks 160 + 144, fp32 in the
//            group's TMEM columns [0,160)),  O (+)= P V (A = P from TMEM, V MN-major, O in columns [160,224))
//   warp 2   TMEM allocator (all 512 columns: two regions of 224)
//   warps 4-11 / 12-19  softmax group 0 / 1: every query row is shared by TWO threads (warps q and q+4 of the group read
//            the same 32 TMEM lanes), each owning half of the key columns of a block: block max through shared memory,
//            exact online softmax (O rescaled in TMEM only when a row's max moved), P (bf16) written over the owner's own
//            consumed score columns; O / l -> bf16 -> the tile's (dead) Q slot as a swizzled staging tile -> ONE TMA store
//            of the strided canonical rows (a warp-wide global store would touch 32 different 128-byte lines).
// Lane quarters wholly past S (the last tile holds 45 of 128 rows at S = 301) skip the softmax and only keep the barriers
// going.  Replaces the two-CTAs-per-SM kernel of round 1 (229.7 -> see profiles/r02_notes.md).
#include <math.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {
constexpr int PP_ROWS = 304;                  // K/V rows (S rounded up to 16)
constexpr int PP_KV_BYTES = PP_ROWS * 128;    // one of K or V: 38912 (multiple of 1024)
constexpr int PP_QTILE_BYTES = 128 * 128;
constexpr int PP_QSLOTS = 4;
constexpr int PP_THREADS = 640;
constexpr int PP_GROUP_THREADS = 256;
constexpr int PP_BLOCK_A = 160;               // keys in the first block (TMEM score columns)
constexpr int PP_TMEM_O = 160;                // O accumulator columns [160, 224) of a region
constexpr int PP_REGION = 224;                // TMEM columns per softmax group
constexpr int PP_TMEM_COLS = 512;
constexpr int PP_XCH_FLOATS = 2 * 256 + 256;  // per group: max partials [2 slots][2 halves][128], row-sum partials [2][128]
constexpr int PP_BAR_BYTES = 256;
constexpr int PP_SMEM = 2 * 2 * PP_KV_BYTES + PP_QSLOTS * PP_QTILE_BYTES + PP_BAR_BYTES + 2 * PP_XCH_FLOATS * 4 + 1024;

struct PpArgs {
  const __nv_bfloat16* qkv;
  int64_t ld_qkv;
  __nv_bfloat16* out;
  int64_t ld_out;
  float* out_cls;
  int B, N, T, heads, use_cls;
  int64_t cls_row0;
  float scale_log2;
  float* lse;  // training: [B*T*heads][304] base-2 log-sum-exp of the scaled scores per query token (or nullptr)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}
// max over the valid entries of one 16-column chunk (keys key0 .. key0+15, valid if < S)
__device__ __forceinline__ float chunk16_max(const uint32_t (&v)[16], int key0, int S, float mx) {
  if (key0 + 16 <= S) {
    float m0 = mx, m1 = -INFINITY;
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
      m0 = max3(m0, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
      m1 = max3(m1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
    }
    return fmaxf(m0, m1);
  }
#pragma unroll
  for (int e = 0; e < 16; ++e)
    if (key0 + e < S) mx = fmaxf(mx, __uint_as_float(v[e]));
  return mx;
}
// p = 2^(s*sc - mxs) for one chunk; packs bf16 pairs into pk; returns the chunk's sum.  Full chunks use packed fp32x2
// FMAs / adds (half the issue slots); the exp2 itself is the MUFU unit (16/clk/SM).
__device__ __forceinline__ float chunk16_exp(const uint32_t (&v)[16], uint32_t (&pk)[8], int key0, int S, float sc, float mxs) {
  if (key0 + 16 <= S) {
    const uint64_t sc2 = f2_pack(sc, sc), nm2 = f2_pack(-mxs, -mxs);
    uint64_t la = f2_pack(0.f, 0.f), lb = la;
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
      float x0, x1, x2, x3;
      f2_unpack(f2_fma(f2_pack_u(v[e], v[e + 1]), sc2, nm2), x0, x1);
      f2_unpack(f2_fma(f2_pack_u(v[e + 2], v[e + 3]), sc2, nm2), x2, x3);
      const float p0 = ex2f(x0), p1 = ex2f(x1), p2 = ex2f(x2), p3 = ex2f(x3);
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
  for (int e = 0; e < 16; e += 2) {
    float p0 = ex2f(fmaf(__uint_as_float(v[e]), sc, -mxs));
    float p1 = ex2f(fmaf(__uint_as_float(v[e + 1]), sc, -mxs));
    if (key0 + e >= S) p0 = 0.f;
    if (key0 + e + 1 >= S) p1 = 0.f;
    l0 += p0;
    l1 += p1;
    pk[e >> 1] = pack_bf16(p0, p1);
  }
  return l0 + l1;
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// Named barriers (id 0 is __syncthreads): 1 + 4*group + quarter = the two warps sharing a lane quarter; 9 + group = the
// whole softmax group.  Immediate ids so that ptxas does not reserve all 16.
__device__ __forceinline__ void pair_sync(int id) {
  switch (id) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
    case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
    case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void group_sync(int group) {
  if (group == 0) asm volatile("bar.sync 9, 256;" ::: "memory");
  else asm volatile("bar.sync 10, 256;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
}  // namespace

__global__ void __launch_bounds__(PP_THREADS, 1)
attn_spatial_pp_kernel(const __grid_constant__ CUtensorMap tmQfull, const __grid_constant__ CUtensorMap tmQtail,
                       const __grid_constant__ CUtensorMap tmKVfull, const __grid_constant__ CUtensorMap tmKVtail,
                       const __grid_constant__ CUtensorMap tmOut, const PpArgs a) {
  extern __shared__ uint8_t smem_pp[];
  const uint32_t raw = smem_u32(smem_pp);
  const uint32_t base = (raw + 1023u) & ~1023u;
  auto k_buf = [&](int b) { return base + b * 2 * PP_KV_BYTES; };
  auto v_buf = [&](int b) { return base + b * 2 * PP_KV_BYTES + PP_KV_BYTES; };
  auto q_buf = [&](int slot) { return base + 4 * PP_KV_BYTES + slot * PP_QTILE_BYTES; };
  const uint32_t bars = base + 4 * PP_KV_BYTES + PP_QSLOTS * PP_QTILE_BYTES;
  auto kv_full = [&](int b) { return bars + 8u * b; };
  auto kv_empty = [&](int b) { return bars + 16u + 8u * b; };
  auto q_full = [&](int s) { return bars + 32u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 64u + 8u * s; };
  auto s_full = [&](int g) { return bars + 96u + 8u * g; };
  auto p_full = [&](int g) { return bars + 112u + 8u * g; };
  auto o_full = [&](int g) { return bars + 128u + 8u * g; };
  const uint32_t tmem_slot = bars + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_pp + (tmem_slot - raw));
  float* xch = reinterpret_cast<float*>(smem_pp + (bars + PP_BAR_BYTES - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int my_items = items > static_cast<int>(blockIdx.x) ? (items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const uint32_t NT = static_cast<uint32_t>(my_items) * nq;           // this CTA's tile stream
  const int n_a = S16 < PP_BLOCK_A ? S16 : PP_BLOCK_A, n_b = S16 - n_a;  // key blocks (multiples of 16)
  const int nblk = n_b > 0 ? 2 : 1;
  const int kv_full_rows = N < 256 ? N : 256, kv_tail_rows = N - kv_full_rows;
  // 16-key chunks of a block are split between the two column halves: [0, split) and [split, n/16).  P (bf16, 8 TMEM
  // columns per chunk) is written inside the owner's own score columns: chunk c of half 0 at column 8c, of half 1 at
  // 16*split + 8(c - split) — neither half ever overwrites scores the other has not read yet.
  const int ca = n_a >> 4, cb = n_b >> 4;
  const int split_a = (ca + 1) >> 1, split_b = (cb + 1) >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQfull);
    prefetch_tmap(&tmQtail);
    prefetch_tmap(&tmKVfull);
    prefetch_tmap(&tmKVtail);
    prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(kv_full(b), 1);
      mbar_init(kv_empty(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(p_full(b), PP_GROUP_THREADS / 32);   // one arrival per softmax warp
      mbar_init(o_full(b), 1);
    }
    for (int s = 0; s < PP_QSLOTS; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, PP_TMEM_COLS);
    tmem_relinquish();
  }
  // Rows [S, 304) of K/V are never written by TMA: zero them once in both buffers (P is 0 there, V must be finite).
  for (int idx = threadIdx.x; idx < (PP_ROWS - S) * 8 * 4; idx += blockDim.x) {
    const int which = idx / ((PP_ROWS - S) * 8), rem = idx % ((PP_ROWS - S) * 8);
    const int row = S + (rem >> 3), chunk = rem & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + which * PP_KV_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4)),
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
    uint32_t n = 0;
    for (int ii = 0; ii < my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      const int kb = ii & 1;
      mbar_wait(kv_empty(kb), ((ii >> 1) & 1) ^ 1);
      if (a.use_cls && lane < 16) {  // cls k / v rows -> row N of the K / V tiles
        const int which = lane >> 3, chunk = lane & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + (1 + which) * D + h * 64 + chunk * 8);
        const uint32_t dst = (which ? v_buf(kb) : k_buf(kb)) + N * 128 + ((chunk ^ (N & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        mbar_expect_tx(kv_full(kb), 2u * N * 128u);
        tma_load_4d(k_buf(kb), &tmKVfull, D + h * 64, t, 0, b, kv_full(kb));
        tma_load_4d(v_buf(kb), &tmKVfull, 2 * D + h * 64, t, 0, b, kv_full(kb));
        if (kv_tail_rows > 0) {
          tma_load_4d(k_buf(kb) + 256 * 128, &tmKVtail, D + h * 64, t, 256, b, kv_full(kb));
          tma_load_4d(v_buf(kb) + 256 * 128, &tmKVtail, 2 * D + h * 64, t, 256, b, kv_full(kb));
        }
      }
      __syncwarp();
      for (int j = 0; j < nq; ++j, ++n) {
        const int slot = n % PP_QSLOTS;
        mbar_wait(q_empty(slot), ((n / PP_QSLOTS) & 1) ^ 1);   // released by the output store of the tile 4 back
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
    // ------------------------------------------------------------------ MMA issuer (both groups, fixed round-robin)
    const uint32_t idesc_a = umma_idesc_bf16(128, n_a);
    const uint32_t idesc_b = umma_idesc_bf16(128, n_b > 0 ? n_b : 16);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    auto p_col = [](int kk, int split) { return kk < split ? 8 * kk : 16 * split + 8 * (kk - split); };
    uint32_t pc[2] = {0, 0};
    // S_a of tile n into its group's score columns
    auto issue_sa = [&](uint32_t n) {
      const uint32_t ii = n / nq, j = n % nq;
      const int g = n & 1, slot = n % PP_QSLOTS, kb = ii & 1;
      if (j == 0) mbar_wait(kv_full(kb), (ii >> 1) & 1);
      mbar_wait(q_full(slot), (n / PP_QSLOTS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = umma_desc_k_sw128(q_buf(slot)), kd = umma_desc_k_sw128(k_buf(kb));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + g * PP_REGION, qd + 2u * k, kd + 2u * k, idesc_a, k > 0 ? 1u : 0u);
        umma_commit(s_full(g));
      }
      __syncwarp();
    };
    // the MMAs that follow block `blk` of tile n's softmax
    auto step = [&](uint32_t n, int blk) {
      const uint32_t ii = n / nq, j = n % nq;
      const int g = n & 1, slot = n % PP_QSLOTS, kb = ii & 1;
      const uint32_t region = tmem_base + g * PP_REGION;
      mbar_wait(p_full(g), pc[g] & 1);
      ++pc[g];
      tc_fence_after();
      if (elect_one()) {
        if (blk == 0) {  // O = P_a V_a ; then S_b = Q K_b^T (in order behind it: P_a is consumed first)
          const uint64_t vd = umma_desc_mn_sw128(v_buf(kb), 1024);
          for (int kk = 0; kk < ca; ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
            umma_bf16_ts(region + PP_TMEM_O, region + p_col(kk, split_a), vd + 128u * kk, idesc_o, kk > 0 ? 1u : 0u);
          if (n_b > 0) {
            const uint64_t qd = umma_desc_k_sw128(q_buf(slot)), kd = umma_desc_k_sw128(k_buf(kb) + PP_BLOCK_A * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(region, qd + 2u * k, kd + 2u * k, idesc_b, k > 0 ? 1u : 0u);
            umma_commit(s_full(g));
          }
        } else {  // O += P_b V_b
          const uint64_t vd = umma_desc_mn_sw128(v_buf(kb) + PP_BLOCK_A * 128, 1024);
          for (int kk = 0; kk < cb; ++kk)
            umma_bf16_ts(region + PP_TMEM_O, region + p_col(kk, split_b), vd + 128u * kk, idesc_o, 1u);
        }
        if (blk == nblk - 1) {
          umma_commit(o_full(g));
          if (j == static_cast<uint32_t>(nq) - 1) umma_commit(kv_empty(kb));
        }
      }
      __syncwarp();
    };
    if (NT > 0) issue_sa(0);
    if (NT > 1) issue_sa(1);
    for (uint32_t n0 = 0; n0 < NT; n0 += 2) {
      for (int blk = 0; blk < nblk; ++blk) {
        step(n0, blk);
        if (blk == nblk - 1 && n0 + 2 < NT) issue_sa(n0 + 2);   // the scores region is free once its last P is consumed
        if (n0 + 1 < NT) {
          step(n0 + 1, blk);
          if (blk == nblk - 1 && n0 + 3 < NT) issue_sa(n0 + 3);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output: two threads per query row
    const int g = (warp - 4) >> 3;                 // softmax group = parity of the tiles it takes
    const int gw = (warp - 4) & 7;                 // warp within the group
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access (warp id % 4)
    const int half = gw >> 2;                      // column half
    const int row = quarter * 32 + lane;
    const uint32_t region = tmem_base + g * PP_REGION;
    const uint32_t t_lane = region + (static_cast<uint32_t>(quarter * 32) << 16);
    float* x_max = xch + g * PP_XCH_FLOATS;
    float* x_sum = x_max + 2 * 256;
    const int pair_id = g * 4 + quarter;
    const float sc = a.scale_log2;
    uint32_t s_ct = 0, o_ct = 0;
    for (uint32_t n = g; n < NT; n += 2, ++o_ct) {
      const uint32_t ii = n / nq;
      const int j = n % nq, slot = n % PP_QSLOTS;
      const int item = blockIdx.x + ii * gridDim.x;
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      const int tok = 128 * j + row;
      const bool valid = tok < S;
      const bool live = 128 * j + quarter * 32 < S;   // warp-uniform: a lane quarter wholly past S only keeps the barriers going
      float m_run = -INFINITY, l_run = 0.f;
      for (int blk = 0; blk < nblk; ++blk, ++s_ct) {
        const int key_base = blk ? PP_BLOCK_A : 0;
        const int nc = blk ? cb : ca, split = blk ? split_b : split_a;
        const int c0 = half ? split : 0, c1 = half ? nc : split;         // this warp's 16-key chunks of the block
        const uint32_t s_col = t_lane + 16 * c0;                         // first score column of this half
        mbar_wait(s_full(g), s_ct & 1);
        tc_fence_after();
        if (live) {
          uint32_t va[16], vb[16], pk[8];
          // ---- pass 1: maximum over this half's columns, then over both halves through shared memory
          float mx = -INFINITY;
          if (c0 < c1) {
            tmem_ld_32x16(s_col, va);
            tmem_ld_wait();
            for (int c = c0; c < c1; c += 2) {
              if (c + 1 < c1) tmem_ld_32x16(s_col + 16 * (c + 1 - c0), vb);
              mx = chunk16_max(va, key_base + 16 * c, S, mx);
              tmem_ld_wait();
              if (c + 1 < c1) {
                if (c + 2 < c1) tmem_ld_32x16(s_col + 16 * (c + 2 - c0), va);
                mx = chunk16_max(vb, key_base + 16 * (c + 1), S, mx);
                tmem_ld_wait();
              }
            }
          }
          float* xm = x_max + (s_ct & 1) * 256;
          xm[half * 128 + row] = mx;
          pair_sync(pair_id);
          mx = max3(m_run, mx, xm[(half ^ 1) * 128 + row]);
          // ---- online softmax: when block B raises a row's maximum, O (block A's partial result, in TMEM) is rescaled by
          // 2^(m_old - m_new); each half rescales its own 32 of the 64 columns.  The exact maximum is kept (not a lazy
          // threshold): the dominant probability is then exactly 1.0 in bf16, which measurably tightens the result.
          if (blk == 1) {
            const bool moved = valid && (mx > m_run);
            const float alpha = moved ? ex2f((m_run - mx) * sc) : 1.f;
            l_run *= alpha;
            if (__any_sync(0xffffffffu, moved)) {
              const uint64_t al2 = f2_pack(alpha, alpha);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint32_t oc = t_lane + PP_TMEM_O + 32 * half + 16 * hh;
                tmem_ld_32x16(oc, va);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  float r0, r1;
                  f2_unpack(f2_mul(f2_pack_u(va[e], va[e + 1]), al2), r0, r1);
                  vb[e] = __float_as_uint(r0);
                  vb[e + 1] = __float_as_uint(r1);
                }
                tmem_st_32x16(oc, vb);
              }
            }
          }
          const float mxs = mx * sc;  // rows past S compute on stale data; their results are never stored
          m_run = mx;
          // ---- pass 2: p = 2^(s*sc - mx*sc), row sum, P (bf16) over this half's score columns already consumed
          if (c0 < c1) {
            tmem_ld_32x16(s_col, va);
            tmem_ld_wait();
            for (int c = c0; c < c1; c += 2) {
              if (c + 1 < c1) tmem_ld_32x16(s_col + 16 * (c + 1 - c0), vb);
              l_run += chunk16_exp(va, pk, key_base + 16 * c, S, sc, mxs);
              tmem_ld_wait();
              tmem_st_32x8(s_col + 8 * (c - c0), pk);
              if (c + 1 < c1) {
                if (c + 2 < c1) tmem_ld_32x16(s_col + 16 * (c + 2 - c0), va);
                l_run += chunk16_exp(vb, pk, key_base + 16 * (c + 1), S, sc, mxs);
                tmem_ld_wait();
                tmem_st_32x8(s_col + 8 * (c + 1 - c0), pk);
              }
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));   // one arrival per warp: 256 single arrivals serialise on the barrier word
      }
      // ---- O / l -> bf16 -> staging tile (this tile's Q slot: Q is dead once the last Q K^T has completed) -> TMA store
      mbar_wait(o_full(g), o_ct & 1);
      tc_fence_after();
      const uint32_t stage = q_buf(slot);
      if (live) {
        // row sum = both halves' partials.  One slot suffices: the next tile's write comes after at least one of its
        // block-level pair barriers, which the partner only reaches after this read.
        x_sum[half * 128 + row] = l_run;
        uint32_t o0[32];
        tmem_ld_32x32(t_lane + PP_TMEM_O + 32 * half, o0);
        pair_sync(pair_id);
        l_run += x_sum[(half ^ 1) * 128 + row];
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l_run;
          if (a.lse && !half) a.lse[static_cast<int64_t>(item) * PP_ROWS + tok] = fmaf(m_run, sc, log2f(l_run));
          uint32_t ob[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) ob[e] = pack_bf16(__uint_as_float(o0[2 * e]) * inv, __uint_as_float(o0[2 * e + 1]) * inv);
          if (a.use_cls && tok == N) {   // the cls query: fp32 per frame, and frame 0 doubles as the projection's cls input row
            float4* dst = reinterpret_cast<float4*>(a.out_cls + (static_cast<int64_t>(b) * T + t) * D + h * 64 + 32 * half);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              dst[e] = make_float4(__uint_as_float(o0[4 * e]) * inv, __uint_as_float(o0[4 * e + 1]) * inv,
                                   __uint_as_float(o0[4 * e + 2]) * inv, __uint_as_float(o0[4 * e + 3]) * inv);
            if (t == 0) {                // vit.py:198
              uint4* dc = reinterpret_cast<uint4*>(a.out + (a.cls_row0 + b) * a.ld_out + h * 64 + 32 * half);
#pragma unroll
              for (int e = 0; e < 4; ++e) dc[e] = make_uint4(ob[4 * e], ob[4 * e + 1], ob[4 * e + 2], ob[4 * e + 3]);
            }
          } else {                       // patch row -> swizzled staging row (16-byte chunk c at c ^ (row & 7))
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t dst = stage + row * 128 + (((4 * half + e) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(ob[4 * e]), "r"(ob[4 * e + 1]),
                           "r"(ob[4 * e + 2]), "r"(ob[4 * e + 3])
                           : "memory");
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      group_sync(g);
      if (gw == 0 && elect_one()) {
        if (128 * j < N) {               // rows past N are clipped by the tensor map (the cls row is not part of it)
          tma_store_4d(&tmOut, stage, h * 64, t, 128 * j, b);
          tma_commit_group();
          tma_wait_group_read<0>();
        }
        mbar_arrive(q_empty(slot));
      }
    }
    if (gw == 0 && elect_one()) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, PP_TMEM_COLS);
}

// Launch for N + use_cls <= 304; lse != nullptr also writes the per-row log-sum-exp (training).
int launch_spatial_pp(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                      int heads, int use_cls, int64_t cls_row0, cudaStream_t stream, float* lse) {
  alignas(64) CUtensorMap tmQf, tmQt, tmKVf, tmKVt, tmO;
  const int cols = 3 * heads * 64;
  const int q_tail = N % 128, kv_full = N < 256 ? N : 256, kv_tail = N - kv_full;
  int rc;
  if ((rc = make_patch_tmap(&tmQf, qkv, ld_qkv, cols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&tmQt, qkv, ld_qkv, cols, B, N, T, q_tail > 0 ? q_tail : 1))) return rc;
  if ((rc = make_patch_tmap(&tmKVf, qkv, ld_qkv, cols, B, N, T, kv_full))) return rc;
  if ((rc = make_patch_tmap(&tmKVt, qkv, ld_qkv, cols, B, N, T, kv_tail > 0 ? kv_tail : 1))) return rc;
  if ((rc = make_patch_tmap(&tmO, out, ld_out, heads * 64, B, N, T, N >= 128 ? 128 : N))) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  PpArgs a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out), ld_out, out_cls,
           B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, lse};
  const int items = B * T * heads;
  const int sms = sm_count();
  const int grid = items < sms ? items : sms;
  attn_spatial_pp_kernel<<<grid, PP_THREADS, PP_SMEM, stream>>>(tmQf, tmQt, tmKVf, tmKVt, tmO, a);
  return check_launch("attn_spatial_pp_kernel");
}

}  // namespace tcow
