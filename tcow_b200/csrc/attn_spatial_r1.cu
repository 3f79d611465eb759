// Spatial attention on tcgen05 + TMEM, resident K/V, ONE persistent CTA per SM, two query tiles in flight, and every
// score read from TMEM exactly ONCE: one softmax thread per query row keeps the 128 scores of a key block in registers
// between the maximum and the exponentials.
//
// Work item = (clip b, frame t, head h): full softmax attention over S = N (+1 cls) <= 304 tokens of head dim 64
// (vit.py:78-111 as called at vit.py:186 on the tokens assembled at vit.py:179-185).  The CTA's stream of 128-query tiles
// alternates between two groups; each group owns 256 TMEM columns: scores [0,128), P (bf16) [128,192), O [192,256).
//   warp 0      TMA producer: K/V of item i into buffer i&1 (4-D gather of the strided canonical rows, SWIZZLE_128B; the cls
//               row is appended by hand), Q tiles into a 4-slot ring
//   warp 1 / 2  MMA issuer of group 0 / 1:  S_b = Q K_b^T for key blocks of 128 (SS), O (+)= P_b V_b (A = P from TMEM, V
//               MN-major).  S_{b+1} is issued as soon as the softmax threads have READ S_b (s_free), i.e. it runs under the
//               exponentials of block b; P has its own columns, so nothing waits for the P V product but the P store of the
//               next block (p_free) and the output.
//   warp 3      TMEM allocator (all 512 columns)
//   warps 4-7 / 8-11  softmax group 0 / 1, one thread per query row: eight tcgen05.ld behind ONE wait (the ~190-cycle TMEM
//               round trip is paid once per block instead of ten times: profiles/r02_notes.md), exact running maximum,
//               exponentials straight from registers, P (bf16) -> TMEM, O rescaled in TMEM only when a row's maximum
//               moved; O / l -> bf16 -> the tile's (dead) Q slot as a swizzled staging tile -> ONE TMA store.
// The softmax warpgroups take the registers the service warpgroup does not need (setmaxnreg 72 / 216).
// Lane quarters wholly past S (the last tile holds 45 of 128 rows at S = 301) skip the softmax and only keep the barriers
// going.
#include <math.h>
#include <stdlib.h>

#include "attn_r1.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

template <int NL>   // 16-key chunks in the last key block (1..8): both block bodies are straight-line code
__global__ void __launch_bounds__(R1_THREADS, 1)
attn_spatial_r1_kernel(const __grid_constant__ CUtensorMap tmQfull, const __grid_constant__ CUtensorMap tmQtail,
                       const __grid_constant__ CUtensorMap tmKVfull, const __grid_constant__ CUtensorMap tmKVtail,
                       const __grid_constant__ CUtensorMap tmOut, const R1Args a) {
  extern __shared__ uint8_t smem_r1[];
  const uint32_t raw = smem_u32(smem_r1);
  const uint32_t base = (raw + 1023u) & ~1023u;
  auto k_buf = [&](int b) { return base + b * 2 * R1_KV_BYTES; };
  auto v_buf = [&](int b) { return base + b * 2 * R1_KV_BYTES + R1_KV_BYTES; };
  auto q_buf = [&](int slot) { return base + 4 * R1_KV_BYTES + slot * R1_QTILE_BYTES; };
  const uint32_t bars = base + 4 * R1_KV_BYTES + R1_QSLOTS * R1_QTILE_BYTES;
  auto kv_full = [&](int b) { return bars + 8u * b; };
  auto kv_empty = [&](int b) { return bars + 16u + 8u * b; };
  auto q_full = [&](int s) { return bars + 32u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 64u + 8u * s; };
  auto s_full = [&](int g) { return bars + 96u + 8u * g; };    // MMA -> softmax: scores of a block are in TMEM
  auto s_free = [&](int g) { return bars + 112u + 8u * g; };   // softmax -> MMA: the scores are in registers
  auto p_full = [&](int g) { return bars + 128u + 8u * g; };   // softmax -> MMA: P of a block is in TMEM
  auto p_free = [&](int g) { return bars + 144u + 8u * g; };   // MMA -> softmax: P V of a block has completed
  auto o_full = [&](int g) { return bars + 160u + 8u * g; };   // MMA -> softmax: the tile's O is final
  const uint32_t tmem_slot = bars + 176;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_r1 + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int nblk = (S16 + R1_KB - 1) / R1_KB;                          // key blocks of 128 (the last one shorter)
  const int nk_last = S16 - R1_KB * (nblk - 1);
  const int D = heads * 64;
  const int items = a.B * T * heads;
  const int my_items = items > static_cast<int>(blockIdx.x) ? (items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const uint32_t NT = static_cast<uint32_t>(my_items) * nq;           // this CTA's tile stream
  const int kv_full_rows = N < 256 ? N : 256, kv_tail_rows = N - kv_full_rows;

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
      mbar_init(kv_empty(b), nq >= 2 ? 2 : 1);   // one commit per MMA warp that has tiles in the item
      mbar_init(s_full(b), 1);
      mbar_init(s_free(b), 4);                   // one arrival per softmax warp
      mbar_init(p_full(b), 4);
      mbar_init(p_free(b), 1);
      mbar_init(o_full(b), 1);
    }
    for (int s = 0; s < R1_QSLOTS; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, R1_TMEM_COLS);
    tmem_relinquish();
  }
  // Rows [S, 304) of K/V are never written by TMA: zero them once in both buffers (P is 0 there, V must be finite).
  for (int idx = threadIdx.x; idx < (R1_ROWS - S) * 8 * 4; idx += blockDim.x) {
    const int which = idx / ((R1_ROWS - S) * 8), rem = idx % ((R1_ROWS - S) * 8);
    const int row = S + (rem >> 3), chunk = rem & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + which * R1_KV_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4)),
                 "r"(0)
                 : "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    setmaxnreg_dec<72>();   // frees 128 x 96 registers: exactly the 256 x 48 the softmax warpgroups add
    if (warp == 0) {
      // ---------------------------------------------------------------- producer
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
          const int slot = n % R1_QSLOTS;
          mbar_wait(q_empty(slot), ((n / R1_QSLOTS) & 1) ^ 1);   // released by the output store of the tile 4 back
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
    } else if (warp <= 2) {
      // ---------------------------------------------------------------- MMA issuer of group g = warp - 1
      const int g = warp - 1;
      const uint32_t region = tmem_base + g * R1_REGION;
      const uint32_t idesc_full = umma_idesc_bf16(128, R1_KB);
      const uint32_t idesc_last = umma_idesc_bf16(128, nk_last);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      uint32_t kc = 0;   // blocks issued so far by this group (= phase index of s_free / p_full)
      // S_b of tile n into the group's score columns
      auto issue_s = [&](uint32_t n, int b) {
        const uint32_t ii = n / nq;
        const int slot = n % R1_QSLOTS, kb = ii & 1;
        if (b == 0) {
          mbar_wait(kv_full(kb), (ii >> 1) & 1);
          mbar_wait(q_full(slot), (n / R1_QSLOTS) & 1);
        }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qd = umma_desc_k_sw128(q_buf(slot)), kd = umma_desc_k_sw128(k_buf(kb) + b * R1_KB * 128);
          const uint32_t idesc = b == nblk - 1 ? idesc_last : idesc_full;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(region, qd + 2u * k, kd + 2u * k, idesc, k > 0 ? 1u : 0u);
          umma_commit(s_full(g));
        }
        __syncwarp();
      };
      if (static_cast<uint32_t>(g) < NT) issue_s(g, 0);
      for (uint32_t n = g; n < NT; n += 2) {
        const uint32_t ii = n / nq;
        const int j = n % nq, kb = ii & 1;
        for (int b = 0; b < nblk; ++b, ++kc) {
          // the next scores (of this tile, or the first block of the group's next tile) as soon as S_b has been read.  With
          // one tile per item (nq == 1) the group's next tile is two items on, in the SAME K/V buffer, which is only refilled
          // after this tile's P V: there the next scores are issued behind it (below).
          const bool next_in_tile = b + 1 < nblk, next_tile = !next_in_tile && n + 2 < NT;
          if (next_in_tile || (next_tile && nq >= 2)) {
            mbar_wait(s_free(g), kc & 1);
            if (next_in_tile) issue_s(n, b + 1);
            else issue_s(n + 2, 0);
          }
          mbar_wait(p_full(g), kc & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t vd = umma_desc_mn_sw128(v_buf(kb) + b * R1_KB * 128, 1024);
            const int nkk = (b == nblk - 1 ? nk_last : R1_KB) >> 4;
            for (int kk = 0; kk < nkk; ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
              umma_bf16_ts(region + R1_TMEM_O, region + R1_TMEM_P + 8 * kk, vd + 128u * kk, idesc_o, (b | kk) != 0 ? 1u : 0u);
            umma_commit(p_free(g));
            if (b == nblk - 1) {
              umma_commit(o_full(g));
              if (j + 2 >= nq) umma_commit(kv_empty(kb));   // this group's last tile of the item
            }
          }
          __syncwarp();
          if (next_tile && nq < 2) {
            mbar_wait(s_free(g), kc & 1);
            issue_s(n + 2, 0);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output: one thread per query row
    setmaxnreg_inc<216>();
    const int g = (warp - 4) >> 2;                 // softmax group = parity of the tiles it takes
    const int gw = (warp - 4) & 3;                 // warp within the group = TMEM lane quarter (warp id % 4)
    const int row = gw * 32 + lane;
    const uint32_t region = tmem_base + g * R1_REGION;
    const uint32_t t_lane = region + (static_cast<uint32_t>(gw * 32) << 16);
    const float sc = a.scale_log2;
    uint32_t kc = 0, o_ct = 0;
    for (uint32_t n = g; n < NT; n += 2, ++o_ct) {
      const uint32_t ii = n / nq;
      const int j = n % nq, slot = n % R1_QSLOTS;
      const int tok = 128 * j + row;
      const bool valid = tok < S;
      const bool live = 128 * j + gw * 32 < S;   // warp-uniform: a lane quarter wholly past S only keeps the barriers going
      float m_run = -INFINITY, l_run = 0.f;
      for (int blk = 0; blk < nblk - 1; ++blk)
        r1_softmax_block<8, false>(t_lane, s_full(g), s_free(g), p_full(g), p_free(g), kc, m_run, l_run, live, valid, lane, S,
                                   S16, sc);
      r1_softmax_block<NL, true>(t_lane, s_full(g), s_free(g), p_full(g), p_free(g), kc, m_run, l_run, live, valid, lane, S, S16,
                                 sc);
      // ---- O / l -> bf16 -> staging tile (this tile's Q slot: Q is dead once the last Q K^T has completed) -> TMA store
      const uint32_t stage = q_buf(slot);
      const int item = blockIdx.x + ii * gridDim.x;   // the divisions run under the wait for the last P V
      const int h = item % heads, t = (item / heads) % T, b = item / (heads * T);
      mbar_wait(o_full(g), o_ct & 1);
      tc_fence_after();
      if (live) {
        const float inv = 1.0f / l_run;
        if (valid && a.lse) a.lse[static_cast<int64_t>(item) * R1_ROWS + tok] = fmaf(m_run, sc, log2f(l_run));
        uint32_t oo[2][32];
        tmem_ld_32x32(t_lane + R1_TMEM_O, oo[0]);
        tmem_ld_32x32(t_lane + R1_TMEM_O + 32, oo[1]);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t (&o0)[32] = oo[half];
          if (valid) {
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
  if (warp == 3) tmem_dealloc(tmem_base, R1_TMEM_COLS);
}

// Launch for N + use_cls <= 304; lse != nullptr also writes the per-row log-sum-exp (training).
int launch_spatial_r1(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
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
  const int S16 = (N + use_cls + 15) & ~15;
  const int nl = (S16 - R1_KB * ((S16 + R1_KB - 1) / R1_KB - 1)) >> 4;   // chunks of the last key block (1..8)
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, R1Args);
  static const Kern kerns[8] = {attn_spatial_r1_kernel<1>, attn_spatial_r1_kernel<2>, attn_spatial_r1_kernel<3>,
                                attn_spatial_r1_kernel<4>, attn_spatial_r1_kernel<5>, attn_spatial_r1_kernel<6>,
                                attn_spatial_r1_kernel<7>, attn_spatial_r1_kernel<8>};
  const Kern kern = kerns[nl - 1];
  static bool configured[64][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63][nl - 1]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R1_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63][nl - 1] = true;
  }
  R1Args a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out), ld_out, out_cls,
           B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, lse};
  const int items = B * T * heads;
  const int sms = sm_count();
  const int grid = items < sms ? items : sms;
  kern<<<grid, R1_THREADS, R1_SMEM, stream>>>(tmQf, tmQt, tmKVf, tmKVt, tmO, a);
  return check_launch("attn_spatial_r1_kernel");
}

}  // namespace tcow
