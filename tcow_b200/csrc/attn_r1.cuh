// Shared by the two tcgen05 spatial-attention kernels with register-resident scores (attn_spatial_r1.cu: K/V resident in
// shared memory, S <= 304; attn_spatial_rs.cu: K/V streamed, any S): the TMEM layout of a softmax group and the softmax of
// one key block by one thread per query row.  Reference: vit.py:78-111 as called at vit.py:186.
#pragma once
#include <math.h>

#include "ptx.cuh"

namespace tcow {
namespace {
constexpr int R1_ROWS = 304;                  // K/V rows (S rounded up to 16)
constexpr int R1_KV_BYTES = R1_ROWS * 128;    // one of K or V: 38912 (multiple of 1024)
constexpr int R1_QTILE_BYTES = 128 * 128;
constexpr int R1_QSLOTS = 4;
constexpr int R1_THREADS = 384;
constexpr int R1_KB = 128;                    // keys per block
constexpr int R1_TMEM_P = 128;                // P columns [128, 192) of a region
constexpr int R1_TMEM_O = 192;                // O accumulator columns [192, 256)
constexpr int R1_REGION = 256;                // TMEM columns per group
constexpr int R1_TMEM_COLS = 512;
constexpr int R1_BAR_BYTES = 256;
constexpr int R1_SMEM = 2 * 2 * R1_KV_BYTES + R1_QSLOTS * R1_QTILE_BYTES + R1_BAR_BYTES + 1024;

struct R1Args {
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
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void group_sync(int group) {   // named barriers 1 / 2: the four warps of a softmax group
  if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// One key block of NCH 16-key chunks (compile time: 8, or NL for the last block).  LAST: the block that may hold keys
// past S — they all sit in its final chunk (S16 - S < 16) and are set to -inf, so p = 0 without a second code path.
template <int NCH, bool LAST>
__device__ __forceinline__ void r1_softmax_block(const uint32_t t_lane, const uint32_t bar_s_full, const uint32_t bar_s_free,
                                                 const uint32_t bar_p_full, const uint32_t bar_p_free, uint32_t& kc,
                                                 float& m_run, float& l_run, const bool live, const bool valid,
                                                 const int lane, const int S, const int S16, const float sc) {
  mbar_wait(bar_s_full, kc & 1);
  tc_fence_after();
  if (!live) {
    // a lane quarter wholly past S: keep the four barriers of the block going.  It waits for p_free like everybody else,
    // which also keeps it from arriving twice in one phase of p_full.
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_s_free);
    if (kc > 0) mbar_wait(bar_p_free, (kc - 1) & 1);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p_full);
    ++kc;
    return;
  }
  uint32_t v[NCH][16];
#pragma unroll
  for (int c = 0; c < NCH; ++c) tmem_ld_32x16(t_lane + 16 * c, v[c]);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_s_free);     // the scores are in registers: the MMA warp may overwrite them
  if (LAST && S16 > S) {
    const int k0 = S16 - 16;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (k0 + e >= S) v[NCH - 1][e] = 0xff800000u;
  }
  // ---- exact maximum: four independent FMNMX3 chains
  float m0 = m_run, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int e = 0; e < 16; e += 8) {
      m0 = max3(m0, __uint_as_float(v[c][e]), __uint_as_float(v[c][e + 1]));
      m1 = max3(m1, __uint_as_float(v[c][e + 2]), __uint_as_float(v[c][e + 3]));
      m2 = max3(m2, __uint_as_float(v[c][e + 4]), __uint_as_float(v[c][e + 5]));
      m3 = max3(m3, __uint_as_float(v[c][e + 6]), __uint_as_float(v[c][e + 7]));
    }
  }
  const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  const bool moved = m_run != -INFINITY && valid && (mx > m_run);
  const float alpha = moved ? ex2f((m_run - mx) * sc) : 1.f;
  l_run *= alpha;
  m_run = mx;
  // ---- p = 2^(s*sc - mx*sc) from registers (each chunk's 16 scores die into 8 packed bf16 pairs), row sum
  uint32_t pk[NCH][8];
  {
    const float mxs = mx * sc;
    const uint64_t sc2 = f2_pack(sc, sc), nm2 = f2_pack(-mxs, -mxs);
    uint64_t la = f2_pack(0.f, 0.f), lb = la;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float x0, x1, x2, x3;
        f2_unpack(f2_fma(f2_pack_u(v[c][e], v[c][e + 1]), sc2, nm2), x0, x1);
        f2_unpack(f2_fma(f2_pack_u(v[c][e + 2], v[c][e + 3]), sc2, nm2), x2, x3);
        const float p0 = ex2f(x0), p1 = ex2f(x1), p2 = ex2f(x2), p3 = ex2f(x3);
        la = f2_add(la, f2_pack(p0, p1));
        lb = f2_add(lb, f2_pack(p2, p3));
        pk[c][e >> 1] = pack_bf16(p0, p1);
        pk[c][(e >> 1) + 1] = pack_bf16(p2, p3);
      }
    }
    float s0, s1;
    f2_unpack(f2_add(la, lb), s0, s1);
    l_run += s0 + s1;
  }
  // ---- the P columns (and O) are free once the previous block's P V has completed: it was issued a TMEM round trip, a
  // maximum pass and an exponential pass ago
  if (kc > 0) mbar_wait(bar_p_free, (kc - 1) & 1);
  tc_fence_after();
  // exact online softmax: O (the partial result of the earlier blocks) is rescaled by 2^(m_old - m_new) when a row's
  // maximum moved
  if (__any_sync(0xffffffffu, moved)) {
    const uint64_t al2 = f2_pack(alpha, alpha);
    uint32_t o[4][16];
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) tmem_ld_32x16(t_lane + R1_TMEM_O + 16 * hh, o[hh]);
    tmem_ld_wait();
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        float r0, r1;
        f2_unpack(f2_mul(f2_pack_u(o[hh][e], o[hh][e + 1]), al2), r0, r1);
        o[hh][e] = __float_as_uint(r0);
        o[hh][e + 1] = __float_as_uint(r1);
      }
      tmem_st_32x16(t_lane + R1_TMEM_O + 16 * hh, o[hh]);
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) tmem_st_32x8(t_lane + R1_TMEM_P + 8 * c, pk[c]);
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_p_full);   // one arrival per warp
  ++kc;
}

}  // namespace
}  // namespace tcow
