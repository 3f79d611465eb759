// Shared pieces of the warp-level (mma.sync) attention kernels: swizzled-smem fragment loaders and the
// single-pass causal attention of one short sequence by one warp (vit.py:88-109 with the mask of :93-99).
#pragma once
#include <math.h>

#include "ptx.cuh"

namespace tcow {

constexpr int HD = 64;          // head dim
constexpr int ROW_BYTES = 128;  // 64 bf16

// smem tile: rows of 128 bytes, 16-byte chunk c of row r stored at chunk (c ^ (r & 7)).
__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int chunk) {
  return base + row * ROW_BYTES + ((chunk ^ (row & 7)) << 4);
}

// A fragments (16 rows x 16 k) of a row-major [row][64] tile: m-tile rows r0.., k-step ks.
__device__ __forceinline__ void load_a_frag(uint32_t base, int r0, int ks, uint32_t (&a)[4]) {
  const int l = lane_id();
  const int mat = l >> 3, r = l & 7;
  ldmatrix_x4(sw_addr(base, r0 + (mat & 1) * 8 + r, ks * 2 + (mat >> 1)), a[0], a[1], a[2], a[3]);
}
// B fragments for S = Q K^T from K stored [key][64]: two n-tiles (keys k0..k0+15), k-step ks.
// returns b[0],b[1] for keys k0..k0+7 and b[2],b[3] for keys k0+8..k0+15.
__device__ __forceinline__ void load_bk_frag(uint32_t base, int k0, int ks, uint32_t (&b)[4]) {
  const int l = lane_id();
  const int mat = l >> 3, r = l & 7;
  ldmatrix_x4(sw_addr(base, k0 + (mat >> 1) * 8 + r, ks * 2 + (mat & 1)), b[0], b[1], b[2], b[3]);
}
// B fragments for O = P V from V stored [key][64]: keys kk0..kk0+15, two n-tiles (d = nd*8 .. nd*8+15).
__device__ __forceinline__ void load_bv_frag(uint32_t base, int kk0, int nd, uint32_t (&b)[4]) {
  const int l = lane_id();
  const int mat = l >> 3, r = l & 7;
  ldmatrix_x4_trans(sw_addr(base, kk0 + (mat & 1) * 8 + r, nd + (mat >> 1)), b[0], b[1], b[2], b[3]);
}

// One warp, one (sequence, head): q, k, v are rows [r0, r0+T) of three [rows][64] bf16 tiles stored with the
// 128-byte XOR swizzle (16-byte chunk c of row r at chunk c ^ (r & 7)); rows [r0+T, r0+T_PAD) of K and V must hold
// finite values (they are masked, but 0 * NaN would poison P V).  On return the normalised output (bf16) sits in
// rows [r0, r0+T) of the Q tile.  T <= T_PAD in {32, 64}; key j is visible to query i iff j <= i + causal_diag
// (causal_diag < 0: no mask).
template <int T_PAD>
__device__ __forceinline__ void temporal_attend_seq(uint32_t sQ, uint32_t sK, uint32_t sV, int r0, int T, int causal_diag,
                                                    float scale_log2) {
  constexpr int NT = T_PAD / 8;   // key n-tiles
  constexpr int KT = T_PAD / 16;  // key k-steps for P V
  const int lane = lane_id();
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll 1
  for (int mh = 0; mh < T_PAD / 32; ++mh) {
    if (mh * 32 >= T) break;
    uint32_t qa[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) load_a_frag(sQ, r0 + mh * 32 + mt * 16, ks, qa[mt][ks]);
    float s[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[mt][nt][c] = 0.f;
#pragma unroll
    for (int np = 0; np < NT / 2; ++np)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b[4];
        load_bk_frag(sK, r0 + np * 16, ks, b);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16_16816(s[mt][2 * np], qa[mt][ks], b[0], b[1]);
          mma_bf16_16816(s[mt][2 * np + 1], qa[mt][ks], b[2], b[3]);
        }
      }
    // ---- mask + softmax (fp32, base-2)
    float inv_l[2][2];
    uint32_t pa[2][KT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = mh * 32 + mt * 16 + g + (c >> 1) * 8;
          const int j = nt * 8 + tq * 2 + (c & 1);
          const bool ok = (j < T) && (causal_diag < 0 || j <= i + causal_diag);
          const float v = ok ? s[mt][nt][c] * scale_log2 : -INFINITY;
          s[mt][nt][c] = v;
          mx[c >> 1] = fmaxf(mx[c >> 1], v);
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float p;   // masked entries are -inf: ex2.approx gives exactly 0
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(s[mt][nt][c] - mx[c >> 1]));
          s[mt][nt][c] = p;
          sum[c >> 1] += p;
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
        inv_l[mt][h] = 1.0f / sum[h];
      }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        pa[mt][kt][0] = pack_bf16(s[mt][2 * kt][0], s[mt][2 * kt][1]);
        pa[mt][kt][1] = pack_bf16(s[mt][2 * kt][2], s[mt][2 * kt][3]);
        pa[mt][kt][2] = pack_bf16(s[mt][2 * kt + 1][0], s[mt][2 * kt + 1][1]);
        pa[mt][kt][3] = pack_bf16(s[mt][2 * kt + 1][2], s[mt][2 * kt + 1][3]);
      }
    }
    // ---- O = P V
    float o[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[mt][nd][c] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        load_bv_frag(sV, r0 + kt * 16, np * 2, b);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16_16816(o[mt][2 * np], pa[mt][kt], b[0], b[1]);
          mma_bf16_16816(o[mt][2 * np + 1], pa[mt][kt], b[2], b[3]);
        }
      }
    // ---- normalise, stage as bf16 in the (already consumed) Q rows of this half
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int lrow = mh * 32 + mt * 16 + g + h * 8;
          const uint32_t v = pack_bf16(o[mt][nd][2 * h] * inv_l[mt][h], o[mt][nd][2 * h + 1] * inv_l[mt][h]);
          if (lrow < T)  // rows past T may belong to the next sequence of a shared tile (qkv_tattn.cu)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw_addr(sQ, r0 + lrow, nd) + tq * 4), "r"(v) : "memory");
        }
  }
  __syncwarp();
}

}  // namespace tcow
