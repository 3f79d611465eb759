// Fused temporal sub-block front half:  O = TemporalAttention( A W_qkv^T + b )  in ONE kernel, all on tcgen05.
//
// Replaces, for the temporal branch of Block.forward (vit.py:169-173), the qkv Linear (vit.py:81) AND the whole
// Attention.forward body (vit.py:82-109, causal mask :93-99) — the 3*D-wide qkv tensor never goes to HBM.
//
// A GEMM tile is 128 rows x 192 columns: the rows are `seqs_per_tile` complete temporal sequences of T consecutive
// token rows (4 x 30 = 120 rows at T=30; rows beyond are padding), the columns are [q | k | v] (64 each) of ONE head
// (the weight rows are permuted head-major on the host).
//   mainloop (warps 0/1) : the TMA + tcgen05 pipeline of gemm_tcgen05.cu (M=128, N=192, fp32 accumulators in TMEM x2)
//   epilogue (warps 4-7) : accumulator + bias -> bf16 q / k / v tiles in swizzled smem (exactly the bytes the unfused
//                          path would have written to HBM), then one query row per thread:
//       S = Q K^T  for the whole tile at once (128x128x64, tcgen05 SS; only the T x T diagonal blocks are used —
//                  the off-diagonal waste is ~0.1 us of tensor time and saves all warp-level MMA work),
//       masked softmax over the row's own sequence straight out of TMEM, P (bf16, zero off the block) back into TMEM,
//       O = P V    (128x64x128, tcgen05 with A = P from TMEM, V as an MN-major smem operand), O / l -> HBM.
//   warp 3 issues the two small attention MMAs so that the mainloop issuer never waits on the epilogue.
// The epilogue is software-pipelined over three tiles so that the two tensor-pipe round trips of a tile (S, then P V —
// each queued behind up to three k-blocks of mainloop MMAs) hide behind other tiles' work:
//   iteration i:  output(i-2)  |  drain accumulator(i) -> q/k/v smem[i&1] -> issue S(i)  |  softmax(i-1) -> issue PV(i-1)
// TMEM: accumulator [0,192) (drained into registers in one go, so a single stage costs the mainloop ~300 cycles per
// tile), S/P/O regions [192,320) and [320,448).
#include <math.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

constexpr int FQ_BM = 128, FQ_BN = 192, FQ_BK = 64;
constexpr int FQ_STAGES = 3;
constexpr int FQ_A_BYTES = FQ_BM * 128;
constexpr int FQ_B_BYTES = FQ_BN * 128;
constexpr int FQ_STAGE_BYTES = FQ_A_BYTES + FQ_B_BYTES;
constexpr int FQ_TILE_BYTES = FQ_BM * 128;  // one of the q / k / v staging tiles
constexpr int FQ_QKV_BYTES = 3 * FQ_TILE_BYTES;  // one q|k|v staging buffer; two of them (software-pipelined epilogue)
constexpr int FQ_SMEM = FQ_STAGES * FQ_STAGE_BYTES + 2 * FQ_QKV_BYTES + 256 + 1024;
constexpr int FQ_TMEM_COLS = 512;
constexpr int FQ_TMEM_S = 192;  // two S regions of 128 columns: S -> P (first 64) and O (last 64)

struct FqArgs {
  const float* bias;      // [heads * 192], permuted like the weights
  __nv_bfloat16* out;     // [M, heads * 64]
  int64_t ld_out;
  int num_seq, T, heads, K, seqs_per_tile, causal_diag;
  float scale_log2;
};

__device__ __forceinline__ float fq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 1)
qkv_tattn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FqArgs a) {
  extern __shared__ uint8_t smem_fq[];
  const uint32_t raw = smem_u32(smem_fq);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t s_qkv = base + FQ_STAGES * FQ_STAGE_BYTES;  // buffer b: q at +b*FQ_QKV_BYTES, then k, then v
  const uint32_t bars = s_qkv + 2 * FQ_QKV_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (FQ_STAGES + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * FQ_STAGES), tempty_bar = tfull_bar + 8;
  auto qkv_ready = [&](int b) { return tfull_bar + 16u + 8u * b; };
  auto s_ready = [&](int b) { return tfull_bar + 32u + 8u * b; };
  auto p_ready = [&](int b) { return tfull_bar + 48u + 8u * b; };
  auto o_ready = [&](int b) { return tfull_bar + 64u + 8u * b; };
  const uint32_t tmem_slot = tfull_bar + 80u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_fq + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T, heads = a.heads, spt = a.seqs_per_tile;
  const int rows_per_tile = spt * T;
  const int num_m = (a.num_seq + spt - 1) / spt;
  const int num_tiles = num_m * heads;
  const int num_kb = a.K / FQ_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < FQ_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(qkv_ready(i), 128);
      mbar_init(s_ready(i), 1);
      mbar_init(p_ready(i), 128);
      mbar_init(o_ready(i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, FQ_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: A rows of `spt` sequences, W rows of one head
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / heads, h = tile % heads;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % FQ_STAGES;
        mbar_wait(empty_bar(s), ((it / FQ_STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), static_cast<uint32_t>(rows_per_tile) * 128u + FQ_B_BYTES);
          const uint32_t sa = base + s * FQ_STAGE_BYTES;
          tma_load_2d(sa, &tmA, kb * FQ_BK, m_blk * rows_per_tile, full_bar(s));
          tma_load_2d(sa + FQ_A_BYTES, &tmB, kb * FQ_BK, h * FQ_BN, full_bar(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ mainloop MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(FQ_BM, FQ_BN);
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      mbar_wait(tempty_bar, (t & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % FQ_STAGES;
        mbar_wait(full_bar(s), (it / FQ_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + s * FQ_STAGE_BYTES;
          const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + FQ_A_BYTES);
#pragma unroll
          for (int k = 0; k < FQ_BK / 16; ++k)
            umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (kb == num_kb - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ attention MMA issuer: S(j), then P V of the previous tile
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    int n_mine = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) ++n_mine;
    for (int j = 0; j <= n_mine; ++j) {
      if (j < n_mine) {
        const int b = j & 1;
        mbar_wait(qkv_ready(b), (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qd = umma_desc_k_sw128(s_qkv + b * FQ_QKV_BYTES);
          const uint64_t kd = umma_desc_k_sw128(s_qkv + b * FQ_QKV_BYTES + FQ_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + FQ_TMEM_S + 128 * b, qd + 2u * k, kd + 2u * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_ready(b));
        }
        __syncwarp();
      }
      if (j >= 1) {
        const int i = j - 1, b = i & 1;
        mbar_wait(p_ready(b), (i >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = umma_desc_mn_sw128(s_qkv + b * FQ_QKV_BYTES + 2 * FQ_TILE_BYTES, 1024);
          const uint32_t reg = tmem_base + FQ_TMEM_S + 128 * b;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
            umma_bf16_ts(reg + 64, reg + 8u * kk, vd + 128u * kk, idesc_o, kk > 0 ? 1u : 0u);
          umma_commit(o_ready(b));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: one accumulator / query row per thread
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const bool live = row < rows_per_tile;   // padding rows are staged as zeros (V must stay finite)
    const uint32_t sw = row & 7;
    const int sidx = live ? row / T : 0;     // sequence of this row inside the tile
    const int k0 = sidx * T;                 // its keys are tile rows [k0, k0+T)
    const int qi = row - k0;                 // position of the query inside its sequence
    const int kvis = (a.causal_diag < 0 || qi + a.causal_diag + 1 > T) ? T : qi + a.causal_diag + 1;  // visible keys
    // 32-column chunks of S that hold keys of any row of this warp (warp-uniform)
    const int w_lo = ((ew * 32) / T) * T, w_hi_row = (ew * 32 + 31 < rows_per_tile ? ew * 32 + 31 : rows_per_tile - 1);
    const int c_lo = (ew * 32 < rows_per_tile) ? w_lo / 32 : 4;
    const int c_hi = (ew * 32 < rows_per_tile) ? ((w_hi_row / T) * T + T - 1) / 32 : -1;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const float sc = a.scale_log2;
    int n_mine = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) ++n_mine;
    float l_keep[2] = {1.f, 1.f};  // row sums of the tiles in flight (indexed by tile parity)
    for (int i = 0; i < n_mine + 2; ++i) {
      // ---------------- C: output of tile i-2:  O / l -> bf16 -> HBM (128 contiguous bytes per row)
      if (i >= 2) {
        const int j = i - 2, b = j & 1;
        const int tile = blockIdx.x + j * gridDim.x;
        const int m_blk = tile / heads, h = tile % heads;
        mbar_wait(o_ready(b), (j >> 1) & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + FQ_TMEM_S + 128 * b + 64, o0);
        tmem_ld_32x32(t_lane + FQ_TMEM_S + 128 * b + 96, o1);
        tmem_ld_wait();
        tc_fence_before();
        if (live && m_blk * spt + sidx < a.num_seq) {
          const float inv = 1.0f / l_keep[b];
          uint4* dst = reinterpret_cast<uint4*>(a.out + (static_cast<int64_t>(m_blk) * rows_per_tile + row) * a.ld_out + h * 64);
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
      // ---------------- A: drain accumulator of tile i, q / k / v (+bias) -> bf16 -> swizzled smem buffer i&1
      if (i < n_mine) {
        const int b = i & 1;
        const int tile = blockIdx.x + i * gridDim.x;
        const int h = tile % heads;
        const float* bias_h = a.bias + h * FQ_BN;
        mbar_wait(tfull_bar, i & 1);
        tc_fence_after();
        uint32_t v[6][32];
#pragma unroll
        for (int c = 0; c < 6; ++c) tmem_ld_32x32(t_lane + 32 * c, v[c]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tempty_bar);  // the whole accumulator row is in registers: the mainloop may start the next tile
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const uint32_t dst = s_qkv + b * FQ_QKV_BYTES + (c >> 1) * FQ_TILE_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_h + 32 * c) + 2 * j);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_h + 32 * c) + 2 * j + 1);
            uint32_t p0 = pack_bf16(__uint_as_float(v[c][8 * j + 0]) + b0.x, __uint_as_float(v[c][8 * j + 1]) + b0.y);
            uint32_t p1 = pack_bf16(__uint_as_float(v[c][8 * j + 2]) + b0.z, __uint_as_float(v[c][8 * j + 3]) + b0.w);
            uint32_t p2 = pack_bf16(__uint_as_float(v[c][8 * j + 4]) + b1.x, __uint_as_float(v[c][8 * j + 5]) + b1.y);
            uint32_t p3 = pack_bf16(__uint_as_float(v[c][8 * j + 6]) + b1.z, __uint_as_float(v[c][8 * j + 7]) + b1.w);
            if (!live) p0 = p1 = p2 = p3 = 0u;
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + ((((c & 1) * 4 + j) ^ sw) << 4)), "r"(p0),
                         "r"(p1), "r"(p2), "r"(p3)
                         : "memory");
          }
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(qkv_ready(b));
      }
      // ---------------- B: masked softmax of tile i-1 over the row's own sequence, straight out of TMEM
      if (i >= 1 && i - 1 < n_mine) {
        const int j = i - 1, b = j & 1;
        mbar_wait(s_ready(b), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t t_s = t_lane + FQ_TMEM_S + 128 * b;
        float mx = -INFINITY;
        for (int c = c_lo; c <= c_hi; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_s + 32 * c, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int jl = 32 * c + e - k0;  // key position inside the row's sequence
            if (jl >= 0 && jl < kvis) mx = fmaxf(mx, __uint_as_float(v[e]));
          }
        }
        const float mxs = mx * sc;
        float l = 0.f;
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          if (c >= c_lo && c <= c_hi) {
            uint32_t v[32];
            tmem_ld_32x32(t_s + 32 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const int jl = 32 * c + e - k0;
              float p0 = fq_ex2(fmaf(__uint_as_float(v[e]), sc, -mxs));
              float p1 = fq_ex2(fmaf(__uint_as_float(v[e + 1]), sc, -mxs));
              if (!(live && jl >= 0 && jl < kvis)) p0 = 0.f;
              if (!(live && jl + 1 >= 0 && jl + 1 < kvis)) p1 = 0.f;
              l += p0 + p1;
              pk[e >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = 0u;
          }
          tmem_st_32x16(t_s + 16 * c, pk);  // P chunk c over S columns this thread has already consumed
        }
        l_keep[b] = l;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_ready(b));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, FQ_TMEM_COLS);
}

static int launch_fq(const void* A, int64_t lda, const void* W, int64_t ldw, const FqArgs& a, int M, cudaStream_t stream) {
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  const int rows_per_tile = a.seqs_per_tile * a.T;
  if ((rc = make_tmap_2d(&tmA, false, A, a.K, M, lda, FQ_BK, rows_per_tile))) return rc;
  if ((rc = make_tmap_2d(&tmB, false, W, a.K, static_cast<uint64_t>(a.heads) * FQ_BN, ldw, FQ_BK, FQ_BN))) return rc;
  auto kern = qkv_tattn_kernel;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FQ_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const int tiles = ((a.num_seq + a.seqs_per_tile - 1) / a.seqs_per_tile) * a.heads;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, 256, FQ_SMEM, stream>>>(tmA, tmB, a);
  return check_launch("qkv_tattn_kernel");
}

}  // namespace tcow

extern "C" int tcow_qkv_temporal_attn(const void* A, int64_t lda, const void* Wperm, int64_t ldw, const float* bias_perm,
                                      void* out, int64_t ld_out, int num_seq, int T, int heads, int K, int causal_diag,
                                      void* stream) {
  using namespace tcow;
  if (!A || !Wperm || !bias_perm || !out || num_seq <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: bad argument");
  if (T < 1 || T > 64) return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: T=%d unsupported (1..64)", T);
  if (K % 64 != 0) return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: K (%d) must be a multiple of 64", K);
  if ((ld_out % 8) || (reinterpret_cast<uintptr_t>(bias_perm) & 15))
    return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: output pitch must be a multiple of 8, bias 16-byte aligned");
  // as many whole sequences as fit the 128 accumulator rows
  FqArgs a{bias_perm, static_cast<__nv_bfloat16*>(out), ld_out, num_seq, T, heads, K, 128 / T, causal_diag,
           0.125f * 1.4426950408889634f};
  const long long M = static_cast<long long>(num_seq) * T;
  if (M > 0x7fffffffLL) return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return launch_fq(A, lda, Wperm, ldw, a, static_cast<int>(M), s);
}
