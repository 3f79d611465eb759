// Fused temporal sub-block front half:  O = TemporalAttention( A W_qkv^T + b )  in ONE kernel.
//
// Replaces, for the temporal branch of Block.forward (vit.py:169-173), the qkv Linear (vit.py:81) AND the whole
// Attention.forward body (vit.py:82-109, causal mask :93-99) — the 3*D-wide qkv tensor never goes to HBM.
//
// A GEMM tile is 128 rows x 192 columns: the rows are `seqs_per_tile` complete temporal sequences of T consecutive
// token rows (4 x 30 = 120 rows at T=30; rows beyond are padding), the columns are [q | k | v] (64 each) of ONE head
// (the weight rows are permuted head-major on the host).  Mainloop = the tcgen05/TMA pipeline of gemm_tcgen05.cu
// (M=128, N=192, fp32 accumulator in TMEM, two accumulator stages).  Epilogue (4 warps): accumulator + bias -> bf16
// q/k/v tiles in swizzled shared memory -> each warp runs the causal softmax attention of one sequence with
// mma.sync (attn_frag.cuh; 30x30 problems are far below a tcgen05 tile) -> normalised output rows -> HBM.
// Numerics are identical to running the two kernels back to back (q, k, v are rounded to bf16 exactly once).
#include "attn_frag.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

constexpr int FQ_BM = 128, FQ_BN = 192, FQ_BK = 64;
constexpr int FQ_STAGES = 4;
constexpr int FQ_A_BYTES = FQ_BM * 128;
constexpr int FQ_B_BYTES = FQ_BN * 128;
constexpr int FQ_STAGE_BYTES = FQ_A_BYTES + FQ_B_BYTES;
constexpr int FQ_TILE_BYTES = FQ_BM * 128;  // one of the q / k / v staging tiles
constexpr int FQ_SMEM = FQ_STAGES * FQ_STAGE_BYTES + 3 * FQ_TILE_BYTES + 256 + 1024;
constexpr int FQ_TMEM_COLS = 512;  // 2 accumulator stages x 192 columns, rounded up to a power of two

struct FqArgs {
  const float* bias;      // [heads * 192], permuted like the weights
  __nv_bfloat16* out;     // [M, heads * 64]
  int64_t ld_out;
  int num_seq, T, heads, K, seqs_per_tile, causal_diag;
  float scale_log2;
};

template <int T_PAD>
__global__ void __launch_bounds__(256, 1)
qkv_tattn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FqArgs a) {
  extern __shared__ uint8_t smem_fq[];
  const uint32_t raw = smem_u32(smem_fq);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t s_q = base + FQ_STAGES * FQ_STAGE_BYTES, s_k = s_q + FQ_TILE_BYTES, s_v = s_k + FQ_TILE_BYTES;
  const uint32_t bars = s_v + FQ_TILE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (FQ_STAGES + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * FQ_STAGES + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * FQ_STAGES + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * FQ_STAGES + 4);
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 128);
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
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(FQ_BM, FQ_BN);
    uint32_t it = 0, t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      mbar_wait(tempty_bar(acc), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
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
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: q/k/v tile -> smem -> per-sequence attention -> HBM
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const bool live = row < rows_per_tile;  // padding rows are staged as zeros (V must stay finite)
    const uint32_t sw = row & 7;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m_blk = tile / heads, h = tile % heads;
      const int acc = t & 1;
      mbar_wait(tfull_bar(acc), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 256;
      const float* bias_h = a.bias + h * FQ_BN;
#pragma unroll 1
      for (int part = 0; part < 3; ++part) {
        const uint32_t dst = (part == 0 ? s_q : (part == 1 ? s_k : s_v)) + row * 128;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_row + part * 64, v0);
        tmem_ld_32x32(t_row + part * 64 + 32, v1);
        tmem_ld_wait();
        if (part == 2) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));  // accumulator drained: the MMA warp may start the tile after next
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* v = (j < 4) ? (v0 + 8 * j) : (v1 + 8 * (j - 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_h + part * 64) + 2 * j);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_h + part * 64) + 2 * j + 1);
          uint32_t p0 = pack_bf16(__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y);
          uint32_t p1 = pack_bf16(__uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w);
          uint32_t p2 = pack_bf16(__uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y);
          uint32_t p3 = pack_bf16(__uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w);
          if (!live) p0 = p1 = p2 = p3 = 0u;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + ((j ^ sw) << 4)), "r"(p0), "r"(p1), "r"(p2),
                       "r"(p3)
                       : "memory");
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the three tiles are complete
      const int seq0 = m_blk * spt;
      for (int s = ew; s < spt && seq0 + s < a.num_seq; s += 4) {
        const int r0 = s * T;
        temporal_attend_seq<T_PAD>(s_q, s_k, s_v, r0, T, a.causal_diag, a.scale_log2);
        __nv_bfloat16* dst = a.out + (static_cast<int64_t>(seq0 + s) * T) * a.ld_out + h * HD;
        for (int idx = lane; idx < T * 8; idx += 32) {
          const int r = idx >> 3, chunk = idx & 7;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(sw_addr(s_q, r0 + r, chunk)));
          *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(r) * a.ld_out + chunk * 8) = v;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with the tiles before they are refilled
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, FQ_TMEM_COLS);
}

template <int T_PAD>
static int launch_fq(const void* A, int64_t lda, const void* W, int64_t ldw, const FqArgs& a, int M, cudaStream_t stream) {
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  const int rows_per_tile = a.seqs_per_tile * a.T;
  if ((rc = make_tmap_2d(&tmA, false, A, a.K, M, lda, FQ_BK, rows_per_tile))) return rc;
  if ((rc = make_tmap_2d(&tmB, false, W, a.K, static_cast<uint64_t>(a.heads) * FQ_BN, ldw, FQ_BK, FQ_BN))) return rc;
  auto kern = qkv_tattn_kernel<T_PAD>;
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
  const int T_PAD = T <= 32 ? 32 : 64;
  // every sequence's T_PAD-row window must stay inside the 128-row tile
  FqArgs a{bias_perm, static_cast<__nv_bfloat16*>(out), ld_out, num_seq, T, heads, K, (128 - T_PAD) / T + 1, causal_diag,
           0.125f * 1.4426950408889634f};
  const long long M = static_cast<long long>(num_seq) * T;
  if (M > 0x7fffffffLL) return set_error(TCOW_ERR_ARG, "qkv_temporal_attn: too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T_PAD == 32) return launch_fq<32>(A, lda, Wperm, ldw, a, static_cast<int>(M), s);
  return launch_fq<64>(A, lda, Wperm, ldw, a, static_cast<int>(M), s);
}
