// Weight-gradient GEMM for sm_100a:  dW[N1,N2] (fp32) += A[R,N1]^T * B[R,N2]   (A, B bf16 row-major token matrices).
//
// This is the backward of every nn.Linear on the Seeker path (vit.py:50-52,73-74,146; mask_tracker.py:83-86; the
// patch-embed conv of vit.py:233 as an im2col GEMM):  dW = dY^T X with dY [R, out], X [R, in], contraction over the
// R token rows.  Neither operand is transposed in memory: both are fed to tcgen05.mma as MN-MAJOR shared-memory
// operands (the contraction index is the row index), staged by 3-D TMA boxes that split the columns into 64-wide
// swizzle atoms:   smem tile = [64-col block][64 k-rows][64 cols] bf16, SWIZZLE_128B, LBO = 8 KB between blocks.
//
// The output is tiny (<= 9.4 MB) and the contraction long (R ~ 54 000), so the work is split along R: unit =
// (n1 tile, n2 tile, row split); all tiles of one split run concurrently and share the two row slabs through L2;
// partial products are combined with TMA reduce-add into the fp32 gradient (callers zero or accumulate it).
//   warp 0 TMA producer | warp 1 MMA issuer (M=128, N=BN, K=16, fp32 accumulators in TMEM, two stages)
//   warp 2 TMEM allocator | warps 4-7 epilogue: tcgen05.ld -> swizzled staging -> cp.reduce.async.bulk.tensor add
#include <stdlib.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {
constexpr int WG_BM = 128;
constexpr int WG_BK = 64;  // token rows per pipeline stage

template <int BN>
struct WgCfg {
  static constexpr int THREADS = 256;
  static constexpr int A_BYTES = WG_BM * WG_BK * 2;
  static constexpr int B_BYTES = BN * WG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_WARP_BYTES = 32 * 128;
  static constexpr int OUT_BYTES = 4 * 2 * OUT_WARP_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_MAX = 232448;
  static constexpr int STAGES_FIT = (SMEM_MAX - 1024 - BAR_BYTES - OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int SMEM = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int NCHUNK = BN / 32;  // 32 fp32 columns = 128 bytes per staging row
};

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32, BOTH operands MN-major (bits 15 and 16).
__host__ __device__ constexpr uint32_t wg_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
}  // namespace

template <int BN>
__global__ void __launch_bounds__(WgCfg<BN>::THREADS, 1)
gemm_bf16_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, int R, int N1, int N2, int splits, int kb_per_split,
                       float* __restrict__ db, int* __restrict__ sched) {
  using Cfg = WgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_wg[];
  const uint32_t raw = smem_u32(smem_wg);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t s_out = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bars = s_out + Cfg::OUT_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_wg + (tmem_slot - raw));
  // Unit ring: the producer warp decides which unit comes next — statically striped, or (sched != nullptr) by an atomic
  // ticket — and publishes it to the MMA, epilogue and bias warps.  With tickets and several units per SM, SMs that are
  // late or busy with another kernel (NCCL's all-reduce during the data-parallel backward) simply take fewer units
  // instead of leaving a tail: wgrad_fc2 ran 68 % longer on 8 GPUs than on one with static striping.
  constexpr int UR = 4;
  auto ufull_bar = [&](int i) { return bars + 8u * (2 * STAGES + 5 + i); };
  auto uempty_bar = [&](int i) { return bars + 8u * (2 * STAGES + 5 + UR + i); };
  volatile int* s_units = reinterpret_cast<volatile int*>(smem_wg + (bars + 8u * (2 * STAGES + 5 + 2 * UR) - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (N1 + WG_BM - 1) / WG_BM;
  const int num_n = N2 / BN;
  const int tiles = num_m * num_n;
  const int num_kb = (R + WG_BK - 1) / WG_BK;
  const int units = tiles * splits;
  const int unit0 = blockIdx.x, unit_step = gridDim.x;
  // unit -> (tile, split): splits outermost, so the CTAs running at the same time stream the same token rows
  auto unit_at = [&](int u, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    const int sp = u / tiles, tile = u % tiles;
    m_blk = tile / num_n;
    n_blk = tile % num_n;
    kb0 = sp * kb_per_split;
    kb1 = kb0 + kb_per_split < num_kb ? kb0 + kb_per_split : num_kb;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      // bias-gradient mode: the two column-sum warps must have drained the stage too
      mbar_init(empty_bar(s), 1 + (db ? 2 : 0));
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    for (int i = 0; i < UR; ++i) {
      mbar_init(ufull_bar(i), 1);
      mbar_init(uempty_bar(i), 1 + 4 + (db ? 2 : 0));   // MMA warp, four epilogue warps, two bias warps
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // k-th unit of this CTA as published by the producer warp (-1: no more work)
  auto next_unit = [&](uint32_t k) {
    const int slot = k % UR;
    mbar_wait(ufull_bar(slot), (k / UR) & 1);
    const int u = s_units[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(uempty_bar(slot));
    return u;
  };

  if (warp == 0) {
    // ------------------------------------------------ unit scheduler + TMA producer
    uint32_t it = 0;
    for (uint32_t k = 0;; ++k) {
      const int slot = k % UR;
      mbar_wait(uempty_bar(slot), ((k / UR) & 1) ^ 1);
      int u = 0;
      if (lane == 0) {
        u = sched ? atomicAdd(sched, 1) : unit0 + static_cast<int>(k) * unit_step;
        if (u >= units) u = -1;
        s_units[slot] = u;
        mbar_arrive(ufull_bar(slot));
      }
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u < 0) break;
      int m_blk, n_blk, kb0, kb1;
      unit_at(u, m_blk, n_blk, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t sa = base + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          // box = (64 columns, 64 token rows, column blocks); rows / blocks past the tensor are zero-filled
          tma_load_3d(sa, &tmA, 0, kb * WG_BK, m_blk * (WG_BM / 64), full_bar(s));
          tma_load_3d(sa + Cfg::A_BYTES, &tmB, 0, kb * WG_BK, n_blk * (BN / 64), full_bar(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = wg_idesc(WG_BM, BN);
    uint32_t it = 0, t = 0;
    for (;; ++t) {
      const int u = next_unit(t);
      if (u < 0) break;
      int m_blk, n_blk, kb0, kb1;
      unit_at(u, m_blk, n_blk, kb0, kb1);
      const int acc = t & 1;
      mbar_wait(tempty_bar(acc), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + s * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_desc_mn_sw128(sa, WG_BK * 128);
          const uint64_t bdesc = umma_desc_mn_sw128(sa + Cfg::A_BYTES, WG_BK * 128);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k)  // 16 token rows = 2048 bytes per MMA: +128 in addr>>4 units
            umma_bf16(d_tmem, adesc + 128u * k, bdesc + 128u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
    }
  } else if (warp <= 3 && db != nullptr) {
    // ------------------------------------------------ bias gradient: db[n1] += sum over the token rows of dY[:, n1]
    // (nn.Linear bias, vit.py:50-52,73-74).  The dY tile of every stage is already in shared memory for the MMA: warps 2-3
    // add up its 64 rows (thread = two adjacent columns of the 128) for the units of column tile 0 — every n2 tile of a
    // split streams the same dY rows — instead of a separate pass over dY (tcow_colsum_bf16: 84 launches per step).
    const int tcol = (warp - 2) * 32 + lane;                        // columns 2*tcol, 2*tcol+1 of the tile
    const uint32_t col_off = (tcol >> 5) * (WG_BK * 128) + ((tcol & 3) << 2);   // 64-column block, word inside the 16-B chunk
    const uint32_t chunk = (tcol & 31) >> 2;
    uint32_t it = 0;
    for (uint32_t k = 0;; ++k) {
      const int u = next_unit(k);
      if (u < 0) break;
      int m_blk, n_blk, kb0, kb1;
      unit_at(u, m_blk, n_blk, kb0, kb1);
      float s0 = 0.f, s1 = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(s), (it / STAGES) & 1);
        if (n_blk == 0) {
          const uint32_t sa = base + s * Cfg::STAGE_BYTES + col_off;
#pragma unroll 8
          for (int r = 0; r < WG_BK; ++r) {
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sa + r * 128 + ((chunk ^ (r & 7)) << 4)));
            s0 += __uint_as_float(v << 16);
            s1 += __uint_as_float(v & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
      }
      if (n_blk == 0) {
        const int c0 = m_blk * WG_BM + 2 * tcol;
        if (c0 < N1) atomicAdd(db + c0, s0);
        if (c0 + 1 < N1) atomicAdd(db + c0 + 1, s1);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: TMEM -> staging -> TMA reduce-add into dW
    const int ew = warp & 3;
    const uint32_t my_out = s_out + ew * 2 * Cfg::OUT_WARP_BYTES;
    const uint32_t srow = lane * 128;
    const uint32_t sw = lane & 7;
    uint32_t t = 0, cc = 0;
    for (;; ++t) {
      const int u = next_unit(t);
      if (u < 0) break;
      int m_blk, n_blk, kb0, kb1;
      unit_at(u, m_blk, n_blk, kb0, kb1);
      const int acc = t & 1;
      mbar_wait(tfull_bar(acc), (t >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < Cfg::NCHUNK; ++c, ++cc) {
        const uint32_t buf = my_out + (cc & 1) * Cfg::OUT_WARP_BYTES;
        if (elect_one()) tma_wait_group_read<1>();
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (c == Cfg::NCHUNK - 1) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + srow + ((j ^ sw) << 4)), "r"(v[4 * j]),
                       "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_reduce_add_2d(&tmC, buf, n_blk * BN + c * 32, m_blk * WG_BM + ew * 32);  // rows past N1 are clipped
          tma_commit_group();
        }
      }
    }
    __syncwarp();
    if (elect_one()) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (sched && threadIdx.x == 0) {   // the last CTA to leave puts the ticket counters back to zero for the next launch
    __threadfence();
    if (atomicAdd(sched + 1, 1) == static_cast<int>(gridDim.x) - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

// [R, C] row-major bf16 seen as (64 columns, R rows, C/64 column blocks): one box lands in smem as
// [block][row][64] — the MN-major SWIZZLE_128B operand layout.
static int make_mn_tmap(CUtensorMap* m, const void* p, int64_t ld, int R, int C, int blocks) {
  const uint64_t dims[3] = {64, static_cast<uint64_t>(R), static_cast<uint64_t>(C / 64)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, 128};
  const uint32_t box[3] = {64, WG_BK, static_cast<uint32_t>(blocks)};
  return make_tmap_nd(m, false, p, 3, dims, strides, box);
}

template <int BN>
static int launch_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw, int R, int N1,
                        int N2, float* db, int* sched, cudaStream_t stream) {
  using Cfg = WgCfg<BN>;
  alignas(64) CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = make_mn_tmap(&tmA, A, lda, R, N1, WG_BM / 64))) return rc;
  if ((rc = make_mn_tmap(&tmB, B, ldb, R, N2, BN / 64))) return rc;
  if ((rc = make_tmap_2d(&tmC, true, dW, N2, N1, ldw, 32, 32))) return rc;
  auto kern = gemm_bf16_wgrad_kernel<BN>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  // Row splits: fill the SMs (tiles x splits close to a multiple of the SM count) with at least 8 k-blocks each.
  const int tiles = ((N1 + WG_BM - 1) / WG_BM) * (N2 / BN);
  const int num_kb = (R + WG_BK - 1) / WG_BK;
  const int sms = sm_count();
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 64 && (s == 1 || s * 8 <= num_kb); ++s) {
    const int units = tiles * s;
    const double eff = static_cast<double>(units) / (((units + sms - 1) / sms) * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = s;
    }
  }
  if (sched != nullptr) {   // ticket scheduling wants several units per SM (finer row splits), each still >= 8 k-blocks
    const int want = (3 * sms + tiles - 1) / tiles;
    const int cap = num_kb / 8 > 1 ? num_kb / 8 : 1;
    if (want > best) best = want < cap ? want : cap;
  }
  const int per = (num_kb + best - 1) / best;
  const int splits = (num_kb + per - 1) / per;  // no empty split
  const int units = tiles * splits;
  const int grid = units < sms ? units : sms;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(tmA, tmB, tmC, R, N1, N2, splits, per, db, sched);
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "gemm_bf16_wgrad_kernel: launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16_wgrad_kernel");
}

}  // namespace tcow

extern "C" int tcow_gemm_bf16_wgrad_sched(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                                          float* db, int* sched, int R, int N1, int N2, void* stream) {
  using namespace tcow;
  if (!A || !B || !dW) return set_error(TCOW_ERR_ARG, "wgrad: null pointer");
  if (R <= 0 || N1 <= 0 || N2 <= 0) return set_error(TCOW_ERR_ARG, "wgrad: non-positive dimension");
  if (N1 % 64 != 0 || N2 % 64 != 0)
    return set_error(TCOW_ERR_ARG, "wgrad: N1 (%d) and N2 (%d) must be multiples of 64", N1, N2);
  if ((lda % 8) || (ldb % 8) || (ldw % 4)) return set_error(TCOW_ERR_ARG, "wgrad: row pitches must be 16-byte multiples");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // (Clusters of two CTAs sharing the B tile by TMA multicast measured slower for every weight-gradient shape of the model
  // — qkv 1142 vs 1206, fc2 1129 vs 1199 TFLOP/s — and are gone: the long row loop keeps the B slab in L2 anyway.)
  if (N2 % 256 == 0) return launch_wgrad<256>(A, lda, B, ldb, dW, ldw, R, N1, N2, db, sched, s);
  return launch_wgrad<64>(A, lda, B, ldb, dW, ldw, R, N1, N2, db, sched, s);
}

extern "C" int tcow_gemm_bf16_wgrad_bias(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                                         float* db, int R, int N1, int N2, void* stream) {
  return tcow_gemm_bf16_wgrad_sched(A, lda, B, ldb, dW, ldw, db, nullptr, R, N1, N2, stream);
}

extern "C" int tcow_gemm_bf16_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                                    int R, int N1, int N2, void* stream) {
  return tcow_gemm_bf16_wgrad_sched(A, lda, B, ldb, dW, ldw, nullptr, nullptr, R, N1, N2, stream);
}
