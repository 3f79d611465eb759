// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias[N]).
//
// Replaces every nn.Linear on the Seeker hot path of the reference
// (third_party/TimeSformer/timesformer/models/vit.py:50-52 Mlp.fc1/fc2, :73-74 Attention.qkv/proj,
// :146 temporal_fc; model/mask_tracker.py:83-86 head linears; vit.py:233 patch-embed conv as an
// im2col GEMM).  A and W are both K-major (row-major activations, nn.Linear weight layout as is).
//
//   warp 0 (1 lane) : TMA producer   - cp.async.bulk.tensor 2-D loads, SWIZZLE_128B, 128x64 A + BNx64 W / stage
//   warp 1 (1 lane) : MMA issuer     - tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16; fp32 accum in TMEM
//   warp 2          : TMEM allocator - 2 accumulator stages x BN columns
//   warps 4-7(-11)  : epilogue       - tcgen05.ld 32x32b -> +bias (-> exact-erf GELU) -> swizzled smem
//                                      -> TMA store (bf16 / fp32) or TMA reduce-add into the fp32 residual stream;
//                                      every warp runs its own double-buffered staging ring (no CTA barrier)
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), persistent tile loop.
#include <stdlib.h>

#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row

// CL: 1 = stand-alone CTAs; 2 = cluster of two, weight tile TMA-multicast; 3 = CTA pair (cta_group::2 MMA).
template <int BN, int EPI, int CL = 1>
struct GemmCfg {
  static constexpr bool PAIR = (CL == 3);
  static constexpr int CSIZE = (CL == 1) ? 1 : 2;   // CTAs per cluster
  static constexpr bool SCALED = (EPI == TCOW_EPI_F32_ADD_SCALED);  // stochastic depth: per-row scale of the branch
  static constexpr bool RED_ADD = (EPI == TCOW_EPI_F32_ADD || SCALED);
  static constexpr bool OUT_F32 = (EPI == TCOW_EPI_F32_STORE || RED_ADD);
  // training epilogues: GELU_AUX also stores the pre-activation (second TMA store through tmAux); DGELU multiplies
  // the accumulator by gelu'(z), z read from the saved pre-activation
  static constexpr bool GELU_AUX = (EPI == TCOW_EPI_BF16_GELU_AUX);
  static constexpr bool DGELU = (EPI == TCOW_EPI_BF16_DGELU);
  static constexpr bool GELU_LIKE = (EPI == TCOW_EPI_BF16_GELU || GELU_AUX || DGELU);
  // The GELU epilogue is issue-bound (exact-erf on 128x256 values per tile): give it 8 warps, 4 otherwise.
  static constexpr int EPI_WARPS = (GELU_LIKE && BN >= 128) ? 8 : 4;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // pair mode: each CTA holds half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_WARP_BYTES = 32 * 128;  // per-warp staging chunk: 32 rows x 128 bytes
  // 4 epilogue warps: double-buffered staging; 8 warps: single-buffered (two warps per scheduler cover each other's
  // TMA-store drain) so that the mainloop keeps 4 smem stages — with 3 the MMA warp waits on TMA 28 % of the time.
  static constexpr int OUT_BUFS = (EPI_WARPS == 8) ? 1 : 2;
  static constexpr int OUT_BYTES = EPI_WARPS * OUT_BUFS * OUT_WARP_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_MAX = 232448;
  static constexpr int STAGES_FIT = (SMEM_MAX - 1024 - BAR_BYTES - OUT_BYTES) / STAGE_BYTES;
#ifdef GEMM_FORCE_STAGES
  static constexpr int STAGES = GEMM_FORCE_STAGES;
#else
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
#endif
  static constexpr int SMEM = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int CHUNK_COLS = OUT_F32 ? 32 : 64;
  static constexpr int NCHUNK = BN / CHUNK_COLS;             // chunks per tile row
  static constexpr int COL_GROUPS = EPI_WARPS / 4;           // warps sharing a TMEM lane quarter split the columns
  static constexpr int CHUNKS_PER_WARP = NCHUNK / COL_GROUPS;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(NCHUNK % COL_GROUPS == 0, "column split");
};

// Exact-erf GELU (nn.GELU() default, vit.py:46) for two values at once, on the packed-fp32x2 FMA pipe with a
// single MUFU (exp2) per value — the SFU is the scarce unit on sm_100 (16 cycles per warp-wide op):
//   gelu(x) = x*Phi(x) = max(x,0) - |x| * h(|x|),   h(a) = 0.5*erfc(a/sqrt2) = q(a) * exp(-a^2/2),
//   q(a) = 0.5*erfcx(a/sqrt2) ~ degree-10 minimax polynomial on [0, 5.75] (|x| is clamped there; h < 2e-8 beyond).
// Max abs error 2.8e-6, max relative error 3.2e-5 wherever |gelu| > 1e-3 (fit and check: DESIGN.md §4.1) — 60x
// below the bf16 rounding applied to the result.  The polynomial is evaluated in na = -min(|x|, 5.75) (odd
// coefficients sign-flipped) so that the last step is one FMA: na*h + max(x,0).
__device__ __forceinline__ uint64_t gelu_erf2(float x0, float x1) {
  const float n0 = fmaxf(-fabsf(x0), -5.75f), n1 = fmaxf(-fabsf(x1), -5.75f);
  const uint64_t na = f2_pack(n0, n1);
  uint64_t q = f2_pack(1.740049385e-07f, 1.740049385e-07f);
  q = f2_fma(q, na, f2_pack(5.850045000e-06f, 5.850045000e-06f));
  q = f2_fma(q, na, f2_pack(8.705152140e-05f, 8.705152140e-05f));
  q = f2_fma(q, na, f2_pack(7.601087564e-04f, 7.601087564e-04f));
  q = f2_fma(q, na, f2_pack(4.371289164e-03f, 4.371289164e-03f));
  q = f2_fma(q, na, f2_pack(1.773692295e-02f, 1.773692295e-02f));
  q = f2_fma(q, na, f2_pack(5.366283283e-02f, 5.366283283e-02f));
  q = f2_fma(q, na, f2_pack(1.275363415e-01f, 1.275363415e-01f));
  q = f2_fma(q, na, f2_pack(2.482131273e-01f, 2.482131273e-01f));
  q = f2_fma(q, na, f2_pack(3.987075090e-01f, 3.987075090e-01f));
  q = f2_fma(q, na, f2_pack(4.999948144e-01f, 4.999948144e-01f));
  const uint64_t w = f2_mul(na, f2_pack(0.84932180028801904f, 0.84932180028801904f));  // sqrt(log2(e)/2)
  float ww0, ww1;
  f2_unpack(f2_mul(w, w), ww0, ww1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-ww0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-ww1));
  const uint64_t h = f2_mul(q, f2_pack(e0, e1));
  return f2_fma(na, h, f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

// d/dx gelu_erf(x) = Phi(x) + x*phi(x) with the same erfcx polynomial: h = q(|x|) e, e = exp(-x^2/2);
// Phi = 1-h (x >= 0) or h (x < 0); phi = e / sqrt(2 pi).  Backward of nn.GELU() (vit.py:46,56).  Packed fp32x2 (two values
// per issue slot, one MUFU each) for the DGELU epilogue.
__device__ __forceinline__ uint64_t dgelu_erf2(float x0, float x1) {
  const float n0 = fmaxf(-fabsf(x0), -5.75f), n1 = fmaxf(-fabsf(x1), -5.75f);
  const uint64_t na = f2_pack(n0, n1);
  uint64_t q = f2_pack(1.740049385e-07f, 1.740049385e-07f);
  q = f2_fma(q, na, f2_pack(5.850045000e-06f, 5.850045000e-06f));
  q = f2_fma(q, na, f2_pack(8.705152140e-05f, 8.705152140e-05f));
  q = f2_fma(q, na, f2_pack(7.601087564e-04f, 7.601087564e-04f));
  q = f2_fma(q, na, f2_pack(4.371289164e-03f, 4.371289164e-03f));
  q = f2_fma(q, na, f2_pack(1.773692295e-02f, 1.773692295e-02f));
  q = f2_fma(q, na, f2_pack(5.366283283e-02f, 5.366283283e-02f));
  q = f2_fma(q, na, f2_pack(1.275363415e-01f, 1.275363415e-01f));
  q = f2_fma(q, na, f2_pack(2.482131273e-01f, 2.482131273e-01f));
  q = f2_fma(q, na, f2_pack(3.987075090e-01f, 3.987075090e-01f));
  q = f2_fma(q, na, f2_pack(4.999948144e-01f, 4.999948144e-01f));
  const uint64_t w = f2_mul(na, f2_pack(0.84932180028801904f, 0.84932180028801904f));
  float ww0, ww1;
  f2_unpack(f2_mul(w, w), ww0, ww1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-ww0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-ww1));
  const uint64_t e = f2_pack(e0, e1);
  // Phi = 0.5 + sign(x) * (0.5 - h),  h = q e
  const uint64_t half = f2_pack(0.5f, 0.5f);
  const uint64_t t = f2_fma(f2_mul(q, e), f2_pack(-1.f, -1.f), half);
  const uint64_t Phi = f2_fma(f2_pack(copysignf(1.f, x0), copysignf(1.f, x1)), t, half);
  return f2_fma(f2_mul(f2_pack(x0, x1), f2_pack(0.3989422804014327f, 0.3989422804014327f)), e, Phi);
}

// Stochastic-depth form of the residual epilogue (DropPath, vit_utils.py:139-164 applied at vit.py:172,186,216):
//   X[r,:] += row_scale[r] * acc[r,:] + bias_scale[r] * bias + bias2
struct RowScale {
  const float* row_scale;
  const float* bias_scale;
  const float* bias2;
};

// CL = 1: stand-alone CTAs.  CL = 2: clusters of two CTAs working on vertically adjacent 128-row tiles of the same
// n-block; each CTA fetches half of the shared weight tile and TMA-multicasts it to both, which cuts the L2->SM
// operand traffic per CTA from 48 KB to 32 KB per k-block (the mainloop's real limiter at ~14 TB/s of L2 reads).
template <int BN, int EPI, int CL>
__global__ void __launch_bounds__(GemmCfg<BN, EPI, CL>::THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int M, int N, int K,
                    const __grid_constant__ CUtensorMap tmAux, const __nv_bfloat16* __restrict__ aux,
                    int64_t ldaux, const RowScale rsc) {
  using Cfg = GemmCfg<BN, EPI, CL>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool PAIR = Cfg::PAIR;
  constexpr int CS = Cfg::CSIZE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t s_out = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bars = s_out + Cfg::OUT_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = N / BN;
  const int num_kb = K / BK;
  // Work unit = CL vertically adjacent tiles of one n-block; every CTA of a cluster walks the same unit sequence.
  const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
  const int num_munits = (num_m + CS - 1) / CS;
  const int unit0 = blockIdx.x / CS, unit_step = gridDim.x / CS;
  // k-th tile of this CTA (cluster): tiles are dealt round-robin, n fastest (neighbouring CTAs share the A rows in L2).
  auto tile_at = [&](int k, int& m_unit, int& n_blk) -> bool {
    const int lin = unit0 + k * unit_step;
    m_unit = lin / num_n;
    n_blk = lin % num_n;
    return lin < num_munits * num_n;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      // multicast mode: every CTA of the cluster must have drained the stage before anyone refills it;
      // pair mode: one commit from the leader's MMA warp frees the stage in both CTAs
      mbar_init(empty_bar(s), CL == 2 ? 2 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), (PAIR ? 2 : 1) * 32 * Cfg::EPI_WARPS);  // pair: both CTAs' epilogues report to the leader
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast can land on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Roles are warp-uniform: the whole warp runs the loop and waits on the mbarriers; one elected lane issues.
  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    uint32_t it = 0;
    for (int k = 0, m_unit, n_blk; tile_at(k, m_unit, n_blk); ++k) {
      const int m_blk = m_unit * CS + crank;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one()) {
          const uint32_t sa = base + s * Cfg::STAGE_BYTES;
          if (PAIR) {
            // both CTAs' bytes are counted on the leader's barrier; the leader arms it for the pair
            if (crank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
            const uint32_t lead_bar = cluster_map_shared(full_bar(s), 0);
            tma_load_2d_pair(sa, &tmA, kb * BK, m_blk * BM, lead_bar);
            tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, kb * BK, n_blk * BN + crank * (BN / 2), lead_bar);
          } else {
          mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, kb * BK, m_blk * BM, full_bar(s));
          if (CL == 1) {
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * BK, n_blk * BN, full_bar(s));
          } else {  // my half of the weight tile, delivered to both CTAs (the peer sends the other half)
            tma_load_2d_mcast(sa + Cfg::A_BYTES + crank * (Cfg::B_BYTES / 2), &tmB, kb * BK,
                              n_blk * BN + crank * (BN / 2), full_bar(s), static_cast<uint16_t>(3));
          }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && (!PAIR || crank == 0)) {
    // ------------------------------------------------ MMA issuer (pair mode: the leader CTA issues for both SMs)
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN);
    uint32_t it = 0, t = 0;
    for (int k = 0, m_unit, n_blk; tile_at(k, m_unit, n_blk); ++k, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      mbar_wait(tempty_bar(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + s * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in addr>>4 units
            if (PAIR) umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // free the smem stage once these MMAs have read it (in every CTA of the cluster);
          // on the last k-block also hand the accumulator to the epilogue (of both CTAs in pair mode)
          if (CL == 1) {
            umma_commit(empty_bar(s));
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          } else if (CL == 2) {
            umma_commit_mcast(empty_bar(s), static_cast<uint16_t>(3));
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          } else {
            umma_commit_pair(empty_bar(s), static_cast<uint16_t>(3));
            if (kb == num_kb - 1) umma_commit_pair(tfull_bar(acc), static_cast<uint16_t>(3));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 4 + Cfg::EPI_WARPS) {
    // -------------------------------------------------- epilogue: each warp owns 32 accumulator rows (its TMEM
    // lane quarter) x CHUNKS_PER_WARP column chunks and runs its own staging ring + TMA stores (no CTA barrier).
    const int ew = warp & 3;
    const int cg = (warp - 4) >> 2;
    const uint32_t my_out = s_out + (warp - 4) * Cfg::OUT_BUFS * Cfg::OUT_WARP_BYTES;
    const uint32_t srow = lane * 128;
    const uint32_t sw = lane & 7;
    uint32_t t = 0, cc = 0;
    for (int k = 0, m_unit, n_blk; tile_at(k, m_unit, n_blk); ++k, ++t) {
      const int m_blk = m_unit * CS + crank;
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ci = 0; ci < Cfg::CHUNKS_PER_WARP; ++ci, ++cc) {
        const int c = cg * Cfg::CHUNKS_PER_WARP + ci;
        const uint32_t buf = my_out + (cc % Cfg::OUT_BUFS) * Cfg::OUT_WARP_BYTES;
        if (elect_one()) tma_wait_group_read<Cfg::OUT_BUFS - 1>();  // the store that last used `buf` has drained it
        __syncwarp();
        const int col0 = n_blk * BN + c * Cfg::CHUNK_COLS;
        if constexpr (Cfg::OUT_F32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c * 32, v);
          tmem_ld_wait();
          if (ci == Cfg::CHUNKS_PER_WARP - 1) {
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(tempty_bar(acc), 0);  // the pair's MMA issuer lives in the leader CTA
            else mbar_arrive(tempty_bar(acc));
          }
          float row_s = 1.f, bias_s = 1.f;
          if constexpr (Cfg::SCALED) {
            const int grow = m_blk * BM + ew * 32 + lane;
            if (grow < M) {
              row_s = __ldg(rsc.row_scale + grow);
              bias_s = __ldg(rsc.bias_scale + grow);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + j) : make_float4(0, 0, 0, 0);
            float4 o;
            if constexpr (Cfg::SCALED) {
              const float4 b2 = rsc.bias2 ? __ldg(reinterpret_cast<const float4*>(rsc.bias2 + col0) + j) : make_float4(0, 0, 0, 0);
              o.x = fmaf(row_s, __uint_as_float(v[4 * j + 0]), fmaf(bias_s, b.x, b2.x));
              o.y = fmaf(row_s, __uint_as_float(v[4 * j + 1]), fmaf(bias_s, b.y, b2.y));
              o.z = fmaf(row_s, __uint_as_float(v[4 * j + 2]), fmaf(bias_s, b.z, b2.z));
              o.w = fmaf(row_s, __uint_as_float(v[4 * j + 3]), fmaf(bias_s, b.w, b2.w));
            } else {
            o.x = __uint_as_float(v[4 * j + 0]) + b.x;
            o.y = __uint_as_float(v[4 * j + 1]) + b.y;
            o.z = __uint_as_float(v[4 * j + 2]) + b.z;
            o.w = __uint_as_float(v[4 * j + 3]) + b.w;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + srow + ((j ^ sw) << 4)), "f"(o.x),
                         "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
        } else {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(t_row + c * 64, v0);
          tmem_ld_32x32(t_row + c * 64 + 32, v1);
          tmem_ld_wait();
          if (ci == Cfg::CHUNKS_PER_WARP - 1) {
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(tempty_bar(acc), 0);  // the pair's MMA issuer lives in the leader CTA
            else mbar_arrive(tempty_bar(acc));
          }
          const int grow = m_blk * BM + ew * 32 + lane;  // this thread's output row
          const uint4* zrow = nullptr;
          if constexpr (Cfg::DGELU) zrow = reinterpret_cast<const uint4*>(aux + static_cast<int64_t>(grow) * ldaux + col0);
#pragma unroll
          for (int pass = 0; pass < (Cfg::GELU_AUX ? 2 : 1); ++pass) {
            if (pass == 1) {  // the pre-activation store must have drained the staging buffer before gelu(z) reuses it
              if (elect_one()) tma_wait_group_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t* v = (j < 4) ? (v0 + 8 * j) : (v1 + 8 * (j - 4));
              float4 b0 = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + 2 * j) : make_float4(0, 0, 0, 0);
              float4 b1 =
                  bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + 2 * j + 1) : make_float4(0, 0, 0, 0);
              float f[8];
              f[0] = __uint_as_float(v[0]) + b0.x;
              f[1] = __uint_as_float(v[1]) + b0.y;
              f[2] = __uint_as_float(v[2]) + b0.z;
              f[3] = __uint_as_float(v[3]) + b0.w;
              f[4] = __uint_as_float(v[4]) + b1.x;
              f[5] = __uint_as_float(v[5]) + b1.y;
              f[6] = __uint_as_float(v[6]) + b1.z;
              f[7] = __uint_as_float(v[7]) + b1.w;
              if (EPI == TCOW_EPI_BF16_GELU || (Cfg::GELU_AUX && pass == 1)) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) f2_unpack(gelu_erf2(f[e], f[e + 1]), f[e], f[e + 1]);
              }
              if constexpr (Cfg::DGELU) {
                const uint4 zz = grow < M ? __ldg(zrow + j) : make_uint4(0, 0, 0, 0);
                const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint64_t dg = dgelu_erf2(__uint_as_float(zw[e] << 16), __uint_as_float(zw[e] & 0xffff0000u));
                  f2_unpack(f2_mul(f2_pack(f[2 * e], f[2 * e + 1]), dg), f[2 * e], f[2 * e + 1]);
                }
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + srow + ((j ^ sw) << 4)),
                           "r"(pack_bf16(f[0], f[1])), "r"(pack_bf16(f[2], f[3])), "r"(pack_bf16(f[4], f[5])),
                           "r"(pack_bf16(f[6], f[7]))
                           : "memory");
            }
            if (Cfg::GELU_AUX && pass == 0) {  // z = A W^T + b goes out through the second tensor map
              fence_proxy_async_smem();
              __syncwarp();
              if (elect_one()) {
                tma_store_2d(&tmAux, buf, col0, m_blk * BM + ew * 32);
                tma_commit_group();
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if constexpr (Cfg::RED_ADD)
            tma_reduce_add_2d(&tmC, buf, col0, m_blk * BM + ew * 32);
          else
            tma_store_2d(&tmC, buf, col0, m_blk * BM + ew * 32);
          tma_commit_group();
        }
      }
    }
    __syncwarp();
    if (elect_one()) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into it or signal its barriers
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host side
template <int BN, int EPI, int CL>
static int launch_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                       int64_t ldc, int M, int N, int K, cudaStream_t stream,
                       const void* aux = nullptr, int64_t ldaux = 0, const RowScale& rsc = RowScale{}) {
  using Cfg = GemmCfg<BN, EPI, CL>;
  constexpr int CS = Cfg::CSIZE;
  alignas(64) CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = make_tmap_2d(&tmA, false, A, K, M, lda, BK, BM))) return rc;
  if ((rc = make_tmap_2d(&tmB, false, W, K, N, ldw, BK, BN / CS))) return rc;
  if ((rc = make_tmap_2d(&tmC, Cfg::OUT_F32, C, N, M, ldc, Cfg::CHUNK_COLS, 32))) return rc;
  alignas(64) CUtensorMap tmAux = tmC;
  if (Cfg::GELU_AUX && (rc = make_tmap_2d(&tmAux, false, aux, N, M, ldaux, Cfg::CHUNK_COLS, 32))) return rc;
  auto kern = gemm_bf16_tn_kernel<BN, EPI, CL>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const int num_m = (M + BM - 1) / BM;
  const int units = ((num_m + CS - 1) / CS) * (N / BN);
  const int slots = sm_count() / CS;
  const int grid = CS * (units < slots ? units : slots);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, bias, M, N, K, tmAux,
                                     static_cast<const __nv_bfloat16*>(aux), ldaux, rsc);
  if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "gemm_bf16_tn_kernel: launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16_tn_kernel");
}

// CTA organisation for the 256-wide tiles: 3 = CTA pairs (cta_group::2, default once there are enough row tiles),
// 2 = cluster of two with weight multicast, 1 = stand-alone CTAs.  TCOW_GEMM_CLUSTER overrides (experiments).
static int cluster_mode(int M, int epi) {
  static const int forced = [] { const char* e = getenv("TCOW_GEMM_CLUSTER"); return e ? atoi(e) : 0; }();
  if (M <= BM * 2) return 1;
  if (forced >= 1 && forced <= 3) return forced;
  // The GELU epilogue is the longest; coupling two CTAs' epilogues to one accumulator hand-off (pair mode) costs
  // it ~4 % (measured), so it keeps independent CTAs sharing the weight tile by multicast.
  return (epi == TCOW_EPI_BF16_GELU || epi == TCOW_EPI_BF16_GELU_AUX || epi == TCOW_EPI_BF16_DGELU) ? 2 : 3;
}

template <int EPI>
static int dispatch_bn(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                       int64_t ldc, int M, int N, int K, cudaStream_t stream, const void* aux = nullptr,
                       int64_t ldaux = 0, const RowScale& rsc = RowScale{}) {
  if (N % 256 == 0) {
    switch (cluster_mode(M, EPI)) {
      case 3: return launch_gemm<256, EPI, 3>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, aux, ldaux, rsc);
      case 2: return launch_gemm<256, EPI, 2>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, aux, ldaux, rsc);
      default: return launch_gemm<256, EPI, 1>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, aux, ldaux, rsc);
    }
  }
  if (N % 128 == 0) return launch_gemm<128, EPI, 1>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, aux, ldaux, rsc);
  return launch_gemm<64, EPI, 1>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, aux, ldaux, rsc);
}

}  // namespace tcow

extern "C" int tcow_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                              int64_t ldc, int M, int N, int K, int epilogue, void* stream) {
  using namespace tcow;
  if (!A || !W || !C) return set_error(TCOW_ERR_ARG, "gemm: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return set_error(TCOW_ERR_ARG, "gemm: non-positive dimension");
  if (N % 64 != 0 || K % 64 != 0)
    return set_error(TCOW_ERR_ARG, "gemm: N (%d) and K (%d) must be multiples of 64", N, K);
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return set_error(TCOW_ERR_ARG, "gemm: bias must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case TCOW_EPI_BF16: return dispatch_bn<TCOW_EPI_BF16>(A, lda, W, ldw, bias, C, ldc, M, N, K, s);
    case TCOW_EPI_BF16_GELU: return dispatch_bn<TCOW_EPI_BF16_GELU>(A, lda, W, ldw, bias, C, ldc, M, N, K, s);
    case TCOW_EPI_F32_STORE: return dispatch_bn<TCOW_EPI_F32_STORE>(A, lda, W, ldw, bias, C, ldc, M, N, K, s);
    case TCOW_EPI_F32_ADD: return dispatch_bn<TCOW_EPI_F32_ADD>(A, lda, W, ldw, bias, C, ldc, M, N, K, s);
  }
  return set_error(TCOW_ERR_ARG, "gemm: unknown epilogue %d", epilogue);
}

extern "C" int tcow_gemm_bf16_aux(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                                  int64_t ldc, void* aux, int64_t ldaux, int M, int N, int K, int epilogue,
                                  void* stream) {
  using namespace tcow;
  if (!A || !W || !C || !aux) return set_error(TCOW_ERR_ARG, "gemm_aux: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return set_error(TCOW_ERR_ARG, "gemm_aux: non-positive dimension");
  if (N % 64 != 0 || K % 64 != 0)
    return set_error(TCOW_ERR_ARG, "gemm_aux: N (%d) and K (%d) must be multiples of 64", N, K);
  if ((ldaux % 8) || (reinterpret_cast<uintptr_t>(aux) & 15))
    return set_error(TCOW_ERR_ARG, "gemm_aux: aux must be 16-byte aligned with a 16-byte-multiple pitch");
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return set_error(TCOW_ERR_ARG, "gemm_aux: bias must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case TCOW_EPI_BF16_GELU_AUX:
      return dispatch_bn<TCOW_EPI_BF16_GELU_AUX>(A, lda, W, ldw, bias, C, ldc, M, N, K, s, aux, ldaux);
    case TCOW_EPI_BF16_DGELU:
      return dispatch_bn<TCOW_EPI_BF16_DGELU>(A, lda, W, ldw, bias, C, ldc, M, N, K, s, aux, ldaux);
  }
  return set_error(TCOW_ERR_ARG, "gemm_aux: epilogue %d takes no auxiliary tensor", epilogue);
}

extern "C" int tcow_gemm_bf16_add_scaled(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                         const float* bias2, const float* row_scale, const float* bias_scale, float* X,
                                         int64_t ldx, int M, int N, int K, void* stream) {
  using namespace tcow;
  if (!A || !W || !X || !row_scale || !bias_scale) return set_error(TCOW_ERR_ARG, "gemm_add_scaled: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return set_error(TCOW_ERR_ARG, "gemm_add_scaled: non-positive dimension");
  if (N % 64 != 0 || K % 64 != 0)
    return set_error(TCOW_ERR_ARG, "gemm_add_scaled: N (%d) and K (%d) must be multiples of 64", N, K);
  if ((bias && (reinterpret_cast<uintptr_t>(bias) & 15)) || (bias2 && (reinterpret_cast<uintptr_t>(bias2) & 15)))
    return set_error(TCOW_ERR_ARG, "gemm_add_scaled: biases must be 16-byte aligned");
  const RowScale rsc{row_scale, bias_scale, bias2};
  return dispatch_bn<TCOW_EPI_F32_ADD_SCALED>(A, lda, W, ldw, bias, X, ldx, M, N, K, static_cast<cudaStream_t>(stream),
                                              nullptr, 0, rsc);
}
