// Spatial attention on tcgen05 + TMEM for frames whose keys do NOT fit shared memory (480x640 frames: S = 1201): K and V
// streamed in 128-key blocks (flash attention), ONE persistent CTA per SM holding TWO independent workers, and every score
// read from TMEM exactly once by one softmax thread per query row (the softmax of attn_r1.cuh, shared with the
// resident-K/V kernel of attn_spatial_r1.cu that serves S <= 304).  Reference: vit.py:78-111 as called at vit.py:186 on the
// tokens assembled at vit.py:179-185.
//
// Work unit = one 128-query tile of one (clip, frame, head); unit u goes to worker u mod (2 x CTAs), query tile fastest, so
// the ~10 tiles of a head run on neighbouring workers at about the same time and its K/V blocks come out of L2, not HBM.
// A worker (group g of the CTA) owns 256 TMEM columns (scores [0,128), P [128,192), O [192,256)), a two-stage K/V ring, two
// Q slots and its own barriers — nothing is shared between the two workers but the SM:
//   warp 0 / 3  TMA producer of worker 0 / 1: Q tile of the unit, then K_j, V_j block by block (4-D gather of the strided
//               canonical rows, SWIZZLE_128B; the cls rows are appended by hand)
//   warp 1 / 2  MMA issuer of worker 0 / 1: S_j = Q K_j^T (SS), O (+)= P_j V_j (A = P from TMEM, V MN-major); S_{j+1} is
//               issued as soon as the softmax threads have READ S_j, so it runs under the exponentials of block j
//   warps 4-7 / 8-11  softmax of worker 0 / 1, one thread per query row: eight tcgen05.ld behind one wait, exact running
//               maximum, exponentials from registers, P (bf16) -> TMEM, O rescaled in TMEM only when a row's maximum moved;
//               O / l -> bf16 -> the unit's (dead) Q slot as a swizzled staging tile -> ONE TMA store
// Replaces the two-CTAs-per-SM kernel that read every score twice (round 1); profiles/r02_notes.md has the numbers.
#include <math.h>
#include <stdlib.h>

#include "attn_r1.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

namespace {
constexpr int RS_STAGES = 2;                       // K/V ring stages per worker
constexpr int RS_STAGE_BYTES = 2 * R1_KB * 128;    // K block + V block
constexpr int RS_QSLOTS = 2;                       // per worker
constexpr int RS_GROUP_BYTES = RS_STAGES * RS_STAGE_BYTES + RS_QSLOTS * R1_QTILE_BYTES;   // 96 KB
constexpr int RS_BAR_STRIDE = 128;                 // barrier bytes per worker
constexpr int RS_SMEM = 2 * RS_GROUP_BYTES + 2 * RS_BAR_STRIDE + 64 + 1024;
}  // namespace

template <int NL>   // 16-key chunks in the last key block (1..8)
__global__ void __launch_bounds__(R1_THREADS, 1)
attn_spatial_rs_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmTail,
                       const __grid_constant__ CUtensorMap tmOut, const R1Args a) {
  extern __shared__ uint8_t smem_rs[];
  const uint32_t raw = smem_u32(smem_rs);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars0 = base + 2 * RS_GROUP_BYTES;
  const uint32_t tmem_slot = bars0 + 2 * RS_BAR_STRIDE;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_rs + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // worker of this warp: service warps 0,1 -> 0 and 2,3 -> 1; softmax warps 4-7 -> 0, 8-11 -> 1
  const int grp = warp < 4 ? (warp >> 1) : (warp - 4) >> 2;
  const uint32_t gbase = base + grp * RS_GROUP_BYTES;
  auto k_buf = [&](int st) { return gbase + st * RS_STAGE_BYTES; };
  auto v_buf = [&](int st) { return gbase + st * RS_STAGE_BYTES + R1_KB * 128; };
  auto q_buf = [&](int slot) { return gbase + RS_STAGES * RS_STAGE_BYTES + slot * R1_QTILE_BYTES; };
  const uint32_t bars = bars0 + grp * RS_BAR_STRIDE;
  auto kv_full = [&](int st) { return bars + 8u * st; };
  auto kv_empty = [&](int st) { return bars + 16u + 8u * st; };
  auto q_full = [&](int s) { return bars + 32u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 48u + 8u * s; };
  const uint32_t s_full = bars + 64, s_free = bars + 72, p_full = bars + 80, p_free = bars + 88, o_full = bars + 96;

  const int N = a.N, T = a.T, heads = a.heads;
  const int S = N + a.use_cls;
  const int S16 = (S + 15) & ~15;
  const int nq = (S + 127) >> 7;
  const int nblk = (S16 + R1_KB - 1) / R1_KB;
  const int nk_last = S16 - R1_KB * (nblk - 1);
  const int D = heads * 64;
  const long long units = static_cast<long long>(a.B) * T * heads * nq;
  const long long worker = static_cast<long long>(blockIdx.x) * 2 + grp, workers = static_cast<long long>(gridDim.x) * 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmFull);
    prefetch_tmap(&tmTail);
    prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int w = 0; w < 2; ++w) {
      const uint32_t bb = bars0 + w * RS_BAR_STRIDE;
      for (int s = 0; s < 2; ++s) {
        mbar_init(bb + 8u * s, 1);           // kv_full
        mbar_init(bb + 16u + 8u * s, 1);     // kv_empty
        mbar_init(bb + 32u + 8u * s, 1);     // q_full
        mbar_init(bb + 48u + 8u * s, 1);     // q_empty
      }
      mbar_init(bb + 64, 1);                 // s_full
      mbar_init(bb + 72, 4);                 // s_free: one arrival per softmax warp
      mbar_init(bb + 80, 4);                 // p_full
      mbar_init(bb + 88, 1);                 // p_free
      mbar_init(bb + 96, 1);                 // o_full
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, R1_TMEM_COLS);
    tmem_relinquish();
  }
  // Rows that TMA never writes (past the last key / query) keep whatever the buffer held before; they are masked or unused,
  // but V must be finite (0 * NaN): zero everything once, afterwards the buffers only ever hold real data.
  for (int idx = threadIdx.x; idx < 2 * RS_GROUP_BYTES / 16; idx += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + idx * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t region = tmem_base + grp * R1_REGION;

  if (warp < 4) {
    setmaxnreg_dec<72>();   // frees 128 x 96 registers: exactly the 256 x 48 the softmax warpgroups add
    if (warp == 0 || warp == 3) {
      // ---------------------------------------------------------------- producer of this worker
      uint32_t kv_it = 0, q_it = 0;
      for (long long u = worker; u < units; u += workers, ++q_it) {
        const int qt = static_cast<int>(u % nq), h = static_cast<int>((u / nq) % heads);
        const int t = static_cast<int>((u / (static_cast<long long>(nq) * heads)) % T);
        const int b = static_cast<int>(u / (static_cast<long long>(nq) * heads * T));
        const int slot = q_it % RS_QSLOTS;
        mbar_wait(q_empty(slot), ((q_it / RS_QSLOTS) & 1) ^ 1);
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
          const int st = kv_it % RS_STAGES;
          mbar_wait(kv_empty(st), ((kv_it / RS_STAGES) & 1) ^ 1);
          const int rows = (N - R1_KB * j) < R1_KB ? (N - R1_KB * j) : R1_KB;
          if (a.use_cls && (N >> 7) == j && lane < 16) {  // cls k / v rows -> row N - 128 j of this block
            const int which = lane >> 3, chunk = lane & 7, r = N - R1_KB * j;
            const uint4 v = *reinterpret_cast<const uint4*>(a.qkv + (a.cls_row0 + b) * a.ld_qkv + (1 + which) * D + h * 64 + chunk * 8);
            const uint32_t dst = (which ? v_buf(st) : k_buf(st)) + r * 128 + ((chunk ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            if (rows > 0) {
              mbar_expect_tx(kv_full(st), 2u * static_cast<uint32_t>(rows) * 128u);
              const CUtensorMap* m = rows == R1_KB ? &tmFull : &tmTail;
              tma_load_4d(k_buf(st), m, D + h * 64, t, R1_KB * j, b, kv_full(st));
              tma_load_4d(v_buf(st), m, 2 * D + h * 64, t, R1_KB * j, b, kv_full(st));
            } else {
              mbar_arrive(kv_full(st));
            }
          }
          __syncwarp();
        }
      }
    } else {
      // ---------------------------------------------------------------- MMA issuer of this worker
      const uint32_t idesc_full = umma_idesc_bf16(128, R1_KB);
      const uint32_t idesc_last = umma_idesc_bf16(128, nk_last);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      uint32_t kv_it = 0, q_it = 0, kc = 0;
      for (long long u = worker; u < units; u += workers, ++q_it) {
        const int slot = q_it % RS_QSLOTS;
        mbar_wait(q_full(slot), (q_it / RS_QSLOTS) & 1);
        const uint64_t qd = umma_desc_k_sw128(q_buf(slot));
        auto issue_s = [&](int b, uint32_t it) {   // S_b = Q K_b^T, K_b in ring position `it`
          const int st = it % RS_STAGES;
          mbar_wait(kv_full(st), (it / RS_STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t kd = umma_desc_k_sw128(k_buf(st));
            const uint32_t idesc = b == nblk - 1 ? idesc_last : idesc_full;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(region, qd + 2u * k, kd + 2u * k, idesc, k > 0 ? 1u : 0u);
            umma_commit(s_full);
          }
          __syncwarp();
        };
        if (kc > 0) mbar_wait(s_free, (kc - 1) & 1);   // the last scores of the previous unit have been read
        issue_s(0, kv_it);
        for (int b = 0; b < nblk; ++b, ++kc, ++kv_it) {
          if (b + 1 < nblk) {   // the next scores as soon as S_b has been read (K_{b+1} sits in the other stage)
            mbar_wait(s_free, kc & 1);
            issue_s(b + 1, kv_it + 1);
          }
          const int st = kv_it % RS_STAGES;
          mbar_wait(p_full, kc & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t vd = umma_desc_mn_sw128(v_buf(st), 1024);
            const int nkk = (b == nblk - 1 ? nk_last : R1_KB) >> 4;
            for (int kk = 0; kk < nkk; ++kk)  // 16 keys per MMA: 8 TMEM columns of P, 16 rows (2048 B) of V
              umma_bf16_ts(region + R1_TMEM_O, region + R1_TMEM_P + 8 * kk, vd + 128u * kk, idesc_o, (b | kk) != 0 ? 1u : 0u);
            umma_commit(kv_empty(st));
            umma_commit(p_free);
            if (b == nblk - 1) umma_commit(o_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output: one thread per query row
    setmaxnreg_inc<216>();
    const int gw = (warp - 4) & 3;                 // warp within the group = TMEM lane quarter (warp id % 4)
    const int row = gw * 32 + lane;
    const uint32_t t_lane = region + (static_cast<uint32_t>(gw * 32) << 16);
    const float sc = a.scale_log2;
    uint32_t kc = 0, o_ct = 0;
    for (long long u = worker; u < units; u += workers, ++o_ct) {
      const int qt = static_cast<int>(u % nq);
      const int slot = o_ct % RS_QSLOTS;
      const int tok = 128 * qt + row;
      const bool valid = tok < S;
      const bool live = 128 * qt + gw * 32 < S;   // warp-uniform: a lane quarter wholly past S only keeps the barriers going
      float m_run = -INFINITY, l_run = 0.f;
      for (int blk = 0; blk < nblk - 1; ++blk)
        r1_softmax_block<8, false>(t_lane, s_full, s_free, p_full, p_free, kc, m_run, l_run, live, valid, lane, S, S16, sc);
      r1_softmax_block<NL, true>(t_lane, s_full, s_free, p_full, p_free, kc, m_run, l_run, live, valid, lane, S, S16, sc);
      // ---- O / l -> bf16 -> staging tile (the unit's Q slot: Q is dead once the last Q K^T has completed) -> TMA store
      const uint32_t stage = q_buf(slot);
      const int h = static_cast<int>((u / nq) % heads);   // the divisions run under the wait for the last P V
      const int t = static_cast<int>((u / (static_cast<long long>(nq) * heads)) % T);
      const int b = static_cast<int>(u / (static_cast<long long>(nq) * heads * T));
      mbar_wait(o_full, o_ct & 1);
      tc_fence_after();
      if (live) {
        const float inv = 1.0f / l_run;
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
      group_sync(grp);
      if (gw == 0 && elect_one()) {
        if (128 * qt < N) {              // rows past N are clipped by the tensor map (the cls row is not part of it)
          tma_store_4d(&tmOut, stage, h * 64, t, 128 * qt, b);
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
  if (warp == 2) tmem_dealloc(tmem_base, R1_TMEM_COLS);
}

int launch_spatial_stream(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                          int heads, int use_cls, int64_t cls_row0, cudaStream_t stream) {
  alignas(64) CUtensorMap tmFull, tmTail, tmO;
  const int cols = 3 * heads * 64;
  const int tail = N % 128;
  int rc;
  if ((rc = make_patch_tmap(&tmFull, qkv, ld_qkv, cols, B, N, T, N >= 128 ? 128 : N))) return rc;
  if ((rc = make_patch_tmap(&tmTail, qkv, ld_qkv, cols, B, N, T, tail > 0 ? tail : 1))) return rc;
  if ((rc = make_patch_tmap(&tmO, out, ld_out, heads * 64, B, N, T, N >= 128 ? 128 : N))) return rc;
  const int S16 = (N + use_cls + 15) & ~15;
  const int nl = (S16 - R1_KB * ((S16 + R1_KB - 1) / R1_KB - 1)) >> 4;   // chunks of the last key block (1..8)
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, R1Args);
  static const Kern kerns[8] = {attn_spatial_rs_kernel<1>, attn_spatial_rs_kernel<2>, attn_spatial_rs_kernel<3>,
                                attn_spatial_rs_kernel<4>, attn_spatial_rs_kernel<5>, attn_spatial_rs_kernel<6>,
                                attn_spatial_rs_kernel<7>, attn_spatial_rs_kernel<8>};
  const Kern kern = kerns[nl - 1];
  static bool configured[64][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63][nl - 1]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63][nl - 1] = true;
  }
  R1Args a{static_cast<const __nv_bfloat16*>(qkv), ld_qkv, static_cast<__nv_bfloat16*>(out), ld_out, out_cls,
           B, N, T, heads, use_cls, cls_row0, 0.125f * 1.4426950408889634f, nullptr};
  const int S = N + use_cls;
  const long long units = static_cast<long long>(B) * T * heads * ((S + 127) / 128);
  const int sms = sm_count();
  const long long ctas = (units + 1) / 2;
  const int grid = ctas < sms ? static_cast<int>(ctas) : sms;
  kern<<<grid, R1_THREADS, RS_SMEM, stream>>>(tmFull, tmTail, tmO, a);
  return check_launch("attn_spatial_rs_kernel");
}

}  // namespace tcow
