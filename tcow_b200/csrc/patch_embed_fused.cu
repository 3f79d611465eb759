// Patch embedding as ONE kernel (BASELINE.json north_star, first bullet): the query mask concatenated as the 4th channel
// (model/mask_tracker.py:103-108), the optional RGB normalisation (model/vision_tf.py:81-89), the Conv2d(4, D, 16, 16) of
// vit.py:233-241 as an implicit GEMM on tcgen05, and the embeddings of model/vision_tf.py:99-138
//     X[(b*N+n)*T+t, :] = conv(x4)[b,t,n,:] + conv_bias + pos_embed[1+n] + time_embed[t],   X[M+b, :] = cls_token + pos_embed[0]
// written straight into the fp32 residual stream.  Neither the im2col matrix (147 MB of bf16 per 8 clips) nor a
// pre-initialised stream ever touches HBM (the three-kernel form — tcow_patch_gather, tcow_embed_init, the reduce-add
// GEMM — stays for T > 32 and for the training step, which saves the im2col matrix for the weight gradient).
//
// Persistent CTAs.  A tile = up to 128 consecutive patches n of ONE frame (b,t) x 256 channels: the 16-pixel segments of
// neighbouring patches are contiguous in an image row, so the im2col loads are coalesced runs (token-row-major tiles, 128
// consecutive (n,t) rows, put every lane in a different frame and measured 2.5x slower than the three-kernel form).
// K = 4*16*16 = 1024 in 16 steps of 64 (= 4 patch rows of one channel):
//   warp 0      TMA producer of the weight tile (256 x 64 bf16, SWIZZLE_128B)
//   warp 1      MMA issuer: tcgen05.mma M=128, N=256, K=16, fp32 accumulators in TMEM, two stages (512 columns)
//   warp 2      TMEM allocator;  warp 3 (CTA 0 only) writes the cls rows
//   warps 4-11  im2col producers: thread = (patch of the tile, half of the step's 4 patch rows); 16 pixels (fp32 or uint8)
//               -> scale / normalise -> bf16 -> the patch's 128-byte K-major row of the A tile, 16-byte chunks swizzled as
//               TMA would have written them
//   warps 12-15 epilogue: tcgen05.ld -> + conv_bias + time_embed[t] (one row per tile, via shared memory) + pos_embed[1+n]
//               (32 x 32 boxes TMA-loaded one chunk ahead) -> swizzled staging -> TMA store through a 4-D view of the
//               stream (column, t, n, b), which puts patch n of frame t at canonical row (b*N+n)*T+t
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {
namespace {

constexpr int PE_BM = 128, PE_BN = 256, PE_BK = 64;
constexpr int PE_STAGES = 3;
constexpr int PE_A_BYTES = PE_BM * PE_BK * 2;
constexpr int PE_B_BYTES = PE_BN * PE_BK * 2;
constexpr int PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_OUT_WARP_BYTES = 32 * 128;
constexpr int PE_OUT_BYTES = 4 * 2 * PE_OUT_WARP_BYTES;      // per epilogue warp: two staging chunks
constexpr int PE_POS_BYTES = 4 * 2 * PE_OUT_WARP_BYTES;      // per epilogue warp: two pos_embed chunks
constexpr int PE_GATHER_WARPS = 8;
constexpr int PE_BAR_BYTES = 256;
constexpr int PE_SMEM = PE_STAGES * PE_STAGE_BYTES + PE_OUT_BYTES + PE_POS_BYTES + PE_BAR_BYTES + 1024;
constexpr int PE_THREADS = 512;

struct PeArgs {
  const void* frames;
  const void* query;
  const float* conv_bias;
  const float* pos;
  const float* tim;
  const float* cls;
  float* X;
  int B, T, Hf, Wf, N, Wo, D, M;
  int normalize, qpv, sample0;
  float frame_scale;
};

__device__ __forceinline__ void px16(const float* src, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + i);
    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
  }
}
__device__ __forceinline__ void px16(const uint8_t* src, float (&v)[16]) {
  // byte -> float without the (quarter-rate) I2F unit: 0x4B0000xx is the float 2^23 + xx, exactly
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[4 * i] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7650)) - 8388608.f;
    v[4 * i + 1] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7651)) - 8388608.f;
    v[4 * i + 2] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7652)) - 8388608.f;
    v[4 * i + 3] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7653)) - 8388608.f;
  }
}
// One 16-pixel segment in registers: raw bytes when every input is uint8 (converted at use, so the load stays in flight),
// otherwise already converted floats.
template <bool RAW8>
struct Seg;
template <>
struct Seg<true> {
  uint4 d;
  __device__ __forceinline__ void zero() { d = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void load(const uint8_t* src) { d = __ldg(reinterpret_cast<const uint4*>(src)); }
  __device__ __forceinline__ void get(float (&v)[16]) const {
    // byte -> float without the (quarter-rate) I2F unit: 0x4B0000xx is the float 2^23 + xx, exactly
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[4 * i] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7650)) - 8388608.f;
      v[4 * i + 1] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7651)) - 8388608.f;
      v[4 * i + 2] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7652)) - 8388608.f;
      v[4 * i + 3] = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7653)) - 8388608.f;
    }
  }
};
template <>
struct Seg<false> {
  float f[16];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = 0.f;
  }
  template <typename T>
  __device__ __forceinline__ void load(const T* src) { px16(src, f); }
  __device__ __forceinline__ void get(float (&v)[16]) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = f[i];
  }
};

__device__ __forceinline__ void tma_store_4d_pe(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <typename FT, typename QT>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_fused_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                         const __grid_constant__ CUtensorMap tmPos, const PeArgs a) {
  extern __shared__ uint8_t smem_pe[];
  const uint32_t raw = smem_u32(smem_pe);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t s_out = base + PE_STAGES * PE_STAGE_BYTES;
  const uint32_t s_pos = s_out + PE_OUT_BYTES;
  const uint32_t bars = s_pos + PE_POS_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PE_STAGES + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * PE_STAGES + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * PE_STAGES + 2 + i); };
  auto pos_bar = [&](int w, int i) { return bars + 8u * (2 * PE_STAGES + 4 + 2 * w + i); };   // per epilogue warp, two slots
  const uint32_t tmem_slot = bars + 8u * (2 * PE_STAGES + 12);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_pe + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = a.D / PE_BN;
  const int tiles_per_frame = (a.N + PE_BM - 1) / PE_BM;
  const int num_m = a.B * a.T * tiles_per_frame;        // tile index m -> (frame f = m / tiles_per_frame, first patch 128*(m % ..))
  const int n_blk = blockIdx.x % num_n;                 // fixed channel tile per CTA
  const int m0 = blockIdx.x / num_n, m_step = gridDim.x / num_n;
  const int ncol0 = n_blk * PE_BN;
  constexpr int KSTEPS = 4 * 16 * 16 / PE_BK;           // 16

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmPos);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PE_STAGES; ++s) {
      mbar_init(full_bar(s), 1 + PE_GATHER_WARPS);    // weight tile (expect_tx) + one arrival per im2col warp
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
    }
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < 2; ++i) mbar_init(pos_bar(w, i), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * PE_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------ weight-tile producer
    uint32_t it = 0;
    for (int m = m0; m < num_m; m += m_step)
      for (int ks = 0; ks < KSTEPS; ++ks, ++it) {
        const int s = it % PE_STAGES;
        mbar_wait(empty_bar(s), ((it / PE_STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), PE_B_BYTES);
          tma_load_2d(base + s * PE_STAGE_BYTES + PE_A_BYTES, &tmW, ks * PE_BK, ncol0, full_bar(s));
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(PE_BM, PE_BN);
    uint32_t it = 0, t = 0;
    for (int m = m0; m < num_m; m += m_step, ++t) {
      const int acc = t & 1;
      mbar_wait(tempty_bar(acc), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int ks = 0; ks < KSTEPS; ++ks, ++it) {
        const int s = it % PE_STAGES;
        mbar_wait(full_bar(s), (it / PE_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + s * PE_STAGE_BYTES;
          const uint64_t ad = umma_desc_k_sw128(sa), bd = umma_desc_k_sw128(sa + PE_A_BYTES);
#pragma unroll
          for (int k = 0; k < PE_BK / 16; ++k)
            umma_bf16(tmem_base + acc * PE_BN, ad + 2u * k, bd + 2u * k, idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (ks == KSTEPS - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ cls rows (vision_tf.py:99-118): X[M+b] = cls_token + pos_embed[0]
    if (blockIdx.x == 0)
      for (int i = lane; i < a.B * a.D; i += 32)
        a.X[static_cast<int64_t>(a.M) * a.D + i] = __ldg(a.cls + i % a.D) + __ldg(a.pos + i % a.D);
  } else if (warp >= 4 && warp < 4 + PE_GATHER_WARPS) {
    // ------------------------------------------------ im2col producers: thread = (patch of the tile, two of the step's 4 rows)
    const int gt = threadIdx.x - 4 * 32;
    const int row = gt & 127, sp = gt >> 7;             // sp: patch rows 2*sp, 2*sp+1 of the k-step
    const int64_t plane = static_cast<int64_t>(a.Hf) * a.Wf;
    const FT* frames = static_cast<const FT*>(a.frames);
    const QT* query = static_cast<const QT*>(a.query);
    const uint32_t row_off = row * 128, sw = row & 7;
    uint32_t it = 0;
    for (int m = m0; m < num_m; m += m_step) {
      const int f = m / tiles_per_frame, n = (m % tiles_per_frame) * PE_BM + row;
      const int t = f % a.T, b = f / a.T;
      const bool live = n < a.N;
      const int vid = (a.sample0 + b) / a.qpv;   // queries of one video share its RGB frames (pipeline.py:134-158)
      const int64_t pix = static_cast<int64_t>((n / a.Wo) * 16) * a.Wf + (n % a.Wo) * 16;
      const int64_t f_base = (static_cast<int64_t>(vid) * 3 * a.T + t) * plane + pix;   // channel c adds c*T*plane
      const int64_t q_base = (static_cast<int64_t>(b) * a.T + t) * plane + pix;
      // software pipeline: the pixels of the next DEPTH steps are in flight while a step is converted and stored (uint8
      // inputs are 16 bytes per segment: four steps ahead keep as many bytes in flight as one step of fp32)
      constexpr bool ALL8 = sizeof(FT) == 1 && sizeof(QT) == 1;
      constexpr int DEPTH = ALL8 ? 4 : 1;
      Seg<ALL8> ring[DEPTH][2];
      auto fetch = [&](int ks, Seg<ALL8> (&r)[2]) {
        const int c = ks >> 2, pr0 = (ks & 3) * 4 + 2 * sp;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          if (!live) r[pr].zero();
          else if (c < 3) r[pr].load(frames + f_base + static_cast<int64_t>(c) * a.T * plane + static_cast<int64_t>(pr0 + pr) * a.Wf);
          else r[pr].load(query + q_base + static_cast<int64_t>(pr0 + pr) * a.Wf);
        }
      };
#pragma unroll
      for (int d0 = 0; d0 < DEPTH; ++d0) fetch(d0, ring[d0]);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks, ++it) {
        const int s = it % PE_STAGES;
        const int c = ks >> 2;
        Seg<ALL8> cur[2] = {ring[ks % DEPTH][0], ring[ks % DEPTH][1]};
        if (ks + DEPTH < KSTEPS) fetch(ks + DEPTH, ring[ks % DEPTH]);
        mbar_wait(empty_bar(s), ((it / PE_STAGES) & 1) ^ 1);
        const uint32_t dst = base + s * PE_STAGE_BYTES + row_off;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          float v[16];
          cur[pr].get(v);
          if (live && c < 3) {
            if (a.frame_scale != 1.0f) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= a.frame_scale;
            }
            if (a.normalize) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = (v[i] - 0.45f) / 0.225f;   // vision_tf.py:23-24
            }
          }
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {   // K index within the step = (2*sp + pr)*16 + w: 16-byte chunk 2*(2*sp + pr) + hc
            const uint32_t chunk = static_cast<uint32_t>(2 * (2 * sp + pr) + hc);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + ((chunk ^ sw) << 4)),
                         "r"(pack_bf16(v[8 * hc], v[8 * hc + 1])), "r"(pack_bf16(v[8 * hc + 2], v[8 * hc + 3])),
                         "r"(pack_bf16(v[8 * hc + 4], v[8 * hc + 5])), "r"(pack_bf16(v[8 * hc + 6], v[8 * hc + 7]))
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(s));
      }
    }
  } else if (warp >= 4 + PE_GATHER_WARPS) {
    // ------------------------------------------------ epilogue: + embeddings -> fp32 stream
    const int ew = warp & 3;                              // TMEM lane quarter = warp id % 4
    const uint32_t my_out = s_out + ew * 2 * PE_OUT_WARP_BYTES;
    const uint32_t my_pos = s_pos + ew * 2 * PE_OUT_WARP_BYTES;
    const uint32_t srow = lane * 128, sw = lane & 7;
    constexpr int NCH = PE_BN / 32;
    uint32_t t_ct = 0, cc = 0;
    // pos_embed chunk (32 patches x 32 channels) for chunk index q of this warp's stream, one chunk ahead
    auto load_pos = [&](int m, int c, uint32_t q) {
      if (elect_one()) {
        const int n0 = (m % tiles_per_frame) * PE_BM + ew * 32;
        mbar_expect_tx(pos_bar(ew, q & 1), PE_OUT_WARP_BYTES);
        tma_load_2d(my_pos + (q & 1) * PE_OUT_WARP_BYTES, &tmPos, ncol0 + c * 32, 1 + n0, pos_bar(ew, q & 1));
      }
      __syncwarp();
    };
    if (m0 < num_m) load_pos(m0, 0, 0);
    for (int m = m0; m < num_m; m += m_step, ++t_ct) {
      const int acc = t_ct & 1;
      const int f = m / tiles_per_frame, n0 = (m % tiles_per_frame) * PE_BM + ew * 32;
      const int t = f % a.T, b = f / a.T;
      const float* time_row = a.tim + static_cast<int64_t>(t) * a.D + ncol0;
      mbar_wait(tfull_bar(acc), (t_ct >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * PE_BN;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++cc) {
        const uint32_t buf = my_out + (cc & 1) * PE_OUT_WARP_BYTES;
        const uint32_t pbuf = my_pos + (cc & 1) * PE_OUT_WARP_BYTES;
        // next chunk's pos_embed box (this tile's next columns, or the next tile's first): its slot was read two chunks ago
        if (c + 1 < NCH) load_pos(m, c + 1, cc + 1);
        else if (m + m_step < num_m) load_pos(m + m_step, 0, cc + 1);
        if (elect_one()) tma_wait_group_read<1>();
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (c == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        mbar_wait(pos_bar(ew, cc & 1), (cc >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float p0, p1, p2, p3;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(p0), "=f"(p1), "=f"(p2), "=f"(p3) : "r"(pbuf + srow + ((j ^ sw) << 4)));
          const float4 e = __ldg(reinterpret_cast<const float4*>(time_row + c * 32) + j);      // same address in every lane
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.conv_bias + ncol0 + c * 32) + j);
          const float o0 = __uint_as_float(v[4 * j]) + bb.x + e.x + p0;
          const float o1 = __uint_as_float(v[4 * j + 1]) + bb.y + e.y + p1;
          const float o2 = __uint_as_float(v[4 * j + 2]) + bb.z + e.z + p2;
          const float o3 = __uint_as_float(v[4 * j + 3]) + bb.w + e.w + p3;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + srow + ((j ^ sw) << 4)), "r"(__float_as_uint(o0)),
                       "r"(__float_as_uint(o1)), "r"(__float_as_uint(o2)), "r"(__float_as_uint(o3))
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (n0 < a.N) tma_store_4d_pe(&tmX, buf, ncol0 + c * 32, t, n0, b);   // patches past N are clipped by the map
          tma_commit_group();
        }
      }
    }
    __syncwarp();
    if (elect_one()) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * PE_BN);
}

template <typename FT, typename QT>
int launch_pe(const CUtensorMap& tmW, const CUtensorMap& tmX, const CUtensorMap& tmPos, const PeArgs& a, int grid, cudaStream_t s) {
  auto kern = patch_embed_fused_kernel<FT, QT>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  kern<<<grid, PE_THREADS, PE_SMEM, s>>>(tmW, tmX, tmPos, a);
  return check_launch("patch_embed_fused_kernel");
}

}  // namespace
}  // namespace tcow

extern "C" int tcow_patch_embed_fused(const void* frames, int frames_dtype, const void* query, int query_dtype,
                                      const void* weight, const float* conv_bias, const float* pos_embed,
                                      const float* time_embed, const float* cls_token, float* X, int B, int T, int Hf, int Wf,
                                      int patch, int D, int normalize, float frame_scale, int queries_per_video, int sample0,
                                      void* stream) {
  using namespace tcow;
  if (!frames || !query || !weight || !conv_bias || !pos_embed || !time_embed || !cls_token || !X || B <= 0 || T <= 0 ||
      queries_per_video < 1 || sample0 < 0)
    return set_error(TCOW_ERR_ARG, "patch_embed_fused: bad argument");
  if ((frames_dtype != TCOW_DTYPE_F32 && frames_dtype != TCOW_DTYPE_U8) || (query_dtype != TCOW_DTYPE_F32 && query_dtype != TCOW_DTYPE_U8))
    return set_error(TCOW_ERR_ARG, "patch_embed_fused: dtype must be TCOW_DTYPE_F32 or TCOW_DTYPE_U8");
  if (patch != 16 || Hf % 16 || Wf % 16 || D % PE_BN)
    return set_error(TCOW_ERR_ARG, "patch_embed_fused: needs patch 16, frame %% 16 == 0, D %% 256 == 0 (use the "
                                   "tcow_patch_gather + tcow_embed_init + GEMM form otherwise)");
  if ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(pos_embed)) & 15)
    return set_error(TCOW_ERR_ARG, "patch_embed_fused: inputs must be 16-byte aligned");
  const int N = (Hf / 16) * (Wf / 16);
  const int64_t M64 = static_cast<int64_t>(B) * N * T;
  if (M64 > 0x7fffffffLL) return set_error(TCOW_ERR_ARG, "patch_embed_fused: too many tokens");
  const int M = static_cast<int>(M64);
  alignas(64) CUtensorMap tmW, tmX, tmPos;
  int rc;
  if ((rc = make_tmap_2d(&tmW, false, weight, 4 * 16 * 16, D, 4 * 16 * 16, PE_BK, PE_BN))) return rc;
  if ((rc = make_tmap_2d(&tmPos, true, pos_embed, D, N + 1, D, 32, 32))) return rc;
  {  // the stream's patch rows as (column, t, n, b): patch n of frame t of clip b sits at row (b*N+n)*T+t
    const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(T), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(D) * 4, static_cast<uint64_t>(T) * D * 4, static_cast<uint64_t>(N) * T * D * 4};
    const uint32_t box[4] = {32, 1, 32, 1};
    if ((rc = make_tmap_nd(&tmX, true, X, 4, dims, strides, box))) return rc;
  }
  const PeArgs a{frames, query, conv_bias, pos_embed, time_embed, cls_token, X, B, T, Hf, Wf, N, Wf / 16, D, M,
                 normalize, queries_per_video, sample0, frame_scale};
  const int num_n = D / PE_BN, num_m = B * T * ((N + PE_BM - 1) / PE_BM);
  int per_n = sm_count() / num_n;
  if (per_n > num_m) per_n = num_m;
  const int grid = per_n * num_n;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (frames_dtype == TCOW_DTYPE_F32 && query_dtype == TCOW_DTYPE_F32) return launch_pe<float, float>(tmW, tmX, tmPos, a, grid, s);
  if (frames_dtype == TCOW_DTYPE_F32) return launch_pe<float, uint8_t>(tmW, tmX, tmPos, a, grid, s);
  if (query_dtype == TCOW_DTYPE_F32) return launch_pe<uint8_t, float>(tmW, tmX, tmPos, a, grid, s);
  return launch_pe<uint8_t, uint8_t>(tmW, tmX, tmPos, a, grid, s);
}
