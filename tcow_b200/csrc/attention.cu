// Fused attention kernels for the divided space-time block (vit.py:78-123 Attention.forward as called at
// vit.py:172 (temporal, causal) and vit.py:186 (spatial, full)).  Both read q/k/v straight out of the QKV GEMM
// output in the canonical token-row layout — no rearranges (vit.py:170,173,181-185,210 are eliminated) — keep
// the softmax state in registers (fp32, exp2 with pre-scaled logits, warp-shuffle row reductions), apply the
// mask in-kernel and never materialise the score matrix.
//
// The temporal kernel (30x30 problems, far below a tcgen05 tile, bandwidth-shaped) uses warp-level mma.sync m16n8k16 bf16
// with ldmatrix from XOR-swizzled shared memory; the spatial attention is tcgen05/TMEM (attn_spatial_r1.cu for frames of
// up to 304 tokens, attn_spatial_rs.cu streamed for longer ones) and only dispatched from here.
#include <math.h>
#include <stdlib.h>

#include "attn_frag.cuh"
#include "ptx.cuh"
#include "tcow_internal.h"

namespace tcow {

// ============================================================================================ temporal
// One warp per (sequence, head).  T <= T_PAD in {32, 64}; all keys in one pass (no online rescale).
template <int T_PAD>
__global__ void __launch_bounds__(128) attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld_qkv,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                            int num_seq, int T, int heads, int causal_diag,
                                                            float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_att[];
  constexpr int TILE = T_PAD * ROW_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long work = static_cast<long long>(blockIdx.x) * 4 + warp;
  if (work >= static_cast<long long>(num_seq) * heads) return;
  const int seq = static_cast<int>(work / heads), head = static_cast<int>(work % heads);
  const uint32_t sQ = smem_u32(smem_att) + warp * 3 * TILE;
  const uint32_t sK = sQ + TILE, sV = sK + TILE;
  const int D = heads * HD;

  // ---- load q, k, v rows of this (sequence, head): T rows x 128 B each, 16 B per cp.async; lane = (row mod 4, chunk):
  // no divisions in the loops (the kernel is issue-bound: 71 % of the issue slots busy at 57 % of the HBM rate)
  const __nv_bfloat16* src = qkv + static_cast<int64_t>(seq) * T * ld_qkv + head * HD;
  const int lrow = lane >> 3, chunk = lane & 7;
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    const __nv_bfloat16* p = src + which * D + chunk * 8 + static_cast<int64_t>(lrow) * ld_qkv;
    const uint32_t tile = sQ + which * TILE;
#pragma unroll 4
    for (int row = lrow; row < T; row += 4, p += 4 * ld_qkv) cp_async_16(sw_addr(tile, row, chunk), p);
    for (int row = T + lrow; row < T_PAD; row += 4)   // zero the padding rows (V must be finite)
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sw_addr(tile, row, chunk)), "r"(0) : "memory");
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  temporal_attend_seq<T_PAD>(sQ, sK, sV, 0, T, causal_diag, scale_log2);
  __syncwarp();
  __nv_bfloat16* dst = out + static_cast<int64_t>(seq) * T * ld_out + head * HD + chunk * 8 + static_cast<int64_t>(lrow) * ld_out;
#pragma unroll 4
  for (int row = lrow; row < T; row += 4, dst += 4 * ld_out) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw_addr(sQ, row, chunk)));
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

// cls residual input: out[cls_row0+b,:] = out_cls[b,0,:] (mode 1) or mean_t out_cls[b,t,:] (mode 0).
__global__ void cls_merge_kernel(const float* __restrict__ out_cls, __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                 int B, int T, int D, int64_t cls_row0, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx % D;
  float v;
  if (mode == 1) {
    v = out_cls[static_cast<int64_t>(b) * T * D + d];
  } else {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += out_cls[(static_cast<int64_t>(b) * T + t) * D + d];
    v = s / static_cast<float>(T);
  }
  out[(cls_row0 + b) * ld_out + d] = __float2bfloat16_rn(v);
}

template <int T_PAD>
static int launch_temporal(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int num_seq, int T, int heads,
                           int causal_diag, cudaStream_t s) {
  auto kern = attn_temporal_kernel<T_PAD>;
  constexpr int smem = 4 * 3 * T_PAD * ROW_BYTES;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured[dev & 63] = true;
  }
  const long long work = static_cast<long long>(num_seq) * heads;
  const long long blocks = (work + 3) / 4;
  if (blocks > 0x7fffffffLL) return set_error(TCOW_ERR_ARG, "attn_temporal: too many sequences");
  kern<<<static_cast<unsigned>(blocks), 128, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), ld_qkv,
                                                        static_cast<__nv_bfloat16*>(out), ld_out, num_seq, T, heads,
                                                        causal_diag, 0.125f * 1.4426950408889634f);
  return check_launch("attn_temporal_kernel");
}

}  // namespace tcow

extern "C" int tcow_attn_temporal(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int num_seq, int T,
                                  int heads, int causal_diag, void* stream) {
  using namespace tcow;
  if (!qkv || !out || num_seq <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_temporal: bad argument");
  if (T < 1 || T > 64) return set_error(TCOW_ERR_ARG, "attn_temporal: T=%d unsupported (1..64)", T);
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_temporal: row pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T <= 32) return launch_temporal<32>(qkv, ld_qkv, out, ld_out, num_seq, T, heads, causal_diag, s);
  return launch_temporal<64>(qkv, ld_qkv, out, ld_out, num_seq, T, heads, causal_diag, s);
}

extern "C" int tcow_attn_spatial(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B,
                                 int N, int T, int heads, int use_cls, int64_t cls_row0, void* stream) {
  using namespace tcow;
  if (!qkv || !out || B <= 0 || N <= 0 || T <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_spatial: bad argument");
  if (use_cls && !out_cls) return set_error(TCOW_ERR_ARG, "attn_spatial: out_cls required when use_cls");
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_spatial: row pitches must be multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int S = N + (use_cls ? 1 : 0);
  // K/V resident in shared memory whenever one frame's keys fit (S <= 304: attn_spatial_r1.cu), else streamed in 128-key
  // blocks (e.g. 480x640 frames, S = 1201).  TCOW_SPATIAL_IMPL=stream forces the streamed kernel for every S (tests).
  static const char impl = [] { const char* e = getenv("TCOW_SPATIAL_IMPL"); return e ? e[0] : '\0'; }();
  if (S <= 304 && impl != 's')
    return launch_spatial_r1(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
  return launch_spatial_stream(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0, s);
}

extern "C" int tcow_attn_spatial_train(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls,
                                       float* lse, int B, int N, int T, int heads, int use_cls, int64_t cls_row0,
                                       void* stream) {
  using namespace tcow;
  if (!qkv || !out || !lse || B <= 0 || N <= 0 || T <= 0 || heads <= 0) return set_error(TCOW_ERR_ARG, "attn_spatial_train: bad argument");
  if (use_cls && !out_cls) return set_error(TCOW_ERR_ARG, "attn_spatial_train: out_cls required when use_cls");
  if ((ld_qkv % 8) || (ld_out % 8)) return set_error(TCOW_ERR_ARG, "attn_spatial_train: row pitches must be multiples of 8");
  if (N + (use_cls ? 1 : 0) > 304)
    return set_error(TCOW_ERR_ARG, "attn_spatial_train: %d tokens per frame > 304 not supported in training", N + (use_cls ? 1 : 0));
  return launch_spatial_r1(qkv, ld_qkv, out, ld_out, out_cls, B, N, T, heads, use_cls ? 1 : 0, cls_row0,
                           static_cast<cudaStream_t>(stream), lse);
}

extern "C" int tcow_cls_merge(const float* out_cls, void* out, int64_t ld_out, int B, int T, int D, int64_t cls_row0,
                              int mode, void* stream) {
  using namespace tcow;
  if (!out_cls || !out || B <= 0 || T <= 0 || D <= 0) return set_error(TCOW_ERR_ARG, "cls_merge: bad argument");
  const int n = B * D;
  cls_merge_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out_cls, static_cast<__nv_bfloat16*>(out), ld_out, B, T, D, cls_row0, mode);
  return check_launch("cls_merge_kernel");
}
