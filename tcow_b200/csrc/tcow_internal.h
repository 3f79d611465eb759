// Internal helpers shared by the .cu translation units (error reporting, device queries).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tcow_b200.h"

namespace tcow {
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
int make_tmap_nd(void* map, bool is_f32, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box);
int make_tmap_2d(void* map, bool is_f32, const void* ptr, uint64_t inner, uint64_t rows, uint64_t pitch_elems,
                 uint32_t box_inner, uint32_t box_rows);
// tcgen05/TMEM spatial attention with K/V resident in shared memory, one CTA per SM with two query tiles in flight, one
// softmax thread per query row with the scores of a 128-key block held in registers (attn_spatial_r1.cu); valid for
// N + use_cls <= 304.  lse != nullptr also writes the per-row log-sum-exp (training).
int launch_spatial_r1(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                      int heads, int use_cls, int64_t cls_row0, cudaStream_t stream, float* lse = nullptr);
// 4-D TMA view of the patch rows of a token-row matrix: (column, t, n, b) -> ((b*N+n)*T+t)*ld + column; box = 64 columns
// x box_n tokens of one frame (rows T apart), SWIZZLE_128B.
int make_patch_tmap(void* map, const void* base, int64_t ld, int cols, int B, int N, int T, int box_n);
// tcgen05/TMEM backward of the spatial attention (attn_spatial_bwd_tc.cu); valid for N + use_cls <= 304.
int launch_spatial_bwd_tc(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* out_cls,
                          const void* d_out, int64_t ld_do, const float* d_out_cls, const float* lse, void* d_qkv,
                          int64_t ld_dqkv, float* d_cls, float* dsum, int B, int N, int T, int heads, int use_cls,
                          int64_t cls_row0, cudaStream_t stream);
// tcgen05/TMEM spatial attention with K/V streamed in 128-key blocks (flash attention): any N.
int launch_spatial_stream(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N, int T,
                          int heads, int use_cls, int64_t cls_row0, cudaStream_t stream);
}  // namespace tcow
