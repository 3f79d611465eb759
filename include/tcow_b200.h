/* tcow_b200 — C ABI of the B200 (sm_100a) Seeker hot path (forward, and the training step's backward).
 *
 * The reference (basilevh/tcow) is pure Python/PyTorch and has no FFI of its own; every entry point
 * below replaces the torch.nn call(s) cited beside it (paths relative to the reference tree).  A
 * maintainer binds them from Python with ctypes; `tcow_b200/_lib.py` is that
 * binding.
 *
 * Conventions
 *  - plain C types only; all pointers are DEVICE pointers owned by the caller (torch allocations);
 *  - no allocation, no ownership transfer, no implicit synchronisation: work is enqueued on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - every function returns 0 on success or a negative TCOW_ERR_* code; `tcow_last_error()` returns a
 *    thread-local message for the last failure on the calling thread;
 *  - re-entrant: may be called concurrently from one thread per GPU (nn.DataParallel, train.py:223);
 *    the device is the caller's current CUDA device;
 *  - bf16 tensors are raw 16-bit bfloat16, fp32 tensors are IEEE float; "ld*" are row pitches in
 *    ELEMENTS.
 *
 * Canonical activation layout ("token rows"): row r = (b*N + n)*T + t for clip b, patch n = ph*Wo+pw,
 * frame t — the reference's token order x[b, 1 + n*T + t] (model/vision_tf.py:124,137; vit.py:170)
 * without the cls token; the B cls rows are stored after the M = B*N*T patch rows (row M + b).
 */
#ifndef TCOW_B200_H_
#define TCOW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCOW_ERR_ARG (-1)   /* invalid argument (the reference raises AssertionError / ValueError) */
#define TCOW_ERR_CUDA (-2)  /* CUDA runtime / driver failure, including launch errors */
#define TCOW_ERR_ARCH (-3)  /* device is not compute capability 10.x — there is no fallback path */

/* GEMM epilogues */
#define TCOW_EPI_BF16 0      /* C(bf16)  = A W^T + bias                       (qkv: vit.py:81)          */
#define TCOW_EPI_BF16_GELU 1 /* C(bf16)  = gelu_erf(A W^T + bias)             (Mlp.fc1+act: vit.py:55-56) */
#define TCOW_EPI_F32_STORE 2 /* C(fp32)  = A W^T + bias                       (head: mask_tracker.py:113) */
#define TCOW_EPI_F32_ADD 3   /* C(fp32) += A W^T + bias   (proj/temporal_fc/fc2 + residual: vit.py:176,215-216) */
#define TCOW_EPI_BF16_GELU_AUX 5 /* training fc1: aux(bf16) = z = A W^T + bias and C(bf16) = gelu_erf(z) (vit.py:55-56) */
#define TCOW_EPI_F32_ADD_SCALED 7 /* internal: the stochastic-depth residual epilogue of tcow_gemm_bf16_add_scaled     */
#define TCOW_EPI_BF16_DGELU 6    /* backward of the above: C(bf16) = (A W^T) * gelu_erf'(aux)                      */

/* Library / device checks. */
int tcow_abi_version(void);
const char* tcow_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else TCOW_ERR_ARCH. */
int tcow_check_device(void);

/* C[M,N] = epilogue(A[M,K] (bf16) x W[N,K]^T (bf16, nn.Linear layout) + bias[N] (fp32 or NULL)).
 * tcgen05/TMEM tensor-core GEMM fed by TMA.  N and K multiples of 64; M arbitrary.
 * Replaces nn.Linear at vit.py:50-52,73-74,146, mask_tracker.py:83-86 and the Conv2d at vit.py:233. */
int tcow_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                   int64_t ldc, int M, int N, int K, int epilogue, void* stream);

/* y[rows,D] (bf16) = LayerNorm(x[rows,D] (fp32); gamma, beta, eps) with fp32 statistics
 * (nn.LayerNorm(eps=1e-6): vit.py:135,142,150,428).  gamma == NULL: plain fp32 -> bf16 cast. */
int tcow_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int rows, int D,
                        float eps, void* stream);

/* Temporal attention (Attention.forward with the BVH causal mask, vit.py:78-123, called at :172):
 * for each of `num_seq` = B*N sequences of T consecutive token rows and each head,
 *   out = softmax(q k^T * hd^-0.5 + mask) v,   mask: key j allowed for query i iff j <= i + causal_diag
 * (causal_diag < 0 disables the mask; tril(0) -> 0, tril(d) -> d).  qkv is [rows, 3*heads*64] bf16 with
 * column order [q | k | v], head-major (vit.py:81-82); out is [rows, heads*64] bf16.  T <= 64, hd = 64. */
int tcow_attn_temporal(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int num_seq, int T,
                       int heads, int causal_diag, void* stream);

/* Spatial attention (vit.py:186 on the tokens assembled at :179-185): for each clip b, frame t, head:
 * full softmax attention over [cls(b) ; patches n=0..N-1 of frame t] (use_cls=1) or the patches only
 * (use_cls=0, causal_attention >= 2 or -1, vit.py:202-208).  Patch rows are read/written in place in the
 * canonical layout (row (b*N+n)*T+t — no transposes); the cls q/k/v come from row cls_row0 + b of qkv.
 * The cls query's output for every (b,t) goes to out_cls [B,T,heads*64] fp32; the frame-0 value is also written
 * (bf16) to row cls_row0 + b of `out`, which is all causal_attention==1 needs (vit.py:198). */
int tcow_attn_spatial(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, int B, int N,
                      int T, int heads, int use_cls, int64_t cls_row0, void* stream);

/* cls residual input (vit.py:191-198): out[cls_row0+b, :] (bf16) = out_cls[b,0,:] (mode 1, causal_attention==1)
 * or mean_t out_cls[b,t,:] (mode 0, causal_attention==0). */
int tcow_cls_merge(const float* out_cls, void* out, int64_t ld_out, int B, int T, int D, int64_t cls_row0,
                   int mode, void* stream);

/* Patch gather: builds the im2col matrix of the patch-embedding conv with the query mask concatenated as
 * the 4th channel (mask_tracker.py:107-108, vit.py:235-238) in one pass:
 *   P[(b*N+n)*T+t, c*256 + r*16 + w] (bf16) = x4[b,c,t,ph*16+r,pw*16+w],  x4 = cat(frames(3ch), query(1ch))
 * query [B,1,T,Hf,Wf] fp32 (the B samples of this call); frames [V,3,T,Hf,Wf] fp32 where sample b reads video
 * (sample0 + b) / queries_per_video — queries_per_video = 1, sample0 = 0 is the plain one-clip-per-sample case,
 * > 1 lets the Qs queries of a clip (pipeline.py:134-158) share one copy of its RGB frames.
 * normalize != 0 applies (x-0.45)/0.225 to the RGB channels only (vision_tf.py:81-89). */
int tcow_patch_gather(const float* frames, const float* query, void* P, int B, int T, int Hf, int Wf,
                      int patch, int normalize, int queries_per_video, int sample0, void* stream);

/* The same gather reading the frames and/or the query mask as uint8 (TCOW_DTYPE_U8) — the form a video decoder and the
 * reference's loader hold them in before data/data_plugin.py:174 (`rgb / 255.0`) and :182-185 (uint8 query mask) expand
 * them on the host.  RGB values are multiplied by frame_scale (1/255 reproduces :174; 1.0 is the plain cast of
 * mask_tracker.py:103) before the optional normalisation; a clip then crosses PCIe as 9.2 MB instead of 36.9 MB. */
#define TCOW_DTYPE_F32 0
#define TCOW_DTYPE_U8 1
int tcow_patch_gather_typed(const void* frames, int frames_dtype, const void* query, int query_dtype, void* P, int B,
                            int T, int Hf, int Wf, int patch, int normalize, float frame_scale, int queries_per_video,
                            int sample0, void* stream);

/* The whole patch embedding as ONE kernel (north_star bullet 1): query mask as 4th channel (mask_tracker.py:103-108), RGB
 * scale / normalisation (vision_tf.py:81-89), Conv2d(4, D, 16, 16) as an implicit tcgen05 GEMM (vit.py:233-241; weight
 * [D, 4*16*16] bf16, K = c*256 + r*16 + w) and the embeddings of vision_tf.py:99-138 written into the fp32 stream:
 *   X[(b*N+n)*T+t, :] = conv + conv_bias + pos_embed[1+n] + time_embed[t];   X[M+b, :] = cls_token + pos_embed[0].
 * Arguments as tcow_patch_gather_typed; requires patch 16, D % 256 == 0 (TCOW_ERR_ARG otherwise: callers fall
 * back to tcow_patch_gather_typed + tcow_embed_init + tcow_gemm_bf16). */
int tcow_patch_embed_fused(const void* frames, int frames_dtype, const void* query, int query_dtype, const void* weight,
                           const float* conv_bias, const float* pos_embed, const float* time_embed, const float* cls_token,
                           float* X, int B, int T, int Hf, int Wf, int patch, int D, int normalize, float frame_scale,
                           int queries_per_video, int sample0, void* stream);

/* Input path: a decoder-layout uint8 video [F,H,W,C] -> the clip tensor the reference's loader builds on
 * the host, in one pass on the device: frame t = video frame frame_start + t*frame_stride (data/data_plugin.py:156-157),
 * `/ 255.0` (:174), window (y0,x0,h,w) = the centre crop to the target aspect ratio and optional crop rectangle of
 * data/augs.py:170-196 (flip != 0 mirrors the window horizontally first, :189-190), resized to Hf x Wf exactly as
 * torchvision Resize does (:198-206) and written channel-first [C,T,Hf,Wf] (data_plugin.py:199-201).
 * tcow_clip_from_video_u8: antialiased bilinear (ATen's separable triangle filter, align_corners=False), fp32 out in [0,1].
 * tcow_mask_clip_from_video_u8: legacy 'nearest' (src = floor(dst * in/out)), uint8 out — query / target masks. */
int tcow_clip_from_video_u8(const uint8_t* video, int F, int H, int W, int C, int frame_start, int frame_stride, int T,
                            int y0, int x0, int h, int w, int flip, int Hf, int Wf, float* out, void* stream);
int tcow_mask_clip_from_video_u8(const uint8_t* video, int F, int H, int W, int C, int frame_start, int frame_stride,
                                 int T, int y0, int x0, int h, int w, int flip, int Hf, int Wf, uint8_t* out,
                                 void* stream);

/* Residual-stream initialisation (vision_tf.py:99-138): X[(b*N+n)*T+t,:] = conv_bias + pos_embed[1+n] +
 * time_embed[t];  X[M+b,:] = cls_token + pos_embed[0].  The patch GEMM then accumulates into X. */
int tcow_embed_init(float* X, const float* conv_bias, const float* pos_embed, const float* time_embed,
                    const float* cls_token, int B, int N, int T, int D, void* stream);

/* Mask head tail (mask_tracker.py:114-132): `low` [M, ld_low] fp32 holds, per token, the avg-pooled
 * C x pp x pp patch (column c*pp*pp + i*pp + j, pp = patch/stride; the pool is folded into the head
 * weights); writes logits [B,C,T,Hf,Wf] fp32 = upsample x stride of the assembled (Ho*pp) x (Wo*pp) map.
 * mode 0: bilinear, align_corners=True; mode 1: nearest. */
int tcow_mask_upsample(const float* low, int64_t ld_low, float* out, int B, int T, int Ho, int Wo, int C,
                       int pp, int stride, int mode, void* stream);

/* Flags (mask_tracker.py:135-137): flags[b,t,f] = mean_n low[(b*N+n)*T+t, col0+f]. */
int tcow_flag_mean(const float* low, int64_t ld_low, float* flags, int B, int N, int T, int F, int col0,
                   void* stream);

/* Mask areas behind the IoU metrics (eval/metrics.py:18-41; the reference builds three boolean tensors and sums each):
 * for every one of `images` = B*C*T images of hw = Hf*Wf pixels (logits and target contiguous, same layout),
 * areas[img] = (|target > 0.5|, |logit > 0 & target > 0.5|, |logit > 0 or target > 0.5|) as fp32. */
int tcow_mask_iou_areas(const float* logits, const float* target, float* areas, int images, int hw, void* stream);

/* Mask-loss reductions (loss.py:19-31 tversky_loss, :164-212 my_mask_loss) over n logits / targets / optional
 * per-pixel weights (NULL = 1), n a multiple of 4:
 *   sums[0..4] (double) = sum w*bce_with_logits(x,y), sum p*y, sum p*(1-y), sum (1-p)*y, sum y     (p = sigmoid(x))
 *   grad[i] = coef[0]*w*(p - y) + p*(1-p)*(coef[1]*y + coef[2])   — coef (device, fp32[3]) carries the upstream gradients
 * workspace: tcow_mask_loss_workspace_floats() floats. */
int64_t tcow_mask_loss_workspace_floats(void);
int tcow_mask_loss_sums(const float* logits, const float* target, const float* weights, int64_t n, float* workspace,
                        double* sums, void* stream);
int tcow_mask_loss_grad(const float* logits, const float* target, const float* weights, int64_t n, const float* coef,
                        float* grad, void* stream);

/* The reference's complete mask loss (loss.py:164-225 my_mask_loss) over a (B,Q,T,H,W) slice addressed in place as
 * G = B*Q groups of T frames of hw pixels (element (g,t,p) at base + g*group_stride + t*hw + p; hw % 4 == 0):
 *   frames whose weights are all zero are skipped (which_frames, :176-185); l = BCE-with-logits or sigmoid focal loss
 *   (focal != 0; :50-53); custom = mean(w*l); bootstrap = mean of the top int(topk_frac*N) values of l' (:12-16; exact
 *   radix select on the device, l' = w*l if weighted_aot else l); jaccard = Tversky(alpha,beta,eps) (:19-31) or the
 *   bootstrap term itself when weighted_aot (:205); loss = sqrt(selected/total frames) * (aot*(bootstrap+jaccard)/2 +
 *   (1-aot)*custom), or 0 when nothing is selected / mean(w) < 1e-4 (:187,:221).
 * forward writes the scalar loss (device fp32), the per-frame selection flags [G*T] and the `state` block
 * (tcow_mask_loss_state_bytes() bytes) that backward consumes: grad (=|+=) upstream * dloss/dlogits, same addressing. */
int64_t tcow_mask_loss_state_bytes(void);
int tcow_mask_loss_forward(const float* logits, int64_t gs_x, const float* target, int64_t gs_y, const float* weights,
                           int64_t gs_w, int G, int T, int64_t hw, int focal, int weighted_aot, double aot_loss,
                           double topk_frac, double alpha, double beta, double eps, uint8_t* frame_sel, void* state,
                           float* loss_out, void* stream);
int tcow_mask_loss_backward(const float* logits, int64_t gs_x, const float* target, int64_t gs_y, const float* weights,
                            int64_t gs_w, int G, int T, int64_t hw, int focal, int weighted_aot, const uint8_t* frame_sel,
                            const void* state, const float* upstream, float* grad, int64_t gs_g, int accumulate,
                            void* stream);

/* Per-pixel loss weights (loss.py:83-148 get_mask_track_pixel_weights, times the per-frame weights of :55-81):
 *   counts[0..1] = #(target == 1), #(target == 0)                                   (class balancing statistics, :103-107)
 *   out = frame_w[f] * (corr[0] if target == 1, corr[1] if target == 0) * (2 if occl != 0) * (hard_negative_factor if the
 *         pixel is within band/2 of a target > 0 pixel (reflect padding — what gaussian_blur(...) > 0 amounts to, :133-146)
 *         and target < 0.5).   corr: device fp32[2] = (pos_corr, neg_corr) or NULL; occl, frame_w may be NULL;
 *   tmp: G*T*H*W bytes of scratch (needed when hard_negative_factor > 1); band odd. */
int tcow_loss_class_counts(const float* target, int64_t gs_t, int G, int T, int64_t hw, uint64_t* counts, void* stream);
int tcow_loss_pixel_weights(const float* target, int64_t gs_t, const uint8_t* occl, int64_t gs_o, const float* frame_w,
                            const float* corr, int G, int T, int H, int W, float hard_negative_factor, int band,
                            uint8_t* tmp, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step (BASELINE configs[3]: fwd+bwd, data-parallel).  The reference gets its backward from torch.autograd
 * over the modules cited above (train.py:93-101 loss.backward(); optimizer.step()); these entry points are the
 * hand-written adjoints, driven by tcow_b200/train_engine.py behind a torch.autograd.Function.
 * ------------------------------------------------------------------------------------------------------------ */

/* GEMM with an auxiliary bf16 tensor aux[M,N] (pitch ldaux):
 *  TCOW_EPI_BF16_GELU_AUX: aux = z = A W^T + bias, C = gelu_erf(z)        (Mlp.fc1 + act saving the pre-activation)
 *  TCOW_EPI_BF16_DGELU:    C = (A W^T) * gelu_erf'(aux)                    (dZ = (dY W2) o gelu'(z), bias must be NULL) */
int tcow_gemm_bf16_aux(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                       void* aux, int64_t ldaux, int M, int N, int K, int epilogue, void* stream);

/* Residual GEMM under stochastic depth (DropPath, vit_utils.py:139-164 as applied at vit.py:172,186,216):
 *   X[r,:] (fp32) += row_scale[r] * (A W^T)[r,:] + bias_scale[r] * bias + bias2
 * row_scale / bias_scale: fp32 [M] (0 or 1/keep_prob per dropped / kept sequence); bias2 may be NULL. */
int tcow_gemm_bf16_add_scaled(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                              const float* bias2, const float* row_scale, const float* bias_scale, float* X, int64_t ldx,
                              int M, int N, int K, void* stream);

/* out[r,:] (bf16) = scale[r] * x[r,:] (bf16): the branch gradient under stochastic depth. */
int tcow_scale_rows_bf16(const void* x, int64_t ldx, const float* scale, void* out, int64_t ld_out, int rows, int N,
                         void* stream);

/* Weight gradient of nn.Linear:  dW[N1,N2] (fp32) += A[R,N1]^T (bf16) x B[R,N2] (bf16), contraction over the R
 * token rows (A = dY, B = the layer input).  tcgen05 with both operands MN-major (no transposes in memory), split
 * along R, partial products combined with TMA reduce-add: the caller zeroes dW (or accumulates across micro-batches).
 * N1, N2 multiples of 64. */
int tcow_gemm_bf16_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw, int R, int N1,
                         int N2, void* stream);
/* The same weight gradient plus the bias gradient of the layer in the same pass: db[n1] (fp32, N1 entries) += sum over the
 * R rows of A[:, n1] — the column sums come out of the dY tiles the MMA already has in shared memory (replaces a separate
 * tcow_colsum_bf16 pass).  db == NULL: exactly tcow_gemm_bf16_wgrad. */
int tcow_gemm_bf16_wgrad_bias(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw, float* db,
                              int R, int N1, int N2, void* stream);
/* The same with ticket scheduling: sched = two ints of device memory, zero before the first launch (the kernel leaves them
 * zero again).  The row splits become finer (about three units per SM) and every CTA takes its next unit from an atomic
 * counter, so SMs that are late or shared with a concurrent kernel — NCCL's all-reduce of the previous gradient bucket
 * during the data-parallel backward (train.py:222-223 is where the reference parallelises) — take fewer units instead of
 * stretching the kernel.  sched == NULL: static striping. */
int tcow_gemm_bf16_wgrad_sched(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw, float* db,
                               int* sched, int R, int N1, int N2, void* stream);

/* Floats of scratch the reductions below need (LayerNorm-backward / column-sum / time-embedding partial sums). */
int64_t tcow_train_workspace_floats(int max_cols);

/* LayerNorm forward that also saves what the backward needs: y = xhat*gamma+beta (bf16), xhat (bf16), rstd (fp32). */
int tcow_layernorm_bf16_train(const float* x, const float* gamma, const float* beta, void* y, void* xhat, float* rstd,
                              int rows, int D, float eps, void* stream);

/* LayerNorm backward fused with the residual-gradient update: with g = dy*gamma,
 *   dx = rstd * (g - mean(g) - xhat*mean(g*xhat));  G[rows,D] (fp32) = (accumulate ? G : 0) + dx;  Gb = bf16(G);
 *   dgamma += sum_rows dy*xhat;  dbeta += sum_rows dy   (deterministic two-stage reduction through `workspace`). */
int tcow_layernorm_bwd(const void* dy, const void* xhat, const float* rstd, const float* gamma, float* G, void* Gb,
                       float* dgamma, float* dbeta, float* workspace, int rows, int D, int accumulate, void* stream);

/* Same, and additionally Gs[r,:] (bf16) = next_scale[r] * G[r,:]: under stochastic depth the next branch of the backward
 * consumes the row-scaled residual gradient (saves the separate tcow_scale_rows_bf16 pass). */
int tcow_layernorm_bwd_scaled(const void* dy, const void* xhat, const float* rstd, const float* gamma, float* G, void* Gb,
                              float* dgamma, float* dbeta, float* workspace, int rows, int D, int accumulate,
                              const float* next_scale, void* Gs, void* stream);

/* Bias gradient: out[N] (fp32) = (accumulate ? out : 0) + sum_r x[r, :] for x bf16 [rows, N] (pitch ldx). */
int tcow_colsum_bf16(const void* x, int64_t ldx, int rows, int N, float* out, float* workspace, int accumulate,
                     void* stream);

/* Embedding gradients (adjoint of tcow_embed_init, vision_tf.py:99-138) from the residual-stream gradient G[M+B, D]:
 *   dpos[1+n] (+)= sum_{b,t} G[(b*N+n)*T+t];  dtime[t] (+)= sum_{b,n} G[(b*N+n)*T+t];  dcls_pos0 (+)= sum_b G[M+b]
 * (dcls_pos0 is the gradient of both cls_token and pos_embed[0]; the conv bias gradient is sum_t dtime[t]). */
int tcow_embed_bwd(const float* G, float* dpos, float* dtime, float* dcls_pos0, float* workspace, int B, int N, int T,
                   int D, int accumulate, void* stream);

/* Adjoint of tcow_mask_upsample + tcow_flag_mean: d_out [B,C,T,Hf,Wf] fp32 and d_flags [B,T,F] fp32 (or NULL) ->
 * d_low [M, ld_low] bf16: columns [0,col0) the pooled-patch gradients, [col0,col0+F) d_flags/N, rest zero up to
 * ncols (the dY operand of the head's weight-gradient and input-gradient GEMMs). */
int tcow_mask_head_bwd(const float* d_out, const float* d_flags, void* d_low, int64_t ld_low, int B, int T, int Ho,
                       int Wo, int C, int pp, int stride, int mode, int F, int col0, int ncols, void* stream);

/* Spatial attention forward that also writes the base-2 log-sum-exp of every query token: lse [B*T*heads, 304]. */
int tcow_attn_spatial_train(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* out_cls, float* lse,
                            int B, int N, int T, int heads, int use_cls, int64_t cls_row0, void* stream);

/* Backward of tcow_attn_temporal: d_qkv[rows, 3*heads*64] (bf16) from qkv, the saved output `out` and d_out. */
int tcow_attn_temporal_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const void* d_out,
                           int64_t ld_do, void* d_qkv, int64_t ld_dqkv, int num_seq, int T, int heads, int causal_diag,
                           void* stream);

/* Backward of tcow_attn_spatial(_train).  out / d_out: patch rows (bf16); out_cls / d_out_cls: the cls query's
 * per-frame output and its gradient, [B,T,heads*64] fp32; lse from the forward; d_cls: scratch of
 * B*T*heads*(3*64 + 304) floats (per-frame cls gradients [B,T,3,heads*64], then dO.O per token [B*T*heads, 304]).  Writes d_qkv for every patch row and (summed over the frames) for row cls_row0+b. */
int tcow_attn_spatial_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* out_cls,
                          const void* d_out, int64_t ld_do, const float* d_out_cls, const float* lse, void* d_qkv,
                          int64_t ld_dqkv, float* d_cls, int B, int N, int T, int heads, int use_cls, int64_t cls_row0,
                          void* stream);

/* Adjoint of the cls read-out (tcow_cls_merge / the in-kernel frame-0 write): d_out_cls[b,t,:] (fp32) =
 * d_out[cls_row0+b,:] * (mode 0: 1/T for all t; mode 1: 1 for t == 0, else 0). */
int tcow_cls_merge_bwd(const void* d_out, int64_t ld_out, float* d_out_cls, int B, int T, int D, int64_t cls_row0,
                       int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCOW_B200_H_ */
