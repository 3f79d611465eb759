"""Torch-tensor front end of the C ABI: pointer/shape plumbing only (PyTorch owns memory and streams)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_BF16_GELU_AUX, EPI_F32_ADD,  # noqa: F401
                   EPI_F32_STORE)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name}: tensor must live on a CUDA device (tcow_b200 has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise ValueError(f'{name}: innermost dimension must be contiguous')


def gemm(a, w, bias, out, epilogue):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias); a, w bf16; out bf16 or fp32 by epilogue."""
    _chk(a, torch.bfloat16, 'gemm.a'); _chk(w, torch.bfloat16, 'gemm.w')
    _chk(out, torch.float32 if epilogue in (EPI_F32_STORE, EPI_F32_ADD) else torch.bfloat16, 'gemm.out')
    if bias is not None:
        _chk(bias, torch.float32, 'gemm.bias')
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or out.shape[0] != M or out.shape[1] != N:
        raise ValueError(f'gemm: shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} out{tuple(out.shape)}')
    _lib.call('tcow_gemm_bf16', a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _p(bias), out.data_ptr(),
              out.stride(0), M, N, K, epilogue, _stream())
    return out


def layernorm(x, gamma, beta, out, eps=1e-6):
    _chk(x, torch.float32, 'layernorm.x'); _chk(out, torch.bfloat16, 'layernorm.out')
    rows, D = x.shape
    if not (x.is_contiguous() and out.is_contiguous() and out.shape == x.shape):
        raise ValueError('layernorm: x and out must be contiguous and of equal shape')
    _lib.call('tcow_layernorm_bf16', x.data_ptr(), _p(gamma), _p(beta), out.data_ptr(), rows, D, float(eps), _stream())
    return out


def attn_temporal(qkv, out, num_seq, T, heads, causal_diag):
    _chk(qkv, torch.bfloat16, 'attn_temporal.qkv'); _chk(out, torch.bfloat16, 'attn_temporal.out')
    _lib.call('tcow_attn_temporal', qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), num_seq, T, heads,
              causal_diag, _stream())
    return out


def attn_spatial(qkv, out, out_cls, B, N, T, heads, use_cls, cls_row0):
    _chk(qkv, torch.bfloat16, 'attn_spatial.qkv'); _chk(out, torch.bfloat16, 'attn_spatial.out')
    if out_cls is not None:
        _chk(out_cls, torch.float32, 'attn_spatial.out_cls')
    _lib.call('tcow_attn_spatial', qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), _p(out_cls), B, N, T,
              heads, int(use_cls), cls_row0, _stream())
    return out


def cls_merge(out_cls, out, B, T, D, cls_row0, mode):
    _lib.call('tcow_cls_merge', out_cls.data_ptr(), out.data_ptr(), out.stride(0), B, T, D, cls_row0, mode, _stream())


_DTYPE_CODE = {torch.float32: 0, torch.uint8: 1}      # TCOW_DTYPE_F32 / TCOW_DTYPE_U8


def patch_gather(frames, query, out, patch, normalize, queries_per_video=1, sample0=0, frame_scale=1.0):
    """frames (V,3,T,H,W), query (B,1,T,H,W), each fp32 or uint8; sample b uses video (sample0 + b) // queries_per_video.
    RGB values are multiplied by frame_scale (1/255 for decoder-style uint8 frames, data/data_plugin.py:174)."""
    for t, n in ((frames, 'frames'), (query, 'query')):
        if not t.is_cuda:
            raise RuntimeError(f'patch_gather.{n}: tensor must live on a CUDA device (tcow_b200 has no CPU path)')
        if t.dtype not in _DTYPE_CODE:
            raise TypeError(f'patch_gather.{n}: expected float32 or uint8, got {t.dtype}')
    _chk(out, torch.bfloat16, 'patch_gather.out')
    V, C, T, Hf, Wf = frames.shape
    B = query.shape[0]
    if C != 3 or tuple(query.shape) != (B, 1, T, Hf, Wf) or not (frames.is_contiguous() and query.is_contiguous()):
        raise ValueError('patch_gather: frames (V,3,T,H,W) and query (B,1,T,H,W) must be contiguous')
    if (sample0 + B - 1) // queries_per_video >= V:
        raise ValueError('patch_gather: not enough videos for the requested samples')
    _lib.call('tcow_patch_gather_typed', frames.data_ptr(), _DTYPE_CODE[frames.dtype], query.data_ptr(),
              _DTYPE_CODE[query.dtype], out.data_ptr(), B, T, Hf, Wf, patch, int(normalize), float(frame_scale),
              int(queries_per_video), int(sample0), _stream())
    return out


PATCH_EMBED_FUSED_MAX_T = 1 << 20     # no limit on the frame count (kept for callers that gate on it)


def patch_embed_fused(frames, query, weight, conv_bias, pos_embed, time_embed, cls_token, X, patch, normalize,
                      queries_per_video=1, sample0=0, frame_scale=1.0):
    """X[:M] = conv(cat(frames, query)) + conv_bias + pos_embed[1+n] + time_embed[t]; X[M:M+B] = cls_token + pos_embed[0] —
    gather, GEMM and embeddings in one kernel (patch 16, D % 256 == 0)."""
    for t, n in ((frames, 'frames'), (query, 'query')):
        if not t.is_cuda:
            raise RuntimeError(f'patch_embed_fused.{n}: tensor must live on a CUDA device (tcow_b200 has no CPU path)')
        if t.dtype not in _DTYPE_CODE:
            raise TypeError(f'patch_embed_fused.{n}: expected float32 or uint8, got {t.dtype}')
    _chk(weight, torch.bfloat16, 'patch_embed_fused.weight'); _chk(X, torch.float32, 'patch_embed_fused.X')
    V, C, T, Hf, Wf = frames.shape
    B = query.shape[0]
    D = weight.shape[0]
    N = (Hf // patch) * (Wf // patch)
    if C != 3 or tuple(query.shape) != (B, 1, T, Hf, Wf) or not (frames.is_contiguous() and query.is_contiguous()):
        raise ValueError('patch_embed_fused: frames (V,3,T,H,W) and query (B,1,T,H,W) must be contiguous')
    if (sample0 + B - 1) // queries_per_video >= V:
        raise ValueError('patch_embed_fused: not enough videos for the requested samples')
    if tuple(weight.shape) != (D, 4 * patch * patch) or not weight.is_contiguous() or not X.is_contiguous() \
            or X.shape[0] < B * N * T + B or X.shape[1] != D:
        raise ValueError('patch_embed_fused: weight [D, 4*P*P] and X [>= M+B, D] must be contiguous')
    _lib.call('tcow_patch_embed_fused', frames.data_ptr(), _DTYPE_CODE[frames.dtype], query.data_ptr(),
              _DTYPE_CODE[query.dtype], weight.data_ptr(), conv_bias.data_ptr(), pos_embed.data_ptr(), time_embed.data_ptr(),
              cls_token.data_ptr(), X.data_ptr(), B, T, Hf, Wf, patch, D, int(normalize), float(frame_scale),
              int(queries_per_video), int(sample0), _stream())
    return X


def embed_init(X, conv_bias, pos_embed, time_embed, cls_token, B, N, T, D):
    _lib.call('tcow_embed_init', X.data_ptr(), conv_bias.data_ptr(), pos_embed.data_ptr(), time_embed.data_ptr(),
              cls_token.data_ptr(), B, N, T, D, _stream())
    return X


def mask_upsample(low, out, B, T, Ho, Wo, C, pp, stride, mode):
    _chk(low, torch.float32, 'mask_upsample.low'); _chk(out, torch.float32, 'mask_upsample.out')
    _lib.call('tcow_mask_upsample', low.data_ptr(), low.stride(0), out.data_ptr(), B, T, Ho, Wo, C, pp, stride, mode,
              _stream())
    return out


def flag_mean(low, flags, B, N, T, F, col0):
    _lib.call('tcow_flag_mean', low.data_ptr(), low.stride(0), flags.data_ptr(), B, N, T, F, col0, _stream())
    return flags


# ------------------------------------------------------------------------------------------------ training step
def gemm_aux(a, w, bias, out, aux, epilogue):
    """EPI_BF16_GELU_AUX: aux = a @ w^T + bias, out = gelu(aux);  EPI_BF16_DGELU: out = (a @ w^T) * gelu'(aux)."""
    _chk(a, torch.bfloat16, 'gemm_aux.a'); _chk(w, torch.bfloat16, 'gemm_aux.w')
    _chk(out, torch.bfloat16, 'gemm_aux.out'); _chk(aux, torch.bfloat16, 'gemm_aux.aux')
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or tuple(out.shape) != (M, N) or tuple(aux.shape) != (M, N):
        raise ValueError('gemm_aux: shape mismatch')
    _lib.call('tcow_gemm_bf16_aux', a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _p(bias), out.data_ptr(),
              out.stride(0), aux.data_ptr(), aux.stride(0), M, N, K, epilogue, _stream())
    return out


def gemm_wgrad(dy, x, dw, db=None, sched=None):
    """dw[N1,N2] (fp32) += dy[R,N1]^T @ x[R,N2]  (bf16 operands, contraction over the token rows); with db (fp32 [N1]) also
    db += dy.sum(0), the bias gradient, from the same pass over dy; with sched (int32[2], zero-initialised once) the units are
    handed out by an atomic ticket instead of static striping (robust against SMs shared with a concurrent kernel)."""
    _chk(dy, torch.bfloat16, 'gemm_wgrad.dy'); _chk(x, torch.bfloat16, 'gemm_wgrad.x'); _chk(dw, torch.float32, 'gemm_wgrad.dw')
    R, N1 = dy.shape
    N2 = x.shape[1]
    if x.shape[0] != R or tuple(dw.shape) != (N1, N2):
        raise ValueError(f'gemm_wgrad: shape mismatch dy{tuple(dy.shape)} x{tuple(x.shape)} dw{tuple(dw.shape)}')
    if db is not None:
        _chk(db, torch.float32, 'gemm_wgrad.db')
        if db.numel() != N1 or not db.is_contiguous():
            raise ValueError('gemm_wgrad: db must be a contiguous fp32 vector of N1 entries')
    if sched is not None and (sched.dtype != torch.int32 or sched.numel() < 2 or not sched.is_cuda):
        raise ValueError('gemm_wgrad: sched must be a CUDA int32 tensor of 2 entries')
    _lib.call('tcow_gemm_bf16_wgrad_sched', dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(),
              dw.stride(0), _p(db), _p(sched), R, N1, N2, _stream())
    return dw


def train_workspace_floats(max_cols=4096):
    return int(_lib.load().tcow_train_workspace_floats(int(max_cols)))


def layernorm_train(x, gamma, beta, out, xhat, rstd, eps=1e-6):
    _chk(x, torch.float32, 'layernorm_train.x'); _chk(out, torch.bfloat16, 'layernorm_train.out')
    _chk(xhat, torch.bfloat16, 'layernorm_train.xhat'); _chk(rstd, torch.float32, 'layernorm_train.rstd')
    rows, D = x.shape
    if not (x.is_contiguous() and out.is_contiguous() and xhat.is_contiguous() and out.shape == x.shape
            and xhat.shape == x.shape and rstd.numel() >= rows):
        raise ValueError('layernorm_train: contiguous tensors of equal shape required')
    _lib.call('tcow_layernorm_bf16_train', x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
              xhat.data_ptr(), rstd.data_ptr(), rows, D, float(eps), _stream())
    return out


def layernorm_bwd(dy, xhat, rstd, gamma, G, Gb, dgamma, dbeta, workspace, accumulate=True, next_scale=None, Gs=None):
    """next_scale / Gs: also write Gs = next_scale[:, None] * G (bf16) for the next branch under stochastic depth."""
    _chk(dy, torch.bfloat16, 'layernorm_bwd.dy'); _chk(xhat, torch.bfloat16, 'layernorm_bwd.xhat')
    _chk(G, torch.float32, 'layernorm_bwd.G'); _chk(Gb, torch.bfloat16, 'layernorm_bwd.Gb')
    rows, D = dy.shape
    for t in (dy, xhat, G, Gb):
        if not t.is_contiguous() or tuple(t.shape) != (rows, D):
            raise ValueError('layernorm_bwd: contiguous [rows, D] tensors required')
    if next_scale is not None:
        _chk(next_scale, torch.float32, 'layernorm_bwd.next_scale'); _chk(Gs, torch.bfloat16, 'layernorm_bwd.Gs')
        if next_scale.numel() < rows or not Gs.is_contiguous() or tuple(Gs.shape) != (rows, D):
            raise ValueError('layernorm_bwd: next_scale needs one entry per row and Gs the shape of G')
        _lib.call('tcow_layernorm_bwd_scaled', dy.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                  G.data_ptr(), Gb.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), workspace.data_ptr(), rows, D,
                  int(accumulate), next_scale.data_ptr(), Gs.data_ptr(), _stream())
        return
    _lib.call('tcow_layernorm_bwd', dy.data_ptr(), xhat.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), G.data_ptr(),
              Gb.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), workspace.data_ptr(), rows, D, int(accumulate),
              _stream())


def colsum(x, out, workspace, accumulate=True):
    _chk(x, torch.bfloat16, 'colsum.x'); _chk(out, torch.float32, 'colsum.out')
    rows, N = x.shape
    if out.numel() != N or workspace.numel() < 256 * N:
        raise ValueError('colsum: output / workspace size mismatch')
    _lib.call('tcow_colsum_bf16', x.data_ptr(), x.stride(0), rows, N, out.data_ptr(), workspace.data_ptr(),
              int(accumulate), _stream())
    return out


def embed_bwd(G, dpos, dtime, dcls_pos0, workspace, B, N, T, D, accumulate=True):
    _chk(G, torch.float32, 'embed_bwd.G')
    _lib.call('tcow_embed_bwd', G.data_ptr(), dpos.data_ptr(), dtime.data_ptr(), _p(dcls_pos0), workspace.data_ptr(),
              B, N, T, D, int(accumulate), _stream())


def mask_head_bwd(d_out, d_flags, d_low, B, T, Ho, Wo, C, pp, stride, mode, F, col0):
    _chk(d_out, torch.float32, 'mask_head_bwd.d_out'); _chk(d_low, torch.bfloat16, 'mask_head_bwd.d_low')
    if not d_out.is_contiguous() or (d_flags is not None and not d_flags.is_contiguous()):
        raise ValueError('mask_head_bwd: contiguous gradients required')
    _lib.call('tcow_mask_head_bwd', d_out.data_ptr(), _p(d_flags), d_low.data_ptr(), d_low.stride(0), B, T, Ho, Wo, C,
              pp, stride, mode, F, col0, d_low.shape[1], _stream())
    return d_low


def attn_spatial_train(qkv, out, out_cls, lse, B, N, T, heads, use_cls, cls_row0):
    _chk(qkv, torch.bfloat16, 'attn_spatial_train.qkv'); _chk(out, torch.bfloat16, 'attn_spatial_train.out')
    _chk(lse, torch.float32, 'attn_spatial_train.lse')
    if lse.numel() < B * T * heads * 304:
        raise ValueError('attn_spatial_train: lse must hold B*T*heads*304 floats')
    _lib.call('tcow_attn_spatial_train', qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), _p(out_cls),
              lse.data_ptr(), B, N, T, heads, int(use_cls), cls_row0, _stream())
    return out


def attn_temporal_bwd(qkv, out, d_out, d_qkv, num_seq, T, heads, causal_diag):
    for t, n in ((qkv, 'qkv'), (out, 'out'), (d_out, 'd_out'), (d_qkv, 'd_qkv')):
        _chk(t, torch.bfloat16, 'attn_temporal_bwd.' + n)
    _lib.call('tcow_attn_temporal_bwd', qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), d_out.data_ptr(),
              d_out.stride(0), d_qkv.data_ptr(), d_qkv.stride(0), num_seq, T, heads, causal_diag, _stream())
    return d_qkv


def spatial_bwd_scratch_floats(B, T, heads):
    """fp32 scratch of tcow_attn_spatial_bwd: per-frame cls gradients [B,T,3,heads*64] + dO.O per token [B*T*heads,304]."""
    return B * T * heads * (3 * 64 + 304)


def attn_spatial_bwd(qkv, out, out_cls, d_out, d_out_cls, lse, d_qkv, d_cls, B, N, T, heads, use_cls, cls_row0):
    for t, n in ((qkv, 'qkv'), (out, 'out'), (d_out, 'd_out'), (d_qkv, 'd_qkv')):
        _chk(t, torch.bfloat16, 'attn_spatial_bwd.' + n)
    if d_cls is None or d_cls.numel() < spatial_bwd_scratch_floats(B, T, heads):
        raise ValueError('attn_spatial_bwd: scratch must hold spatial_bwd_scratch_floats(B, T, heads) floats')
    _lib.call('tcow_attn_spatial_bwd', qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), _p(out_cls),
              d_out.data_ptr(), d_out.stride(0), _p(d_out_cls), lse.data_ptr(), d_qkv.data_ptr(), d_qkv.stride(0),
              _p(d_cls), B, N, T, heads, int(use_cls), cls_row0, _stream())
    return d_qkv


def cls_merge_bwd(d_out, d_out_cls, B, T, D, cls_row0, mode):
    _lib.call('tcow_cls_merge_bwd', d_out.data_ptr(), d_out.stride(0), d_out_cls.data_ptr(), B, T, D, cls_row0, mode,
              _stream())


def gemm_add_scaled(a, w, bias, bias2, row_scale, bias_scale, x):
    """x[M,N] (fp32) += row_scale[:,None] * (a @ w^T) + bias_scale[:,None] * bias + bias2   (stochastic depth)."""
    _chk(a, torch.bfloat16, 'gemm_add_scaled.a'); _chk(w, torch.bfloat16, 'gemm_add_scaled.w')
    _chk(x, torch.float32, 'gemm_add_scaled.x'); _chk(row_scale, torch.float32, 'gemm_add_scaled.row_scale')
    _chk(bias_scale, torch.float32, 'gemm_add_scaled.bias_scale')
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or tuple(x.shape) != (M, N) or row_scale.numel() < M or bias_scale.numel() < M:
        raise ValueError('gemm_add_scaled: shape mismatch')
    _lib.call('tcow_gemm_bf16_add_scaled', a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _p(bias), _p(bias2),
              row_scale.data_ptr(), bias_scale.data_ptr(), x.data_ptr(), x.stride(0), M, N, K, _stream())
    return x


def scale_rows(x, scale, out):
    _chk(x, torch.bfloat16, 'scale_rows.x'); _chk(out, torch.bfloat16, 'scale_rows.out')
    rows, N = x.shape
    if tuple(out.shape) != (rows, N) or scale.numel() < rows:
        raise ValueError('scale_rows: shape mismatch')
    _lib.call('tcow_scale_rows_bf16', x.data_ptr(), x.stride(0), scale.data_ptr(), out.data_ptr(), out.stride(0), rows, N,
              _stream())
    return out


def mask_iou_areas(logits, target):
    """(..., Hf, Wf) fp32 logits and targets -> (..., 3) fp32: |gt|, |pred & gt|, |pred | gt| per image
    (pred = logit > 0, gt = target > 0.5; eval/metrics.py:18-41)."""
    _chk(logits, torch.float32, 'mask_iou_areas.logits'); _chk(target, torch.float32, 'mask_iou_areas.target')
    if logits.shape != target.shape or not (logits.is_contiguous() and target.is_contiguous()):
        raise ValueError('mask_iou_areas: logits and target must be contiguous and of equal shape')
    hw = logits.shape[-1] * logits.shape[-2]
    images = logits.numel() // hw
    out = torch.empty(*logits.shape[:-2], 3, device=logits.device, dtype=torch.float32)
    _lib.call('tcow_mask_iou_areas', logits.data_ptr(), target.data_ptr(), out.data_ptr(), images, hw, _stream())
    return out
