"""Per-kernel checks against plain fp32 PyTorch on the GPU (TF32 off).  Each check returns
(max_abs_err, tolerance, detail).  Used by tests/test_kernels_gpu.py and tools/gpu_diag.py."""
import math

import torch
import torch.nn.functional as F

from tcow_b200 import ops


def _dev():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device('cuda:0')


def check_gemm(M, N, K, epi, seed=0):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(seed)
    a = (torch.randn(M, K, device=d, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=d, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=d, generator=g) * 0.1
    ref = a.float() @ w.float().t() + bias
    if epi == ops.EPI_BF16_GELU:
        ref = F.gelu(ref)
    if epi in (ops.EPI_BF16, ops.EPI_BF16_GELU):
        out = torch.full((M, N), float('nan'), device=d, dtype=torch.bfloat16)
        ops.gemm(a, w, bias, out, epi)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs()
        tol = 2.0 ** -8 * ref.abs().max().item() + 1e-3
    elif epi == ops.EPI_F32_STORE:
        out = torch.full((M, N), float('nan'), device=d)
        ops.gemm(a, w, bias, out, epi)
        torch.cuda.synchronize()
        err = (out - ref).abs()
        tol = 2e-4 * max(1.0, math.sqrt(K / 768))
    else:
        base = torch.randn(M, N, device=d, generator=g)
        out = base.clone()
        ops.gemm(a, w, bias, out, epi)
        torch.cuda.synchronize()
        err = (out - (base + ref)).abs()
        tol = 2e-4 * max(1.0, math.sqrt(K / 768))
    bad = torch.isnan(err).sum().item()
    e = float('inf') if bad else err.max().item()
    where = ''
    if e > tol:
        idx = torch.nonzero(torch.isnan(err) | (err > tol))
        rows, cols = idx[:, 0], idx[:, 1]
        where = (f' bad={idx.shape[0]} rows[{rows.min().item()}..{rows.max().item()}] '
                 f'cols[{cols.min().item()}..{cols.max().item()}] nan={bad}')
    return e, tol, f'gemm M={M} N={N} K={K} epi={epi}{where}'


def check_gelu_accuracy():
    """fc1 epilogue on inputs spanning [-9, 9]: A = identity-like so that the accumulator equals the bias grid."""
    d = _dev()
    M, N, K = 256, 3072, 64
    a = torch.zeros(M, K, device=d, dtype=torch.bfloat16)
    w = torch.zeros(N, K, device=d, dtype=torch.bfloat16)
    bias = torch.linspace(-9, 9, N, device=d)
    out = torch.empty(M, N, device=d, dtype=torch.bfloat16)
    ops.gemm(a, w, bias, out, ops.EPI_BF16_GELU)
    torch.cuda.synchronize()
    ref = F.gelu(bias.double()).float()
    got = out[0].float()
    err = (got - ref.to(torch.bfloat16).float()).abs()
    # allow one bf16 ulp of the reference value (rounding-boundary flips) plus the documented 3e-6 absolute error
    tol = ref.abs() * 2.0 ** -7 + 4e-6
    worst = (err - tol).max().item()
    return max(worst, 0.0), 0.0, f'gelu epilogue accuracy on [-9,9]: max err {err.max().item():.3e}'


def check_layernorm(rows, D=768, affine=True):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(1)
    x = torch.randn(rows, D, device=d, generator=g) * 2 + 0.3
    gm = 1 + 0.1 * torch.randn(D, device=d, generator=g)
    bt = 0.1 * torch.randn(D, device=d, generator=g)
    out = torch.empty(rows, D, device=d, dtype=torch.bfloat16)
    if affine:
        ops.layernorm(x, gm, bt, out, 1e-6)
        ref = F.layer_norm(x, (D,), gm, bt, 1e-6)
    else:
        ops.layernorm(x, None, None, out)
        ref = x
    torch.cuda.synchronize()
    return (out.float() - ref).abs().max().item(), 2.0 ** -8 * ref.abs().max().item() + 1e-3, f'layernorm rows={rows} affine={affine}'


def _mha_ref(q, k, v, mask):
    a = (q @ k.transpose(-1, -2)) * 0.125
    if mask is not None:
        a = a.masked_fill(~mask, float('-inf'))
    return a.softmax(-1) @ v


def check_attn_temporal(num_seq, T, causal_diag, heads=12):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(2)
    D = heads * 64
    extra = 3
    qkv = (torch.randn(num_seq * T + extra, 3 * D, device=d, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.zeros(num_seq * T + extra, D, device=d, dtype=torch.bfloat16)
    ops.attn_temporal(qkv, out, num_seq, T, heads, causal_diag)
    torch.cuda.synchronize()
    x = qkv[:num_seq * T].float().reshape(num_seq, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    mask = None
    if causal_diag >= 0:
        mask = torch.ones(T, T, dtype=torch.bool, device=d).tril(causal_diag)
    ref = _mha_ref(x[0], x[1], x[2], mask).permute(0, 2, 1, 3).reshape(num_seq * T, D)
    err = (out[:num_seq * T].float() - ref).abs().max().item()
    untouched = out[num_seq * T:].abs().max().item()
    return max(err, untouched), 0.03, f'attn_temporal seq={num_seq} T={T} diag={causal_diag}'


def check_attn_spatial(B, N, T, use_cls, heads=12):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(3)
    D = heads * 64
    M = B * N * T
    qkv = (torch.randn(M + B, 3 * D, device=d, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.zeros(M + B, D, device=d, dtype=torch.bfloat16)
    out_cls = torch.zeros(B, T, D, device=d)
    ops.attn_spatial(qkv, out, out_cls if use_cls else None, B, N, T, heads, use_cls, M)
    torch.cuda.synchronize()
    x = qkv[:M].float().reshape(B, N, T, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)      # (3,B,T,h,N,64)
    if use_cls:
        c = qkv[M:].float().reshape(B, 3, heads, 64).permute(1, 0, 2, 3)              # (3,B,h,64)
        c = c[:, :, None, :, None, :].expand(3, B, T, heads, 1, 64)
        x = torch.cat([c, x], dim=4)
    o = _mha_ref(x[0], x[1], x[2], None)                                               # (B,T,h,S,64)
    err_c = 0.0
    if use_cls:
        ref_cls = o[:, :, :, 0, :].reshape(B, T, D)
        err_c = (out_cls - ref_cls).abs().max().item()
        # the frame-0 cls output is also written (bf16) into the cls row of `out`
        err_c = max(err_c, (out[M:].float() - out_cls[:, 0].to(torch.bfloat16).float()).abs().max().item())
        o = o[:, :, :, 1:, :]
    ref = o.permute(0, 3, 1, 2, 4).reshape(M, D)
    err = (out[:M].float() - ref).abs().max().item()
    return max(err, err_c), 0.03, f'attn_spatial B={B} N={N} T={T} cls={use_cls} (patch {err:.4f} cls {err_c:.4f})'


def check_patch_gather(B, T, Hf, Wf, normalize):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(4)
    P = 16
    fr = torch.rand(B, 3, T, Hf, Wf, device=d, generator=g)
    q = (torch.rand(B, 1, T, Hf, Wf, device=d, generator=g) > 0.7).float()
    Ho, Wo = Hf // P, Wf // P
    out = torch.empty(B * Ho * Wo * T, 4 * P * P, device=d, dtype=torch.bfloat16)
    ops.patch_gather(fr, q, out, P, normalize)
    torch.cuda.synchronize()
    f2 = (fr - 0.45) / 0.225 if normalize else fr
    x4 = torch.cat([f2, q], 1)
    if B % 2 == 0:   # shared-frames addressing: 2 queries per video, second half of the samples
        out2 = torch.empty_like(out[: out.shape[0] // 2])
        ops.patch_gather(fr[: B // 2].contiguous(), q[B // 2:].contiguous(), out2, P, normalize, 2, B // 2)
        torch.cuda.synchronize()
        x4b = torch.cat([f2[: B // 2][(torch.arange(B // 2, B) // 2).to(fr.device)], q[B // 2:]], 1)
        refb = x4b.reshape(B // 2, 4, T, Ho, P, Wo, P).permute(0, 3, 5, 2, 1, 4, 6).reshape(-1, 4 * P * P)
        if not torch.equal(out2.float(), refb.to(torch.bfloat16).float()):
            return float('inf'), 0.0, 'patch_gather shared-frames addressing mismatch'
    ref = x4.reshape(B, 4, T, Ho, P, Wo, P).permute(0, 3, 5, 2, 1, 4, 6).reshape(B * Ho * Wo * T, 4 * P * P)
    return (out.float() - ref.to(torch.bfloat16).float()).abs().max().item(), 1e-2 if normalize else 0.0, \
        f'patch_gather B={B} T={T} {Hf}x{Wf} norm={normalize}'


def check_patch_embed_fused(B, T, Hf, Wf, normalize, u8=False, qpv=1):
    """One-kernel patch embedding vs the three reference steps in fp32 torch (bf16-rounded operands, as the GEMM sees them):
    cat(frames, query) -> (optional normalisation) -> conv2d k=s=16 -> + bias + pos_embed[1+n] + time_embed[t]; cls rows."""
    d = _dev()
    g = torch.Generator(device=d).manual_seed(14)
    P, D = 16, 768
    V = (B + qpv - 1) // qpv
    fr = torch.rand(V, 3, T, Hf, Wf, device=d, generator=g)
    q = (torch.rand(B, 1, T, Hf, Wf, device=d, generator=g) > 0.7).float()
    scale = 1.0
    if u8:
        fr = (fr * 255).round().to(torch.uint8)
        q = q.to(torch.uint8)
        scale = 1.0 / 255.0
    Ho, Wo = Hf // P, Wf // P
    N = Ho * Wo
    M = B * N * T
    w = (torch.randn(D, 4 * P * P, device=d, generator=g) * 0.03).to(torch.bfloat16)
    cb, pos, tim, cls = (torch.randn(sh, device=d, generator=g) * 0.1 for sh in [(D,), (N + 1, D), (T, D), (D,)])
    X = torch.full((M + B, D), float('nan'), device=d)
    ops.patch_embed_fused(fr, q, w, cb, pos, tim, cls, X, P, normalize, qpv, 0, scale)
    torch.cuda.synchronize()
    f2 = fr.float() * scale
    if normalize:
        f2 = (f2 - 0.45) / 0.225
    f2 = f2[(torch.arange(B, device=d) // qpv)]
    x4 = torch.cat([f2, q.float()], 1).to(torch.bfloat16).float()
    pm = x4.reshape(B, 4, T, Ho, P, Wo, P).permute(0, 3, 5, 2, 1, 4, 6).reshape(M, 4 * P * P)
    ref = pm @ w.float().t() + (cb[None, None, None] + pos[None, 1:, None] + tim[None, None]).expand(B, N, T, D).reshape(M, D)
    err = (X[:M] - ref).abs().max().item()
    err_cls = (X[M:] - (cls + pos[0])[None]).abs().max().item()
    if not torch.isfinite(X).all():
        return float('inf'), 0.0, 'patch_embed_fused left rows unwritten'
    return max(err, err_cls), 2e-3, f'patch_embed_fused B={B} T={T} {Hf}x{Wf} norm={normalize} u8={u8} qpv={qpv} (cls {err_cls:.1e})'


def check_embed_init(B, N, T, D=768):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(5)
    cb, pos, tim, cls = (torch.randn(s, device=d, generator=g) for s in [(D,), (N + 1, D), (T, D), (D,)])
    X = torch.empty(B * N * T + B, D, device=d)
    ops.embed_init(X, cb, pos, tim, cls, B, N, T, D)
    torch.cuda.synchronize()
    ref = (cb[None, None, None] + pos[None, 1:, None] + tim[None, None]).expand(B, N, T, D).reshape(-1, D)
    ref = torch.cat([ref, (cls + pos[0])[None].expand(B, D)], 0)
    return (X - ref).abs().max().item(), 1e-6, f'embed_init B={B} N={N} T={T}'


def check_mask_upsample(B, T, Ho, Wo, stride, mode, C=3, P=16):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(6)
    pp = P // stride
    N = Ho * Wo
    ncol = C * pp * pp
    ld = (ncol + 3 + 63) // 64 * 64
    low = torch.randn(B * N * T, ld, device=d, generator=g)
    out = torch.empty(B, C, T, Ho * P, Wo * P, device=d)
    ops.mask_upsample(low, out, B, T, Ho, Wo, C, pp, stride, mode)
    flags = torch.empty(B, T, 3, device=d)
    ops.flag_mean(low, flags, B, N, T, 3, ncol)
    torch.cuda.synchronize()
    img = low[:, :ncol].reshape(B, Ho, Wo, T, C, pp, pp).permute(0, 3, 4, 1, 5, 2, 6).reshape(B * T, C, Ho * pp, Wo * pp)
    if stride > 1:
        img = F.interpolate(img, scale_factor=stride, mode='bilinear', align_corners=True) if mode == 0 else \
            F.interpolate(img, scale_factor=stride, mode='nearest')
    ref = img.reshape(B, T, C, Ho * P, Wo * P).transpose(1, 2)
    fref = low[:, ncol:ncol + 3].reshape(B, N, T, 3).mean(1)
    e = max((out - ref).abs().max().item(), (flags - fref).abs().max().item())
    return e, 2e-5, f'mask_upsample B={B} T={T} {Ho}x{Wo} stride={stride} mode={mode}'


ALL_CHECKS = [
    ('gemm_small_bf16', lambda: check_gemm(300, 256, 128, ops.EPI_BF16)),
    ('gemm_bn128_bf16', lambda: check_gemm(500, 384, 256, ops.EPI_BF16)),
    ('gemm_bn64_store', lambda: check_gemm(1000, 64, 768, ops.EPI_F32_STORE)),
    ('gemm_bn64_bf16', lambda: check_gemm(77, 64, 64, ops.EPI_BF16)),
    ('gemm_qkv', lambda: check_gemm(9030, 2304, 768, ops.EPI_BF16)),
    ('gemm_fc1_gelu', lambda: check_gemm(9001, 3072, 768, ops.EPI_BF16_GELU)),
    ('gemm_fc2_add', lambda: check_gemm(9001, 768, 3072, ops.EPI_F32_ADD)),
    ('gemm_proj_add', lambda: check_gemm(9000, 768, 768, ops.EPI_F32_ADD)),
    ('gemm_patch_add', lambda: check_gemm(9000, 768, 1024, ops.EPI_F32_ADD)),
    ('gemm_head_store', lambda: check_gemm(9000, 64, 768, ops.EPI_F32_STORE)),
    ('gemm_big_m', lambda: check_gemm(72008, 768, 768, ops.EPI_F32_ADD)),
    ('gemm_store_256', lambda: check_gemm(640, 512, 192, ops.EPI_F32_STORE)),
    ('gemm_one_row', lambda: check_gemm(1, 768, 768, ops.EPI_BF16)),
    ('gemm_gelu_accuracy', check_gelu_accuracy),
    ('layernorm', lambda: check_layernorm(9001)),
    ('layernorm_cast', lambda: check_layernorm(333, affine=False)),
    ('layernorm_1024', lambda: check_layernorm(100, 1024)),
    ('attn_temporal_T30', lambda: check_attn_temporal(601, 30, 0)),
    ('attn_temporal_T30_full', lambda: check_attn_temporal(64, 30, -1)),
    ('attn_temporal_T6_d1', lambda: check_attn_temporal(10, 6, 1)),
    ('attn_temporal_T60', lambda: check_attn_temporal(33, 60, 0)),
    ('attn_temporal_T33_d2', lambda: check_attn_temporal(5, 33, 2)),
    ('attn_temporal_T1', lambda: check_attn_temporal(4, 1, 0)),
    ('attn_spatial_301', lambda: check_attn_spatial(2, 300, 3, True)),
    ('attn_spatial_300_nocls', lambda: check_attn_spatial(1, 300, 2, False)),
    ('attn_spatial_7', lambda: check_attn_spatial(2, 6, 4, True)),
    ('attn_spatial_24_nocls', lambda: check_attn_spatial(1, 24, 5, False)),
    ('attn_spatial_1201', lambda: check_attn_spatial(1, 1200, 2, True)),
    ('attn_spatial_129', lambda: check_attn_spatial(1, 128, 1, True)),
    ('attn_spatial_128_nocls', lambda: check_attn_spatial(1, 128, 2, False)),
    ('attn_spatial_128_cls127', lambda: check_attn_spatial(2, 127, 2, True)),
    ('attn_spatial_256_cls255', lambda: check_attn_spatial(1, 255, 2, True)),
    ('attn_spatial_257', lambda: check_attn_spatial(1, 256, 2, True)),
    ('attn_spatial_258', lambda: check_attn_spatial(1, 257, 3, True)),
    ('attn_spatial_304', lambda: check_attn_spatial(1, 303, 2, True)),
    ('attn_spatial_304_nocls', lambda: check_attn_spatial(1, 304, 1, False)),
    ('attn_spatial_305_fallback', lambda: check_attn_spatial(1, 304, 1, True)),
    ('attn_spatial_1', lambda: check_attn_spatial(3, 1, 2, False)),
    ('attn_spatial_many_items', lambda: check_attn_spatial(8, 300, 30, True)),
    # several items per SM with one / two query tiles per item (the K/V double buffer and the tile stream wrap differently)
    ('attn_spatial_25_many_items', lambda: check_attn_spatial(2, 24, 30, True)),
    ('attn_spatial_24_many_items_nocls', lambda: check_attn_spatial(8, 24, 30, False)),
    ('attn_spatial_101_many_items', lambda: check_attn_spatial(4, 100, 10, True)),
    ('attn_spatial_131_many_items', lambda: check_attn_spatial(3, 130, 20, True)),
    ('attn_spatial_201_many_items', lambda: check_attn_spatial(4, 200, 10, True)),
    ('patch_gather', lambda: check_patch_gather(2, 3, 32, 48, False)),
    ('patch_gather_norm', lambda: check_patch_gather(1, 2, 240, 320, True)),
    ('embed_init', lambda: check_embed_init(2, 6, 4)),
    ('patch_embed_fused', lambda: check_patch_embed_fused(2, 3, 32, 48, False)),
    ('patch_embed_fused_norm_u8', lambda: check_patch_embed_fused(1, 2, 240, 320, True, u8=True)),
    ('patch_embed_fused_ragged_qpv', lambda: check_patch_embed_fused(3, 30, 64, 96, False, qpv=3)),
    ('patch_embed_fused_full', lambda: check_patch_embed_fused(8, 30, 240, 320, False, u8=True, qpv=3)),
    ('patch_embed_fused_T60_1200', lambda: check_patch_embed_fused(1, 60, 480, 640, False)),
    ('mask_upsample_bilinear', lambda: check_mask_upsample(2, 3, 15, 20, 4, 0)),
    ('mask_upsample_nearest', lambda: check_mask_upsample(1, 2, 2, 3, 4, 1)),
    ('mask_upsample_stride2', lambda: check_mask_upsample(1, 2, 2, 3, 2, 0)),
    ('mask_upsample_stride1', lambda: check_mask_upsample(1, 2, 2, 3, 1, 0)),
    ('mask_upsample_hires', lambda: check_mask_upsample(1, 2, 30, 40, 4, 0)),
    ('mask_upsample_stride8', lambda: check_mask_upsample(1, 1, 3, 2, 8, 0)),
]
