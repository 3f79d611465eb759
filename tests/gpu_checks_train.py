"""Per-kernel checks of the training-step kernels (weight-gradient GEMM, aux GEMM epilogues, LayerNorm fwd-save /
bwd, column sums, embedding and mask-head adjoints, attention backward) against fp32 PyTorch autograd on the GPU
(TF32 off).  Each check returns (err, tol, detail).  Used by tests/test_train_kernels_gpu.py."""
import torch
import torch.nn.functional as F

from tcow_b200 import ops


def _dev():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device('cuda:0')


def _ws(d):
    return torch.empty(ops.train_workspace_floats(4096), device=d)


def _relerr(got, ref):
    return ((got - ref).norm() / ref.norm().clamp_min(1e-20)).item()


def check_wgrad(R, N1, N2, accumulate=False, pad=0, bias=False, tickets=False):
    d = _dev()
    sched = torch.zeros(2, device=d, dtype=torch.int32) if tickets else None
    g = torch.Generator(device=d).manual_seed(11)
    dy_full = (torch.randn(R, N1 + pad, device=d, generator=g) * 0.3).to(torch.bfloat16)
    x_full = (torch.randn(R, N2 + pad, device=d, generator=g) * 0.7).to(torch.bfloat16)
    dy, x = dy_full[:, :N1], x_full[:, :N2]          # pitch != width when pad > 0
    base = torch.randn(N1, N2, device=d, generator=g) if accumulate else torch.zeros(N1, N2, device=d)
    dw = base.clone()
    db0 = torch.randn(N1, device=d, generator=g) if accumulate else torch.zeros(N1, device=d)
    db = db0.clone() if bias else None
    ops.gemm_wgrad(dy, x, dw, db, sched)
    if tickets:      # the counters come back to zero: a second launch on the same counters must work
        ops.gemm_wgrad(dy, x, dw, db, sched)
        torch.cuda.synchronize()
        if int(sched.abs().sum()) != 0:
            return float('inf'), 0.0, 'wgrad ticket counters not reset'
        dw.sub_(base).mul_(0.5).add_(base)
        if db is not None:
            db.sub_(db0).mul_(0.5).add_(db0)
    torch.cuda.synchronize()
    ref = base + dy.float().t() @ x.float()
    err = (dw - ref).abs().max().item()
    tol = 2e-5 * ref.abs().max().item() + 1e-4 * (R / 1000) ** 0.5
    extra = ''
    if bias:      # bias gradient from the same pass: db += dy.sum(0)
        ref_b = db0 + dy.float().sum(0)
        err_b = (db - ref_b).abs().max().item()
        extra = f' db err {err_b:.2e}'
        err = max(err, err_b * tol / (1e-5 * ref_b.abs().max().item() + 2e-4 * (R / 1000) ** 0.5))
    return err, tol, f'wgrad R={R} N1={N1} N2={N2} acc={accumulate} pad={pad} (rel {_relerr(dw, ref):.2e}){extra}'


def check_gemm_gelu_aux(M, N, K):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(12)
    a = (torch.randn(M, K, device=d, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=d, generator=g) * 0.08).to(torch.bfloat16)
    bias = torch.randn(N, device=d, generator=g) * 0.1
    out = torch.full((M, N), float('nan'), device=d, dtype=torch.bfloat16)
    aux = torch.full((M, N), float('nan'), device=d, dtype=torch.bfloat16)
    ops.gemm_aux(a, w, bias, out, aux, ops.EPI_BF16_GELU_AUX)
    torch.cuda.synchronize()
    z = a.float() @ w.float().t() + bias
    ez = (aux.float() - z).abs().max().item()
    eh = (out.float() - F.gelu(z)).abs().max().item()
    tol = 2.0 ** -8 * z.abs().max().item() + 1e-3
    bad = float('inf') if (torch.isnan(aux.float()).any() or torch.isnan(out.float()).any()) else max(ez, eh)
    return bad, tol, f'gemm_gelu_aux M={M} N={N} K={K} (z {ez:.4f}, gelu {eh:.4f})'


def check_gemm_dgelu(M, N, K):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(13)
    a = (torch.randn(M, K, device=d, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=d, generator=g) * 0.08).to(torch.bfloat16)
    z = (torch.randn(M, N, device=d, generator=g) * 2.0).to(torch.bfloat16)
    out = torch.full((M, N), float('nan'), device=d, dtype=torch.bfloat16)
    ops.gemm_aux(a, w, None, out, z, ops.EPI_BF16_DGELU)
    torch.cuda.synchronize()
    zz = z.float().requires_grad_(True)
    F.gelu(zz).sum().backward()
    ref = (a.float() @ w.float().t()) * zz.grad
    err = (out.float() - ref).abs()
    bad = float('inf') if torch.isnan(err).any() else err.max().item()
    return bad, 2.0 ** -8 * ref.abs().max().item() + 1e-3, f'gemm_dgelu M={M} N={N} K={K}'


def check_gemm_add_scaled(M, N, K, bias2=True):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(21)
    a = (torch.randn(M, K, device=d, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=d, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=d, generator=g) * 0.1
    b2 = torch.randn(N, device=d, generator=g) * 0.1 if bias2 else None
    rs = (torch.rand(M, device=d, generator=g) > 0.4).float() / 0.6
    bs = torch.rand(M, device=d, generator=g)
    x0 = torch.randn(M, N, device=d, generator=g)
    x = x0.clone()
    ops.gemm_add_scaled(a, w, bias, b2, rs, bs, x)
    torch.cuda.synchronize()
    ref = x0 + rs[:, None] * (a.float() @ w.float().t()) + bs[:, None] * bias + (b2 if bias2 else 0)
    return (x - ref).abs().max().item(), 3e-4 * max(1.0, (K / 768) ** 0.5), f'gemm_add_scaled M={M} N={N} K={K} bias2={bias2}'


def check_scale_rows(rows, N):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(22)
    x = torch.randn(rows, N, device=d, generator=g).to(torch.bfloat16)
    sc = (torch.rand(rows, device=d, generator=g) > 0.5).float() * 2
    out = torch.empty_like(x)
    ops.scale_rows(x, sc, out)
    torch.cuda.synchronize()
    return (out.float() - (x.float() * sc[:, None]).to(torch.bfloat16).float()).abs().max().item(), 0.0, f'scale_rows {rows}x{N}'


def check_ln_train_bwd(rows, D=768, accumulate=True):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(14)
    x = torch.randn(rows, D, device=d, generator=g) * 2 + 0.3
    gm = 1 + 0.1 * torch.randn(D, device=d, generator=g)
    bt = 0.1 * torch.randn(D, device=d, generator=g)
    y = torch.empty(rows, D, device=d, dtype=torch.bfloat16)
    xhat = torch.empty_like(y)
    rstd = torch.empty(rows, device=d)
    ops.layernorm_train(x, gm, bt, y, xhat, rstd)
    dy = (torch.randn(rows, D, device=d, generator=g) * 0.1).to(torch.bfloat16)
    G0 = torch.randn(rows, D, device=d, generator=g) * 0.1
    G = G0.clone()
    Gb = torch.empty_like(y)
    dgm = torch.full((D,), 0.5, device=d)
    dbt = torch.full((D,), -0.25, device=d)
    ops.layernorm_bwd(dy, xhat, rstd, gm, G, Gb, dgm, dbt, _ws(d), accumulate)
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    gr = gm.clone().requires_grad_(True)
    br = bt.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br, 1e-6)
    yr.backward(dy.float())
    e_y = (y.float() - yr.detach()).abs().max().item() / (2.0 ** -8 * yr.abs().max().item() + 1e-3)
    ref_G = (G0 if accumulate else 0) + xr.grad
    e_dx = _relerr(G, ref_G) / 6e-3            # xhat is bf16: ~2^-9 relative per element
    e_gb = (Gb.float() - G).abs().max().item() / (2.0 ** -8 * G.abs().max().item() + 1e-6)
    e_dg = _relerr(dgm - 0.5, gr.grad) / 6e-3
    e_db = _relerr(dbt + 0.25, br.grad) / 1e-4
    return max(e_y, e_dx, e_gb, e_dg, e_db), 1.0, \
        f'ln_train_bwd rows={rows} acc={accumulate} (y {e_y:.2f} dx {e_dx:.2f} gb {e_gb:.2f} dgamma {e_dg:.2f} dbeta {e_db:.2f} of tol)'


def check_ln_bwd_scaled(rows, D=768):
    """The stochastic-depth variant also writes Gs = next_scale * G (bf16); everything else must equal the plain kernel."""
    d = _dev()
    g = torch.Generator(device=d).manual_seed(23)
    x = torch.randn(rows, D, device=d, generator=g)
    gm = 1 + 0.1 * torch.randn(D, device=d, generator=g)
    bt = 0.1 * torch.randn(D, device=d, generator=g)
    y, xhat, rstd = torch.empty(rows, D, device=d, dtype=torch.bfloat16), torch.empty(rows, D, device=d, dtype=torch.bfloat16), torch.empty(rows, device=d)
    ops.layernorm_train(x, gm, bt, y, xhat, rstd)
    dy = (torch.randn(rows, D, device=d, generator=g) * 0.1).to(torch.bfloat16)
    G0 = torch.randn(rows, D, device=d, generator=g) * 0.1
    sc = (torch.rand(rows, device=d, generator=g) > 0.3).float() / 0.7
    outs = []
    for scaled in (False, True):
        G, Gb, Gs = G0.clone(), torch.empty_like(y), torch.full_like(y, 7.0)
        dgm, dbt = torch.zeros(D, device=d), torch.zeros(D, device=d)
        ops.layernorm_bwd(dy, xhat, rstd, gm, G, Gb, dgm, dbt, _ws(d), True, sc if scaled else None, Gs if scaled else None)
        outs.append((G, Gb, Gs, dgm, dbt))
    torch.cuda.synchronize()
    (G1, Gb1, _, dg1, db1), (G2, Gb2, Gs2, dg2, db2) = outs
    same = torch.equal(G1, G2) and torch.equal(Gb1, Gb2) and torch.equal(dg1, dg2) and torch.equal(db1, db2)
    ref = (sc[:, None] * G2).to(torch.bfloat16).float()
    err = (Gs2.float() - ref).abs().max().item()
    return (0.0 if same else float('inf')) + err, 2.0 ** -8 * ref.abs().max().item(), f'ln_bwd_scaled rows={rows} (identical to plain: {same}, Gs err {err:.2e})'


def check_colsum(rows, N, pad=0):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(15)
    xf = torch.randn(rows, N + pad, device=d, generator=g).to(torch.bfloat16)
    x = xf[:, :N]
    out = torch.full((N,), 2.0, device=d)
    ops.colsum(x, out, _ws(d), True)
    out2 = torch.full((N,), 2.0, device=d)
    ops.colsum(x, out2, _ws(d), False)
    torch.cuda.synchronize()
    ref = x.float().sum(0)
    err = max((out - 2.0 - ref).abs().max().item(), (out2 - ref).abs().max().item())
    return err, 1e-3 * (rows / 1000) ** 0.5 + 1e-4, f'colsum rows={rows} N={N}'


def check_embed_bwd(B, N, T, D=768):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(16)
    G = torch.randn(B * N * T + B, D, device=d, generator=g)
    dpos = torch.zeros(N + 1, D, device=d)
    dtime = torch.zeros(T, D, device=d)
    ops.embed_bwd(G, dpos, dtime, dpos[0], _ws(d), B, N, T, D, True)
    torch.cuda.synchronize()
    g4 = G[:B * N * T].reshape(B, N, T, D)
    e = max((dpos[1:] - g4.sum((0, 2))).abs().max().item(), (dtime - g4.sum((0, 1))).abs().max().item(),
            (dpos[0] - G[B * N * T:].sum(0)).abs().max().item())
    return e, 1e-3, f'embed_bwd B={B} N={N} T={T}'


def check_mask_head_bwd(B, T, Ho, Wo, stride, mode, C=3, P=16, F_=3):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(17)
    pp = P // stride
    N = Ho * Wo
    ncol = C * pp * pp
    ld = (ncol + F_ + 63) // 64 * 64
    low = torch.randn(B * N * T, ld, device=d, generator=g, requires_grad=True)
    img = low[:, :ncol].reshape(B, Ho, Wo, T, C, pp, pp).permute(0, 3, 4, 1, 5, 2, 6).reshape(B * T, C, Ho * pp, Wo * pp)
    if stride > 1:
        img = F.interpolate(img, scale_factor=stride, mode='bilinear', align_corners=True) if mode == 0 else \
            F.interpolate(img, scale_factor=stride, mode='nearest')
    mask = img.reshape(B, T, C, Ho * P, Wo * P).transpose(1, 2)
    flags = low[:, ncol:ncol + F_].reshape(B, N, T, F_).mean(1)
    d_out = torch.randn(B, C, T, Ho * P, Wo * P, device=d, generator=g)
    d_flags = torch.randn(B, T, F_, device=d, generator=g)
    ((mask * d_out).sum() + (flags * d_flags).sum()).backward()
    d_low = torch.full((B * N * T, ld), float('nan'), device=d, dtype=torch.bfloat16)
    ops.mask_head_bwd(d_out, d_flags, d_low, B, T, Ho, Wo, C, pp, stride, mode, F_, ncol)
    torch.cuda.synchronize()
    err = (d_low.float() - low.grad).abs()
    bad = float('inf') if torch.isnan(err).any() else err.max().item()
    return bad, 2.0 ** -8 * low.grad.abs().max().item() + 1e-5, f'mask_head_bwd B={B} T={T} {Ho}x{Wo} stride={stride} mode={mode}'


def _mha(q, k, v, mask):
    a = (q @ k.transpose(-1, -2)) * 0.125
    if mask is not None:
        a = a.masked_fill(~mask, float('-inf'))
    return a.softmax(-1) @ v


def check_attn_temporal_bwd(num_seq, T, causal_diag, heads=12):
    d = _dev()
    g = torch.Generator(device=d).manual_seed(18)
    D = heads * 64
    R = num_seq * T
    qkv = (torch.randn(R + 2, 3 * D, device=d, generator=g) * 1.2).to(torch.bfloat16)
    out = torch.zeros(R + 2, D, device=d, dtype=torch.bfloat16)
    ops.attn_temporal(qkv, out, num_seq, T, heads, causal_diag)
    d_out = (torch.randn(R + 2, D, device=d, generator=g) * 0.2).to(torch.bfloat16)
    d_qkv = torch.zeros(R + 2, 3 * D, device=d, dtype=torch.bfloat16)
    ops.attn_temporal_bwd(qkv, out, d_out, d_qkv, num_seq, T, heads, causal_diag)
    torch.cuda.synchronize()
    xr = qkv[:R].float().clone().requires_grad_(True)
    x = xr.reshape(num_seq, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    mask = torch.ones(T, T, dtype=torch.bool, device=d).tril(causal_diag) if causal_diag >= 0 else None
    o = _mha(x[0], x[1], x[2], mask).permute(0, 2, 1, 3).reshape(R, D)
    o.backward(d_out[:R].float())
    rel = _relerr(d_qkv[:R].float(), xr.grad)
    mx = (d_qkv[:R].float() - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    untouched = d_qkv[R:].abs().max().item()
    return max(rel / 2e-2, mx / 3e-2, untouched * 1e6), 1.0, \
        f'attn_temporal_bwd seq={num_seq} T={T} diag={causal_diag} (rel-L2 {rel:.4f}, max/absmax {mx:.4f})'


def check_attn_spatial_bwd(B, N, T, use_cls, cls_mode=1, heads=12):
    """cls_mode 1: only frame 0's cls output feeds the loss (causal_attention==1); 0: the mean over frames."""
    d = _dev()
    g = torch.Generator(device=d).manual_seed(19)
    D = heads * 64
    M = B * N * T
    qkv = (torch.randn(M + B, 3 * D, device=d, generator=g) * 1.2).to(torch.bfloat16)
    out = torch.zeros(M + B, D, device=d, dtype=torch.bfloat16)
    out_cls = torch.zeros(B, T, D, device=d)
    lse = torch.zeros(B * T * heads, 304, device=d)
    ops.attn_spatial_train(qkv, out, out_cls if use_cls else None, lse, B, N, T, heads, use_cls, M)
    d_out = (torch.randn(M + B, D, device=d, generator=g) * 0.2).to(torch.bfloat16)
    d_out_cls = torch.zeros(B, T, D, device=d)
    if use_cls:
        ops.cls_merge_bwd(d_out, d_out_cls, B, T, D, M, cls_mode)
    d_qkv = torch.zeros(M + B, 3 * D, device=d, dtype=torch.bfloat16)
    d_cls = torch.zeros(ops.spatial_bwd_scratch_floats(B, T, heads), device=d)
    ops.attn_spatial_bwd(qkv, out, out_cls if use_cls else None, d_out, d_out_cls if use_cls else None, lse, d_qkv,
                         d_cls, B, N, T, heads, use_cls, M)
    torch.cuda.synchronize()
    xr = qkv.float().clone().requires_grad_(True)
    x = xr[:M].reshape(B, N, T, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)               # (3,B,T,h,N,64)
    if use_cls:
        c = xr[M:].reshape(B, 3, heads, 64).permute(1, 0, 2, 3)[:, :, None, :, None, :].expand(3, B, T, heads, 1, 64)
        x = torch.cat([x, c], dim=4)                                                   # cls is the LAST token here
    o = _mha(x[0], x[1], x[2], None)                                                   # (B,T,h,S,64)
    loss = (o[:, :, :, :N].permute(0, 3, 1, 2, 4).reshape(M, D) * d_out[:M].float()).sum()
    if use_cls:
        oc = o[:, :, :, N].reshape(B, T, D)
        ocm = oc[:, 0] if cls_mode == 1 else oc.mean(1)
        loss = loss + (ocm * d_out[M:].float()).sum()
        # forward sanity: lse really is the log-sum-exp (checked through the cls outputs)
    loss.backward()
    R = M + (B if use_cls else 0)
    rel = _relerr(d_qkv[:R].float(), xr.grad[:R])
    mx = (d_qkv[:R].float() - xr.grad[:R]).abs().max().item() / xr.grad[:R].abs().max().item()
    rel_c = _relerr(d_qkv[M:R].float(), xr.grad[M:R]) if use_cls else 0.0
    return max(rel / 2e-2, mx / 3e-2, rel_c / 2e-2), 1.0, \
        f'attn_spatial_bwd B={B} N={N} T={T} cls={use_cls}/{cls_mode} (rel-L2 {rel:.4f}, cls rows {rel_c:.4f}, max/absmax {mx:.4f})'


TRAIN_CHECKS = [
    ('wgrad_small', lambda: check_wgrad(300, 128, 256)),
    ('wgrad_ragged_rows', lambda: check_wgrad(1001, 192, 512, accumulate=True)),
    ('wgrad_n1_64', lambda: check_wgrad(5000, 64, 768)),
    ('wgrad_bn64', lambda: check_wgrad(777, 256, 192)),
    ('wgrad_pitch', lambda: check_wgrad(900, 128, 256, pad=64)),
    ('wgrad_qkv', lambda: check_wgrad(54006, 2304, 768)),
    ('wgrad_fc1', lambda: check_wgrad(54006, 3072, 768)),
    ('wgrad_fc2', lambda: check_wgrad(54006, 768, 3072, accumulate=True)),
    ('wgrad_proj', lambda: check_wgrad(54000, 768, 768)),
    ('wgrad_patch', lambda: check_wgrad(54000, 768, 1024)),
    ('wgrad_one_row', lambda: check_wgrad(1, 64, 64)),
    ('wgrad_bias_small', lambda: check_wgrad(300, 128, 256, bias=True)),
    ('wgrad_bias_ragged', lambda: check_wgrad(1001, 192, 512, accumulate=True, bias=True)),
    ('wgrad_bias_n1_64', lambda: check_wgrad(5000, 64, 768, bias=True)),
    ('wgrad_bias_bn64', lambda: check_wgrad(777, 256, 192, bias=True)),
    ('wgrad_bias_qkv', lambda: check_wgrad(54006, 2304, 768, bias=True)),
    ('wgrad_bias_fc2', lambda: check_wgrad(54006, 768, 3072, accumulate=True, bias=True)),
    ('wgrad_tickets_small', lambda: check_wgrad(300, 128, 256, bias=True, tickets=True)),
    ('wgrad_tickets_bn64', lambda: check_wgrad(777, 256, 192, tickets=True)),
    ('wgrad_tickets_qkv', lambda: check_wgrad(54006, 2304, 768, bias=True, tickets=True)),
    ('wgrad_tickets_fc2', lambda: check_wgrad(54006, 768, 3072, accumulate=True, bias=True, tickets=True)),
    ('wgrad_tickets_proj', lambda: check_wgrad(54000, 768, 768, tickets=True)),
    ('gemm_gelu_aux', lambda: check_gemm_gelu_aux(9001, 3072, 768)),
    ('gemm_gelu_aux_small', lambda: check_gemm_gelu_aux(100, 128, 64)),
    ('gemm_dgelu', lambda: check_gemm_dgelu(9001, 3072, 768)),
    ('gemm_dgelu_small', lambda: check_gemm_dgelu(77, 64, 128)),
    ('gemm_add_scaled_proj', lambda: check_gemm_add_scaled(9008, 768, 768)),
    ('gemm_add_scaled_fc2', lambda: check_gemm_add_scaled(9001, 768, 3072, bias2=False)),
    ('gemm_add_scaled_small', lambda: check_gemm_add_scaled(100, 128, 64)),
    ('scale_rows', lambda: check_scale_rows(9001, 768)),
    ('ln_train_bwd', lambda: check_ln_train_bwd(9001)),
    ('ln_train_bwd_noacc', lambda: check_ln_train_bwd(300, accumulate=False)),
    ('ln_train_bwd_1024', lambda: check_ln_train_bwd(100, 1024)),
    ('ln_train_bwd_one', lambda: check_ln_train_bwd(1)),
    ('ln_bwd_scaled', lambda: check_ln_bwd_scaled(9001)),
    ('ln_bwd_scaled_small', lambda: check_ln_bwd_scaled(5)),
    ('colsum', lambda: check_colsum(9001, 768)),
    ('colsum_wide', lambda: check_colsum(3000, 3072, pad=8)),
    ('colsum_64', lambda: check_colsum(5000, 64)),
    ('embed_bwd', lambda: check_embed_bwd(2, 6, 4)),
    ('embed_bwd_full', lambda: check_embed_bwd(2, 300, 30)),
    ('mask_head_bwd_bilinear', lambda: check_mask_head_bwd(2, 3, 15, 20, 4, 0)),
    ('mask_head_bwd_nearest', lambda: check_mask_head_bwd(1, 2, 2, 3, 4, 1)),
    ('mask_head_bwd_stride2', lambda: check_mask_head_bwd(1, 2, 2, 3, 2, 0)),
    ('mask_head_bwd_stride1', lambda: check_mask_head_bwd(1, 2, 2, 3, 1, 0)),
    ('mask_head_bwd_noflags', lambda: check_mask_head_bwd(1, 2, 2, 3, 4, 0, F_=0)),
    ('attn_temporal_bwd_T30', lambda: check_attn_temporal_bwd(301, 30, 0)),
    ('attn_temporal_bwd_T30_full', lambda: check_attn_temporal_bwd(17, 30, -1)),
    ('attn_temporal_bwd_T6_d1', lambda: check_attn_temporal_bwd(10, 6, 1)),
    ('attn_temporal_bwd_T60', lambda: check_attn_temporal_bwd(9, 60, 0)),
    ('attn_temporal_bwd_T33_d2', lambda: check_attn_temporal_bwd(5, 33, 2)),
    ('attn_temporal_bwd_T1', lambda: check_attn_temporal_bwd(4, 1, 0)),
    ('attn_spatial_bwd_301', lambda: check_attn_spatial_bwd(2, 300, 3, True, 1)),
    ('attn_spatial_bwd_301_mean', lambda: check_attn_spatial_bwd(1, 300, 2, True, 0)),
    ('attn_spatial_bwd_300_nocls', lambda: check_attn_spatial_bwd(1, 300, 2, False)),
    ('attn_spatial_bwd_7', lambda: check_attn_spatial_bwd(2, 6, 4, True, 1)),
    ('attn_spatial_bwd_24_nocls', lambda: check_attn_spatial_bwd(1, 24, 5, False)),
    ('attn_spatial_bwd_129', lambda: check_attn_spatial_bwd(1, 128, 2, True, 0)),
    ('attn_spatial_bwd_304', lambda: check_attn_spatial_bwd(1, 303, 2, True, 1)),
]
