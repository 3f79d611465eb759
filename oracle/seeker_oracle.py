"""TEST INFRASTRUCTURE ONLY — fp32 functional restatement of TCOW's Seeker forward.

This is the parity oracle for the CUDA path.  It restates, in plain fp32
PyTorch (CPU by default), what the reference computes in
  model/seeker.py:24-25            (Seeker.forward: delegation)
  model/mask_tracker.py:92-142     (QueryMaskTracker.forward: query concat, head, pool+upsample, flags)
  model/vision_tf.py:68-169        (DenseTimeSformer.forward: patch embed, embeddings, block loop)
  third_party/TimeSformer/timesformer/models/vit.py:64-123   (Attention, causal mask)
  third_party/TimeSformer/timesformer/models/vit.py:155-217  (Block, divided space-time)
  third_party/TimeSformer/timesformer/models/vit.py:220-241  (PatchEmbed)
All arithmetic in the reference is torch itself (no third-party numeric dependency).

Pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c).  The oracle is therefore pinned against outputs of the
reference ITSELF, imported and run in the build container by
``oracle/make_golden.py`` (committed fixtures: ``tests/golden/*.npz``), and
``tests/test_oracle.py`` re-checks the oracle against those fixtures on every run.

It is written op-by-op in the reference's own order (separate proj and
temporal_fc, un-folded avg_pool + interpolate, per-token flag linear then mean)
so that it is an independent check of the algebraic shortcuts the CUDA path takes.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

PREFIX = 'seeker.tracker_backbone.timesformer.model.'
TIMESFORMER_MEAN = 0.45   # model/vision_tf.py:23
TIMESFORMER_STD = 0.225   # model/vision_tf.py:24


def seeker_forward(sd, input_frames, query_mask, *, causal_attention=1, patch_size=16,
                   num_heads=12, depth=12, track_map_stride=4, track_map_resize='bilinear',
                   output_channels=3, flag_channels=3, norm_embeddings=False,
                   pretrained_norm=False, eps=1e-6, drop_path=None):
    """Return (output_mask (B,C,T,Hf,Wf) fp32 logits, output_flags (B,T,F) fp32 or None).

    ``sd`` is ``Seeker.state_dict()`` (251 tensors, layout of SURVEY.md §8b).
    ``drop_path`` (training mode only): per block ``None`` or ``dict(t=(B*N,), s=(B*T,), m=(B,))`` of 0/1 keep masks
    plus ``keep`` (the keep probability) — DropPath of vit_utils.py:139-164 with the draws made explicit: each branch
    output is multiplied by mask/keep along its leading dim (vit.py:172 (b h w), :186 (b t), :216 b).
    """
    g = lambda k: sd[PREFIX + k].to(torch.float32)
    P = patch_size
    # mask_tracker.py:102-108 — cast, clone, concat query as the 4th channel.
    rgb = input_frames.to(torch.float32).clone()
    qm = query_mask.to(torch.float32)
    assert qm.shape[1] == 1
    B, _, T, Hf, Wf = rgb.shape
    assert Hf % P == 0 and Wf % P == 0
    # vision_tf.py:81-89 — RGB normalisation only when the backbone is flagged pretrained.
    if pretrained_norm:
        rgb[:, 0:3] = (rgb[:, 0:3] - TIMESFORMER_MEAN) / TIMESFORMER_STD
    x4 = torch.cat([rgb, qm], dim=1)                                   # (B,4,T,Hf,Wf)
    Ho, Wo = Hf // P, Wf // P
    N = Ho * Wo
    D = g('cls_token').shape[-1]
    hd = D // num_heads

    # vit.py:235-241 — Conv2d(k=s=P) over every frame, tokens in (ph, pw) order.
    frames = x4.permute(0, 2, 1, 3, 4).reshape(B * T, x4.shape[1], Hf, Wf)
    tok = F.conv2d(frames, g('patch_embed.proj.weight'), g('patch_embed.proj.bias'), stride=P)
    tok = tok.flatten(2).transpose(1, 2)                               # (B*T, N, D)
    # vision_tf.py:99-118 — cls token + positional embedding.
    assert g('pos_embed').shape[1] == N + 1 and g('time_embed').shape[1] == T
    tok = torch.cat([g('cls_token').expand(B * T, -1, -1), tok], dim=1) + g('pos_embed')
    cls = tok[:B, 0, :].unsqueeze(1)                                   # (B,1,D)  vision_tf.py:121
    tok = tok[:, 1:].reshape(B, T, N, D).permute(0, 2, 1, 3)           # (B,N,T,D)
    tok = tok + g('time_embed')[0][None, None]                         # vision_tf.py:133-134
    x = torch.cat([cls, tok.reshape(B, N * T, D)], dim=1)              # (B, 1+N*T, D), order n*T+t

    def ln(v, p):
        return F.layer_norm(v, (D,), g(p + '.weight'), g(p + '.bias'), eps)

    def lin(v, p):
        return F.linear(v, g(p + '.weight'), g(p + '.bias'))

    def attention(v, p, causal):                                       # vit.py:78-123
        Bn, S, C = v.shape
        qkv = lin(v, p + '.qkv').reshape(Bn, S, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        q, k, w = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * hd ** -0.5
        if causal > 0:                                                 # vit.py:93-99
            m = torch.ones(S, S, dtype=torch.bool, device=v.device)
            m = m.tril() if causal <= 2 else m.tril(diagonal=causal - 2)
            attn = attn.masked_fill(~m, -1e10)
        attn = attn.softmax(dim=-1)
        o = (attn @ w).transpose(1, 2).reshape(Bn, S, C)
        return lin(o, p + '.proj')

    def dpath(v, i, which):                                            # vit_utils.py:139-164
        if drop_path is None or drop_path[i] is None:
            return v
        m = drop_path[i][which].to(v.dtype).reshape(-1, *([1] * (v.dim() - 1)))
        return v.div(drop_path[i]['keep']) * m

    for i in range(depth):                                             # vit.py:165-217
        b = f'blocks.{i}.'
        xt = x[:, 1:, :].reshape(B * N, T, D)
        rt = dpath(attention(ln(xt, b + 'temporal_norm1'), b + 'temporal_attn', causal_attention), i, 't')
        rt = lin(rt.reshape(B, N * T, D), b + 'temporal_fc')
        xt = x[:, 1:, :] + rt
        init_cls = x[:, 0, :].unsqueeze(1)
        xs = xt.reshape(B, N, T, D).permute(0, 2, 1, 3).reshape(B * T, N, D)
        if causal_attention in (0, 1):
            c = init_cls.repeat(1, T, 1).reshape(B * T, 1, D)
            rs = dpath(attention(ln(torch.cat([c, xs], 1), b + 'norm1'), b + 'attn', 0), i, 's')
            c = rs[:, 0, :].reshape(B, T, D)
            c = c.mean(1, keepdim=True) if causal_attention == 0 else c[:, 0:1, :]
            rs = rs[:, 1:, :]
        elif causal_attention >= 2 or causal_attention == -1:
            c = torch.zeros_like(init_cls)
            rs = dpath(attention(ln(xs, b + 'norm1'), b + 'attn', 0), i, 's')
        else:
            raise ValueError(causal_attention)
        rs = rs.reshape(B, T, N, D).permute(0, 2, 1, 3).reshape(B, N * T, D)
        x = torch.cat([init_cls, xt], 1) + torch.cat([c, rs], 1)
        h = F.gelu(lin(ln(x, b + 'norm2'), b + 'mlp.fc1'))             # nn.GELU() = exact erf
        x = x + dpath(lin(h, b + 'mlp.fc2'), i, 'm')

    if norm_embeddings:                                                # vision_tf.py:152-153
        x = ln(x, 'norm')
    feat = x[:, 1:].reshape(B, Ho, Wo, T, D).permute(0, 3, 1, 2, 4)     # (B,T,Ho,Wo,D)

    # mask_tracker.py:112-132
    Wt, bt = sd['seeker.tracker_post_linear.weight'].float(), sd['seeker.tracker_post_linear.bias'].float()
    C = output_channels
    patches = F.linear(feat, Wt, bt).reshape(B, T, Ho, Wo, C, P, P)
    mask = patches.permute(0, 4, 1, 2, 5, 3, 6).reshape(B, C, T, Hf, Wf)
    if track_map_stride > 1:
        m2 = mask.permute(0, 2, 1, 3, 4).reshape(B * T, C, Hf, Wf)
        m2 = F.avg_pool2d(m2, track_map_stride, track_map_stride)
        if track_map_resize == 'nearest':
            m2 = F.interpolate(m2, scale_factor=track_map_stride, mode='nearest')
        elif track_map_resize == 'bilinear':
            m2 = F.interpolate(m2, scale_factor=track_map_stride, mode='bilinear', align_corners=True)
        mask = m2.reshape(B, T, C, Hf, Wf).permute(0, 2, 1, 3, 4)
    flags = None
    if flag_channels > 0:                                              # mask_tracker.py:135-137
        fl = F.linear(feat, sd['seeker.flag_post_linear.weight'].float(),
                      sd['seeker.flag_post_linear.bias'].float())
        flags = fl.mean(dim=[-2, -3])
    return mask.contiguous(), flags
