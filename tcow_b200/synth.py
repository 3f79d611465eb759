"""Deterministic synthetic weights and clips for parity tests and the benchmark.

The reference's own initialisation leaves ``time_embed``, every ``temporal_fc``,
every bias and every LayerNorm affine at 0 / identity
(third_party/TimeSformer/timesformer/models/vit.py:264-306), which would hide the
temporal path, the causal mask, the biases and the LN affines from a parity check
(SURVEY.md §0 traps 1-2).  These generators therefore fill *every* tensor of the
251-key ``Seeker.state_dict()`` layout from one seeded CPU generator, independent
of the reference's RNG order, so the same weights can be rebuilt anywhere.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

BACKBONE_PREFIX = 'seeker.tracker_backbone.timesformer.model.'


def state_dict_shapes(num_frames=30, frame_height=240, frame_width=320, patch_size=16,
                      in_channels=4, embed_dim=768, depth=12, mlp_ratio=4,
                      output_channels=3, flag_channels=3):
    """Ordered {key: shape} of Seeker.state_dict() (train.py:281 'net_seeker')."""
    D, P = embed_dim, patch_size
    N = (frame_height // P) * (frame_width // P)
    s = OrderedDict()
    b = BACKBONE_PREFIX
    s[b + 'cls_token'] = (1, 1, D)
    s[b + 'pos_embed'] = (1, N + 1, D)
    s[b + 'time_embed'] = (1, num_frames, D)
    s[b + 'patch_embed.proj.weight'] = (D, in_channels, P, P)
    s[b + 'patch_embed.proj.bias'] = (D,)
    for i in range(depth):
        k = f'{b}blocks.{i}.'
        s[k + 'norm1.weight'] = (D,)
        s[k + 'norm1.bias'] = (D,)
        s[k + 'attn.qkv.weight'] = (3 * D, D)
        s[k + 'attn.qkv.bias'] = (3 * D,)
        s[k + 'attn.proj.weight'] = (D, D)
        s[k + 'attn.proj.bias'] = (D,)
        s[k + 'temporal_norm1.weight'] = (D,)
        s[k + 'temporal_norm1.bias'] = (D,)
        s[k + 'temporal_attn.qkv.weight'] = (3 * D, D)
        s[k + 'temporal_attn.qkv.bias'] = (3 * D,)
        s[k + 'temporal_attn.proj.weight'] = (D, D)
        s[k + 'temporal_attn.proj.bias'] = (D,)
        s[k + 'temporal_fc.weight'] = (D, D)
        s[k + 'temporal_fc.bias'] = (D,)
        s[k + 'norm2.weight'] = (D,)
        s[k + 'norm2.bias'] = (D,)
        s[k + 'mlp.fc1.weight'] = (mlp_ratio * D, D)
        s[k + 'mlp.fc1.bias'] = (mlp_ratio * D,)
        s[k + 'mlp.fc2.weight'] = (D, mlp_ratio * D)
        s[k + 'mlp.fc2.bias'] = (D,)
    s[b + 'norm.weight'] = (D,)
    s[b + 'norm.bias'] = (D,)
    s['seeker.tracker_post_linear.weight'] = (output_channels * P * P, D)
    s['seeker.tracker_post_linear.bias'] = (output_channels * P * P,)
    if flag_channels > 0:
        s['seeker.flag_post_linear.weight'] = (flag_channels, D)
        s['seeker.flag_post_linear.bias'] = (flag_channels,)
    return s


def make_state_dict(seed=901, **shape_kwargs):
    """Every tensor non-trivial: weights ~ tn(0.02)-like, biases N(0,0.02), LN weight 1+N(0,0.02)."""
    gen = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in state_dict_shapes(**shape_kwargs).items():
        r = torch.randn(shp, generator=gen, dtype=torch.float32)
        leaf = k.split('.')[-1]
        is_norm = ('norm' in k.split('.')[-2]) if '.' in k else False
        if k.endswith('patch_embed.proj.weight'):
            fan_in = shp[1] * shp[2] * shp[3]
            t = (torch.rand(shp, generator=gen) * 2 - 1) * fan_in ** -0.5
        elif k.startswith('seeker.tracker_post_linear') or k.startswith('seeker.flag_post_linear'):
            t = (torch.rand(shp, generator=gen) * 2 - 1) * shape_kwargs.get('embed_dim', 768) ** -0.5
        elif is_norm and leaf == 'weight':
            t = 1.0 + 0.02 * r
        elif len(shp) == 1:
            t = 0.02 * r
        else:
            t = (0.02 * r).clamp_(-0.04, 0.04)
        sd[k] = t.contiguous()
    return sd


def make_clip(sample, num_frames=30, frame_height=240, frame_width=320, query_frame=0):
    """One (clip, query) sample: rgb ~ U[0,1) (3,T,Hf,Wf); binary query rectangle (~2 % of
    the frame) at ``query_frame`` only, as data/data_utils.py:431 builds it."""
    g = torch.Generator().manual_seed(1000 + sample)
    rgb = torch.rand(3, num_frames, frame_height, frame_width, generator=g)
    h, w = max(frame_height // 6, 1), max(frame_width // 8, 1)
    y0 = int(torch.randint(0, frame_height - h + 1, (1,), generator=g))
    x0 = int(torch.randint(0, frame_width - w + 1, (1,), generator=g))
    q = torch.zeros(1, num_frames, frame_height, frame_width)
    q[0, query_frame, y0:y0 + h, x0:x0 + w] = 1.0
    return rgb, q


def make_batch(samples, **kw):
    clips = [make_clip(s, **kw) for s in samples]
    return torch.stack([c[0] for c in clips]), torch.stack([c[1] for c in clips])


def bench_clips(rank, batch, num_frames=30, frame_height=240, frame_width=320, queries_per_video=3):
    """The clips bench.py times on ``rank`` (BASELINE configs[1]): videos come with 3 queries each (README.md:42), so
    clips form triples sharing the RGB and differing in the query.  oracle/make_golden.py runs the unmodified
    reference on the same clips (tests/golden/full_bench_b8.npz) so the bench can check the batch it times."""
    rgb, q = [], []
    for i in range(batch):
        vid = (rank * batch + i) // queries_per_video
        rgb.append(make_clip(vid, num_frames, frame_height, frame_width)[0])
        q.append(make_clip(1000 + rank * batch + i, num_frames, frame_height, frame_width)[1])
    return torch.stack(rgb), torch.stack(q)


def make_targets(samples, num_frames=30, frame_height=240, frame_width=320, output_channels=3, flag_channels=3):
    """Seeded binary mask / flag targets for the training-step parity tests and benchmark
    (stand-in for the Kubric ground truth consumed by loss.py:164-225)."""
    masks, flags = [], []
    for s in samples:
        g = torch.Generator().manual_seed(5000 + s)
        m = torch.zeros(output_channels, num_frames, frame_height, frame_width)
        for c in range(output_channels):
            h, w = max(frame_height // (3 + c), 1), max(frame_width // (4 + c), 1)
            for t in range(num_frames):
                y0 = int(torch.randint(0, frame_height - h + 1, (1,), generator=g))
                x0 = int(torch.randint(0, frame_width - w + 1, (1,), generator=g))
                m[c, t, y0:y0 + h, x0:x0 + w] = 1.0
        masks.append(m)
        flags.append((torch.rand(num_frames, max(flag_channels, 1), generator=g) > 0.5).float())
    return torch.stack(masks), torch.stack(flags)[..., :flag_channels]


def training_loss(output_mask, output_flags, target_mask, target_flags):
    """Mean BCE-with-logits on the mask logits plus mean BCE on the flag logits: the differentiable core of
    loss.py:164-225 (the reference adds weighting / top-k bootstrapping on top of exactly these terms)."""
    import torch.nn.functional as F
    loss = F.binary_cross_entropy_with_logits(output_mask, target_mask)
    if output_flags is not None and output_flags.numel() > 0:
        loss = loss + F.binary_cross_entropy_with_logits(output_flags, target_flags)
    return loss
