"""Parameter containers for the TimeSformer-style divided space-time ViT of the Seeker backbone.

These modules exist to (a) own ``nn.Parameter``s under exactly the reference's state-dict names
(third_party/TimeSformer/timesformer/models/vit.py:126-153 Block, :220-233 PatchEmbed, :244-297
VisionTransformer, :416-430 TimeSformer — 251 tensors in total) so that
``checkpoint['net_seeker']`` round-trips, and (b) reproduce the reference initialisation.  They do NOT
implement a forward: all arithmetic runs in the CUDA engine (``tcow_b200/engine.py``).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard against accidental eager use
        raise RuntimeError(f'{type(self).__name__} is a parameter container; the forward runs in '
                           'tcow_b200.engine (CUDA, sm_100a) — there is no eager/CPU path')


class Mlp(_NoForward):  # vit.py:45-61
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Attention(_NoForward):  # vit.py:64-76
    def __init__(self, dim, num_heads, causal_attention=0):
        super().__init__()
        self.num_heads = num_heads
        self.causal_attention = causal_attention
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Block(_NoForward):  # vit.py:126-153 (attention_type == 'divided_space_time')
    def __init__(self, dim, num_heads, mlp_ratio, causal_attention, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, num_heads, 0)  # the spatial module is never causal (vit.py:136-138)
        self.temporal_norm1 = nn.LayerNorm(dim, eps=eps)
        self.temporal_attn = Attention(dim, num_heads, causal_attention)
        self.temporal_fc = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(_NoForward):  # vit.py:220-233
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size = tuple(img_size)
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size[1] // patch_size) * (img_size[0] // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def _trunc_normal_(t, std):
    # same distribution as vit_utils.py:58-76 (mean 0, cut at +-2 in absolute units)
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class VisionTransformer(_NoForward):  # vit.py:244-306
    def __init__(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, num_frames,
                 causal_attention, eps=1e-6):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.depth = depth
        self.num_heads = num_heads
        self.causal_attention = causal_attention
        self.attention_type = 'divided_space_time'
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.time_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))  # stays zero at init (vit.py:268)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, causal_attention, eps)
                                     for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=eps)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        _trunc_normal_(self.pos_embed, 0.02)
        _trunc_normal_(self.cls_token, 0.02)
        for m in self.modules():  # vit.py:299-306 (Conv2d keeps the torch default init)
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, 0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        # vit.py:289-297 zeroes temporal_fc of every block: its loop counts the ModuleList itself as the
        # first "Block", so block 0 is zeroed as well.
        for blk in self.blocks:
            nn.init.zeros_(blk.temporal_fc.weight)
            nn.init.zeros_(blk.temporal_fc.bias)


class TimeSformer(_NoForward):  # vit.py:416-468
    def __init__(self, img_size, patch_size=16, num_frames=8, attention_type='divided_space_time',
                 causal_attention=0, drop_path_rate=0.1, network_depth=12, pretrained=False,
                 pretrained_model='', in_chans=3):
        super().__init__()
        assert attention_type in ['divided_space_time', 'space_only', 'joint_space_time']  # vit.py:133
        if attention_type != 'divided_space_time':
            raise NotImplementedError('tcow_b200 implements attention_type="divided_space_time" only '
                                      '(the only mode TCOW trains and evaluates, args.py:154)')
        if network_depth != 12:
            if network_depth in (18, 24):
                raise NotImplementedError('tcow_b200 implements network_depth=12 (ViT-B/16) only')
            raise ValueError(f'Invalid network depth {network_depth}, must be one of 12, 18, 24.')  # vit.py:449
        if pretrained:
            # vit.py:462-464 downloads ImageNet ViT weights; there is no network path here.
            raise NotImplementedError('tracker_pretrained=True needs a download; construct with False and '
                                      'load_state_dict a checkpoint (set backbone.pretrained=True for the '
                                      'RGB normalisation of vision_tf.py:81-89)')
        self.pretrained = pretrained
        self.attention_type = attention_type
        self.drop_path_rate = drop_path_rate
        self.model = VisionTransformer(img_size, patch_size, in_chans, 768, 12, 12, 4, num_frames, causal_attention)
        self.num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
