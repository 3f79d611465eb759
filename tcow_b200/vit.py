"""Parameter containers for the TimeSformer-style divided space-time ViT of the Seeker backbone.

These modules exist to (a) own ``nn.Parameter``s under exactly the reference's state-dict names
(third_party/TimeSformer/timesformer/models/vit.py:126-153 Block, :220-233 PatchEmbed, :244-297
VisionTransformer, :416-430 TimeSformer — 251 tensors in total) so that
``checkpoint['net_seeker']`` round-trips, and (b) reproduce the reference initialisation.  They do NOT
implement a forward: all arithmetic runs in the CUDA engine (``tcow_b200/engine.py``).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# File name under which torch.hub caches the ImageNet ViT-B/16 the reference downloads (vit.py:33-36); when that file is
# already in the hub cache the reference constructs without touching the network, and so does this module.
IMAGENET_VIT_FILE = 'jx_vit_base_p16_224-80ecf9dd.pth'


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
        self.pretrained = pretrained
        self.attention_type = attention_type
        self.drop_path_rate = drop_path_rate
        self.model = VisionTransformer(img_size, patch_size, in_chans, 768, 12, 12, 4, num_frames, causal_attention)
        self.num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        if pretrained:                                                     # vit.py:462-464
            load_pretrained(self.model, find_pretrained_file(pretrained_model), in_chans=in_chans, patch_size=patch_size,
                            num_frames=num_frames, num_patches=self.num_patches)


# ------------------------------------------------------------------------------------------------ pretrained weights
def find_pretrained_file(pretrained_model=''):
    """Local file holding the image-ViT (or TimeSformer) weights to start from.  The reference fetches
    jx_vit_base_p16_224 over the network when no path is given (helpers.py:109-110); there is no network path here, so
    an explicit path, $TCOW_PRETRAINED_VIT or a copy already sitting in the torch hub cache is required."""
    if pretrained_model:
        if not os.path.isfile(pretrained_model):
            raise FileNotFoundError(pretrained_model)                      # helpers.py:96-97
        return pretrained_model
    candidates = [os.environ.get('TCOW_PRETRAINED_VIT', ''),
                  os.path.join(torch.hub.get_dir(), 'checkpoints', IMAGENET_VIT_FILE)]
    for c in candidates:
        if c and os.path.isfile(c):
            return c
    raise RuntimeError(
        "tracker_pretrained is truthy but no pretrained ViT file is available offline: pass its path as "
        f"tracker_pretrained='/path/to/{IMAGENET_VIT_FILE}', set $TCOW_PRETRAINED_VIT, or place the file in "
        f"{candidates[1]!r}.  To evaluate a TCOW checkpoint no ImageNet weights are needed: "
        "tcow_b200.checkpoint.build_seeker constructs without them and restores the RGB-normalisation flag.")


def _read_weights(path):
    """The tensors of a checkpoint file, whatever wrapper they sit in (helpers.py:22-47, :112-114)."""
    ckpt = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(ckpt, dict) and 'state_dict' in ckpt:
        sd = OrderedDict((k[7:] if k.startswith('module') else k, v) for k, v in ckpt['state_dict'].items())
    elif isinstance(ckpt, dict) and 'model_state' in ckpt:
        sd = OrderedDict((k[6:] if k.startswith('model') else k, v) for k, v in ckpt['model_state'].items())
    else:
        sd = ckpt
    return sd['model'] if 'model' in sd else sd


@torch.no_grad()
def load_pretrained(model, path, in_chans=4, patch_size=16, num_frames=30, num_patches=300):
    """Inflate image-ViT weights into the divided space-time model the way helpers.py:100-202 does:
    patch projection reshaped to a conv kernel and its 3 input channels tiled to `in_chans` (scaled by 3/in_chans);
    classifier dropped; positional / temporal embeddings resampled (nearest) to this model's grid and length;
    every block's temporal attention and temporal_norm1 initialised from the spatial ones unless the file has them."""
    sd = OrderedDict(_read_weights(path))
    key = 'patch_embed.proj.weight'
    w = sd[key]
    if w.dim() != 4:                                                       # vit.py:381-390: stored as a linear projection
        w = w.reshape(w.shape[0], 3, patch_size, patch_size)
    if in_chans != 3:
        if w.shape[1] != 3:
            del sd[key]                                                    # helpers.py:140-143: unusable first conv
            w = None
        elif in_chans == 1:
            w = w.float().sum(1, keepdim=True).to(w.dtype)                 # helpers.py:119-133
        else:
            rep = int(math.ceil(in_chans / 3))
            w = (w.float().repeat(1, rep, 1, 1)[:, :in_chans] * (3.0 / in_chans)).to(w.dtype)   # helpers.py:144-150
    if w is not None:
        sd[key] = w
    sd.pop('head.weight', None)                                            # num_classes=0 here (vision_tf.py:59)
    sd.pop('head.bias', None)
    pos = sd['pos_embed']
    if pos.shape[1] != num_patches + 1:                                    # helpers.py:170-178
        grid = F.interpolate(pos[:, 1:].transpose(1, 2), size=num_patches, mode='nearest').transpose(1, 2)
        sd['pos_embed'] = torch.cat([pos[:, :1], grid], 1)
    if 'time_embed' in sd and sd['time_embed'].shape[1] != num_frames:     # helpers.py:180-184
        sd['time_embed'] = F.interpolate(sd['time_embed'].transpose(1, 2), size=num_frames, mode='nearest').transpose(1, 2)
    for k in list(sd):                                                     # helpers.py:186-202
        if 'blocks' in k and 'attn' in k:
            sd.setdefault(k.replace('attn', 'temporal_attn'), sd[k])
        if 'blocks' in k and 'norm1' in k:
            sd.setdefault(k.replace('norm1', 'temporal_norm1'), sd[k])
    return model.load_state_dict(sd, strict=False)
