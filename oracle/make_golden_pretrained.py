"""TEST INFRASTRUCTURE ONLY — pin tcow_b200.vit.load_pretrained against the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_pretrained
A seeded stand-in for the ImageNet ViT-B/16 file (same keys and shapes as jx_vit_base_p16_224; the real file cannot be
downloaded here) is written to a temporary path, the reference Seeker is constructed with
``tracker_pretrained=<that path>`` (model/mask_tracker.py:52-67 -> third_party/TimeSformer/timesformer/models/vit.py:462-464
-> helpers.py:100-202) and a digest of every backbone tensor it ends up with is stored in
tests/golden/pretrained_inflate.npz.  tests/test_host.py rebuilds the same file and requires the drop-in's state dict to
reproduce the digest exactly (the inflate only copies, tiles and rescales tensors).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'pretrained_inflate.npz')
KW = dict(num_total_frames=6, num_visible_frames=6, frame_height=64, frame_width=96, attention_type='divided_space_time',
          patch_size=16, causal_attention=1, norm_embeddings=False, drop_path_rate=0.1, network_depth=12,
          track_map_stride=4, track_map_resize='bilinear', query_channels=1, output_channels=3, flag_channels=3)
PREFIX = 'seeker.tracker_backbone.timesformer.model.'


def fake_imagenet_vit(seed=4242, linear_patch_proj=False):
    """Keys / shapes of the timm ViT-B/16 (224x224: 196 patches + cls) checkpoint the reference inflates from."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *shape: torch.randn(*shape, generator=g) * 0.02
    D = 768
    sd = {'cls_token': r(1, 1, D), 'pos_embed': r(1, 197, D),
          'patch_embed.proj.weight': r(D, 3 * 16 * 16) if linear_patch_proj else r(D, 3, 16, 16),
          'patch_embed.proj.bias': r(D)}
    for i in range(12):
        b = f'blocks.{i}.'
        sd[b + 'norm1.weight'], sd[b + 'norm1.bias'] = 1 + r(D), r(D)
        sd[b + 'attn.qkv.weight'], sd[b + 'attn.qkv.bias'] = r(3 * D, D), r(3 * D)
        sd[b + 'attn.proj.weight'], sd[b + 'attn.proj.bias'] = r(D, D), r(D)
        sd[b + 'norm2.weight'], sd[b + 'norm2.bias'] = 1 + r(D), r(D)
        sd[b + 'mlp.fc1.weight'], sd[b + 'mlp.fc1.bias'] = r(4 * D, D), r(4 * D)
        sd[b + 'mlp.fc2.weight'], sd[b + 'mlp.fc2.bias'] = r(D, 4 * D), r(D)
    sd['norm.weight'], sd['norm.bias'] = 1 + r(D), r(D)
    sd['head.weight'], sd['head.bias'] = r(1000, D), r(1000)
    return sd


def digest(t):
    """Order-sensitive fingerprint of a tensor: sum, weighted sum, first/last entries (float64)."""
    f = t.detach().double().flatten()
    w = torch.arange(1, f.numel() + 1, dtype=torch.float64) % 9973
    return np.array([f.sum().item(), (f * w).sum().item(), f[0].item(), f[-1].item(), float(f.numel())])


def backbone_digests(state_dict):
    return {k[len(PREFIX):]: digest(v) for k, v in state_dict.items() if k.startswith(PREFIX)}


if __name__ == '__main__':
    from oracle import ref_import
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'jx_vit_base_p16_224-fake.pth')
        torch.save(fake_imagenet_vit(), path)
        net = ref_import.build_reference(None, tracker_pretrained=path, **KW)
    assert net.seeker.tracker_backbone.pretrained is True
    d = backbone_digests(net.state_dict())
    arrs = {k.replace('.', '/'): v for k, v in d.items()}
    arrs['meta'] = np.frombuffer(json.dumps(dict(kwargs=KW, seed=4242, torch=torch.__version__)).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrs)
    print(f'{len(d)} backbone tensors digested -> {OUT}')
