"""TEST INFRASTRUCTURE ONLY — pin the device input path (tcow_b200/input_path.py, csrc/input_path.cu) against the
UNMODIFIED reference loader code.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_input
Seeded synthetic uint8 videos go through exactly what data/data_plugin.py:168-205 does per sample — `rgb / 255.0`,
'T H W C -> C T H W', then the reference's own MyAugmentationPipeline.apply_augs_2d_frames (data/augs.py:138-210: centre
crop, flip, crop rectangle, torchvision Resize antialiased-bilinear / nearest) — and the resulting clip tensors are stored
in tests/golden/input_*.npz with the raw video (tiny sizes).
"""
from __future__ import annotations

import json
import logging
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    # the demo's geometry in miniature: 4:3 source, exact 2x downscale, no crop needed
    dict(name='input_down2', F=7, H=48, W=64, Hf=24, Wf=32, start=1, stride=2, T=3),
    # wider than the target: centre crop of the columns, non-integer downscale
    dict(name='input_wide', F=5, H=45, W=100, Hf=32, Wf=48, start=0, stride=1, T=4),
    # taller than the target: centre crop of the rows, upscale
    dict(name='input_tall_up', F=4, H=40, W=30, Hf=48, Wf=64, start=3, stride=-1, T=3),
    # already the right size: identity resize (bit-exact /255)
    dict(name='input_same', F=3, H=32, W=48, Hf=32, Wf=48, start=0, stride=1, T=3),
    # train-time 2D augmentation: horizontal flip + crop rectangle (data/augs.py:189-196)
    dict(name='input_flip_crop', F=4, H=60, W=90, Hf=32, Wf=48, start=0, stride=1, T=4, flip=True,
         crop_rect=[0.07, 0.93, 0.11, 0.88], center_crop=False),
]


def make_video(c, seed=77):
    g = torch.Generator().manual_seed(seed + c['F'] * 1000 + c['H'])
    # smooth-ish content plus noise so that interpolation errors show
    base = torch.rand(c['F'], c['H'] // 4 + 1, c['W'] // 4 + 1, 3, generator=g)
    up = torch.nn.functional.interpolate(base.permute(0, 3, 1, 2), size=(c['H'], c['W']), mode='bilinear').permute(0, 2, 3, 1)
    rgb = (up * 200 + torch.rand(c['F'], c['H'], c['W'], 3, generator=g) * 55).round().clamp(0, 255).to(torch.uint8)
    mask = (torch.rand(c['F'], c['H'] // 6 + 1, c['W'] // 6 + 1, generator=g) > 0.6).to(torch.uint8)
    mask = mask.repeat_interleave(6, 1).repeat_interleave(6, 2)[:, :c['H'], :c['W']].contiguous()
    return rgb.contiguous(), mask


def run_case(c, augs_mod):
    rgb, mask = make_video(c)
    T = c['T']
    pipe = augs_mod.MyAugmentationPipeline(logging.getLogger('x'), T, T, c['Hf'], c['Wf'], 1, False, False, 0.0, 0.0,
                                           c.get('center_crop', True))
    params = pipe.sample_augs_params()
    params['horz_flip'] = bool(c.get('flip', False))
    if 'crop_rect' in c:
        params['crop_rect'] = np.array(c['crop_rect'])
    inds = [c['start'] + t * c['stride'] for t in range(T)]
    # data/data_plugin.py:168-176, 199-200
    pv_rgb = np.stack([(rgb[t].numpy() / 255.0).astype(np.float32) for t in inds], 0)
    pv_mask = np.stack([mask[t].numpy()[..., None] for t in inds], 0)
    mods = {'rgb': torch.tensor(pv_rgb).permute(3, 0, 1, 2), 'query_mask': torch.tensor(pv_mask).permute(3, 0, 1, 2)}
    out = pipe.apply_augs_2d_frames(mods, params)
    meta = dict(c)
    meta['torch'] = torch.__version__
    np.savez_compressed(os.path.join(OUT, c['name'] + '.npz'), video=rgb.numpy(), masks=mask.numpy(),
                        rgb=out['rgb'].numpy().astype(np.float32), query_mask=out['query_mask'].numpy().astype(np.uint8),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print(f"{c['name']:18s} video {tuple(rgb.shape)} -> rgb {tuple(out['rgb'].shape)} mask {tuple(out['query_mask'].shape)}")


if __name__ == '__main__':
    ref_import.import_reference_seeker()          # registers the stub modules the reference's `from __init__ import *` needs
    cwd = os.getcwd()
    os.chdir(ref_import.REF)
    sys.path[:0] = [ref_import.REF, ref_import.REF + '/data', ref_import.REF + '/utils']
    try:
        import augs
    finally:
        os.chdir(cwd)
    os.makedirs(OUT, exist_ok=True)
    for c in CASES:
        run_case(c, augs)
