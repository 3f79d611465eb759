"""Device input path.  CPU: the host geometry (tcow_b200/input_path.py) and the numpy oracle against the
fixtures produced by the reference's own loader code (oracle/make_golden_input.py).  GPU: the kernels against both."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import input_oracle
from tcow_b200 import input_path

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('input_') and f.endswith('.npz'))
TOL = 2e-6      # fp32 summation order of the separable filter (values are in [0, 1])


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return json.loads(bytes(z['meta']).decode()), z


def window_of(meta):
    return input_path.source_window(meta['H'], meta['W'], meta['Hf'], meta['Wf'], meta.get('center_crop', True),
                                    meta.get('crop_rect'), meta.get('flip', False))


def test_fixture_set_is_complete():
    assert set(CASES) >= {'input_down2', 'input_wide', 'input_tall_up', 'input_same', 'input_flip_crop'}


def test_center_crop_window_rules():
    assert input_path.center_crop_window(480, 640, 240, 320) == (0, 0, 480, 640)         # demo/teaduck2.mp4: nothing cropped
    assert input_path.center_crop_window(45, 100, 32, 48) == (0, 16, 45, 67)              # wider: int(H * ar), round((W-w)/2)
    assert input_path.center_crop_window(40, 30, 48, 64) == (9, 0, 22, 30)                # taller: int(W / ar)
    with pytest.raises(ValueError):
        input_path.source_window(60, 90, 32, 48, False, [0.5, 0.5, 0.1, 0.9])


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_loader(name):
    meta, z = load(name)
    win = window_of(meta)
    rgb = input_oracle.clip_from_video(z['video'], meta['start'], meta['stride'], meta['T'], meta['Hf'], meta['Wf'], win,
                                       meta.get('flip', False))
    assert rgb.shape == z['rgb'].shape
    assert np.abs(rgb - z['rgb']).max() <= TOL, np.abs(rgb - z['rgb']).max()
    if name == 'input_same':
        assert np.array_equal(rgb, z['rgb'])                   # identity resize: exactly uint8 / 255 in fp32
    m = input_oracle.clip_from_video(z['masks'][..., None], meta['start'], meta['stride'], meta['T'], meta['Hf'], meta['Wf'],
                                     win, meta.get('flip', False), nearest=True)
    assert np.array_equal(m, z['query_mask'])


def test_cpu_video_fails_loudly():
    with pytest.raises(RuntimeError, match='CUDA'):
        input_path.clip_from_video(torch.zeros(2, 8, 8, 3, dtype=torch.uint8), 0, 1, 2, 8, 8)


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_kernels_match_reference_loader(name):
    meta, z = load(name)
    video = torch.from_numpy(z['video']).cuda()
    masks = torch.from_numpy(z['masks']).cuda()
    kw = dict(center_crop=meta.get('center_crop', True), crop_rect=meta.get('crop_rect'), horz_flip=meta.get('flip', False))
    rgb = input_path.clip_from_video(video, meta['start'], meta['stride'], meta['T'], meta['Hf'], meta['Wf'], **kw)
    assert rgb.dtype == torch.float32 and tuple(rgb.shape) == z['rgb'].shape
    err = (rgb.cpu() - torch.from_numpy(z['rgb'])).abs().max().item()
    assert err <= TOL, err
    if name == 'input_same':
        assert torch.equal(rgb.cpu(), torch.from_numpy(z['rgb']))
    m = input_path.mask_clip_from_video(masks, meta['start'], meta['stride'], meta['T'], meta['Hf'], meta['Wf'], **kw)
    assert m.dtype == torch.uint8 and torch.equal(m.cpu(), torch.from_numpy(z['query_mask']))
    q = input_path.query_clip(masks[meta['start']], 1, meta['T'], meta['Hf'], meta['Wf'], **kw)
    assert torch.equal(q[:, 1].cpu(), torch.from_numpy(z['query_mask'])[:, 0]) and int(q[:, 0].sum()) == 0
    with pytest.raises(ValueError):
        input_path.clip_from_video(video, meta['F'], 1, meta['T'], meta['Hf'], meta['Wf'])      # frames outside the video


@pytest.mark.gpu
def test_uint8_division_is_exact_for_all_256_values():
    """`rgb / 255.0` in float64 then cast (data_plugin.py:174) == the kernel's fp32 division, for every byte value."""
    video = torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16, 1).cuda()
    out = input_path.clip_from_video(video, 0, 1, 1, 16, 16)
    want = torch.from_numpy((np.arange(256) / 255.0).astype(np.float32)).reshape(1, 1, 16, 16)
    assert torch.equal(out.cpu(), want)


@pytest.mark.gpu
def test_full_size_clip_feeds_the_seeker(logger):
    """480x640 uint8 video -> 240x320 clip on the device -> Seeker.forward, against the oracle-prepared clip."""
    import tcow_b200
    from conftest import cached_state_dict
    g = torch.Generator().manual_seed(5)
    video = torch.randint(0, 256, (40, 480, 640, 3), generator=g, dtype=torch.uint8)
    T, Hf, Wf = 30, 240, 320
    clip = input_path.clip_from_video(video.cuda(), 3, 1, T, Hf, Wf)
    ref = input_oracle.clip_from_video(video.numpy()[:, ::1], 3, 1, 2, Hf, Wf, (0, 0, 480, 640))     # first two frames
    assert (clip[:, :2].cpu() - torch.from_numpy(ref)).abs().max().item() <= TOL
    net = tcow_b200.Seeker(logger, num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                           tracker_pretrained=False, causal_attention=1)
    net.load_state_dict(cached_state_dict(901, T, Hf, Wf))
    net = net.cuda().eval()
    q = torch.zeros(1, 1, T, Hf, Wf, dtype=torch.uint8, device='cuda')
    q[0, 0, 0, 40:80, 60:100] = 1
    with torch.no_grad():
        mask, flags = net(clip[None], q)
    assert torch.isfinite(mask).all() and tuple(mask.shape) == (1, 3, T, Hf, Wf)
