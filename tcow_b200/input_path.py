"""Input path on the device: from a decoded uint8 video to the tensors Seeker.forward consumes.

The reference does this on the host per sample — data/data_plugin.py:152-233 (frame selection, `rgb / 255.0`, masks,
'T H W C -> C T H W') and data/augs.py:138-210 (centre crop to the target aspect ratio, optional flip / crop rectangle,
torchvision Resize) — and then ships 36.9 MB of fp32 per clip over PCIe.  Here the uint8 video crosses PCIe once
(0.9 MB per 480x640 frame) and every clip of the sweep is cut out of it on the GPU by one kernel
(csrc/input_path.cu); the geometry below restates the reference's rules so the result is what its loader would yield.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def center_crop_window(H, W, frame_height, frame_width):
    """(y0, x0, h, w) kept by the test-time centre crop of data/augs.py:170-178 (torchvision CenterCrop rounding)."""
    current_ar, desired_ar = W / H, frame_width / frame_height
    h, w = H, W
    if current_ar > desired_ar:
        w = int(H * desired_ar)
    elif current_ar < desired_ar:
        h = int(W / desired_ar)
    return int(round((H - h) / 2.0)), int(round((W - w) / 2.0)), h, w


def source_window(H, W, frame_height, frame_width, center_crop=True, crop_rect=None, horz_flip=False):
    """Window of the raw frame that ends up in the clip, as (y0, x0, h, w) for the kernels (flip handled there).
    crop_rect = (y1, y2, x1, x2) fractions in post-flip coordinates, scaled by the RAW frame size and applied to the
    centre-cropped frame exactly as data/augs.py:192-196 slices it (python slicing clips at the window's edge)."""
    cy0, cx0, ch, cw = center_crop_window(H, W, frame_height, frame_width) if center_crop else (0, 0, H, W)
    a_y, b_y, a_x, b_x = 0, ch, 0, cw
    if crop_rect is not None and all(float(v) >= 0.0 for v in crop_rect):
        y1, y2, x1, x2 = (float(v) for v in crop_rect)
        a_y, b_y = min(int(y1 * H), ch), min(int(y2 * H), ch)
        a_x, b_x = min(int(x1 * W), cw), min(int(x2 * W), cw)
    h, w = b_y - a_y, b_x - a_x
    if h <= 0 or w <= 0:
        raise ValueError(f'empty crop window from crop_rect={crop_rect}')
    x0 = cx0 + cw - a_x - w if horz_flip else cx0 + a_x      # post-flip column u is raw column cx0 + cw - 1 - u
    return cy0 + a_y, x0, h, w


def _check_video(video, name):
    if not video.is_cuda:
        raise RuntimeError(f'{name}: the video must live on a CUDA device (tcow_b200 has no CPU path)')
    if video.dtype != torch.uint8:
        raise TypeError(f'{name}: expected a uint8 video as a decoder leaves it, got {video.dtype}')
    if video.dim() == 3:
        video = video[..., None]
    if video.dim() != 4 or not video.is_contiguous():
        raise ValueError(f'{name}: expected a contiguous (F, H, W, C) tensor')
    return video


def clip_from_video(video, frame_start, frame_stride, num_frames, frame_height, frame_width, center_crop=True,
                    crop_rect=None, horz_flip=False):
    """video (F,H,W,3) uint8 on the GPU -> (3, T, Hf, Wf) fp32 in [0,1]: `pv_rgb_tf` of data/data_plugin.py:229."""
    video = _check_video(video, 'clip_from_video')
    F, H, W, C = video.shape
    y0, x0, h, w = source_window(H, W, frame_height, frame_width, center_crop, crop_rect, horz_flip)
    out = torch.empty((C, num_frames, frame_height, frame_width), device=video.device, dtype=torch.float32)
    _lib.call('tcow_clip_from_video_u8', video.data_ptr(), F, H, W, C, int(frame_start), int(frame_stride), int(num_frames),
              y0, x0, h, w, int(bool(horz_flip)), int(frame_height), int(frame_width), out.data_ptr(), _stream())
    return out


def mask_clip_from_video(masks, frame_start, frame_stride, num_frames, frame_height, frame_width, center_crop=True,
                         crop_rect=None, horz_flip=False):
    """masks (F,H,W) or (F,H,W,C) uint8 on the GPU -> (C, T, Hf, Wf) uint8, nearest resize: the query / target mask
    modalities of data/data_plugin.py:178-200 ('mask' in the modality name selects nearest, data/augs.py:199-202)."""
    masks = _check_video(masks, 'mask_clip_from_video')
    F, H, W, C = masks.shape
    y0, x0, h, w = source_window(H, W, frame_height, frame_width, center_crop, crop_rect, horz_flip)
    out = torch.empty((C, num_frames, frame_height, frame_width), device=masks.device, dtype=torch.uint8)
    _lib.call('tcow_mask_clip_from_video_u8', masks.data_ptr(), F, H, W, C, int(frame_start), int(frame_stride),
              int(num_frames), y0, x0, h, w, int(bool(horz_flip)), int(frame_height), int(frame_width), out.data_ptr(),
              _stream())
    return out


def query_clip(query_frame_mask, query_time, num_frames, frame_height, frame_width, center_crop=True, crop_rect=None,
               horz_flip=False):
    """One annotated query frame (H, W) uint8 -> the (1, T, Hf, Wf) uint8 query mask that is zero except at clip frame
    `query_time` (data/data_plugin.py:178-181, data/data_utils.py:431)."""
    m = mask_clip_from_video(query_frame_mask[None].contiguous(), 0, 1, 1, frame_height, frame_width, center_crop, crop_rect,
                             horz_flip)
    out = torch.zeros((1, num_frames, frame_height, frame_width), device=m.device, dtype=torch.uint8)
    out[:, query_time] = m[:, 0]
    return out
