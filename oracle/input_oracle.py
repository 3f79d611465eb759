"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's host-side clip preparation for the device input path.

Follows data/data_plugin.py:156-176,199-200 (frame selection, `rgb / 255.0`, channel-first) and data/augs.py:166-206
(centre crop, flip, crop rectangle, torchvision Resize) with the resize written out as ATen computes it
(UpSampleKernel.cpp: separable anti-aliased triangle filter, align_corners=False, horizontal pass then vertical pass;
legacy 'nearest' for masks).  Pinned against the reference's own pipeline by tests/golden/input_*.npz
(oracle/make_golden_input.py); the CUDA kernels of csrc/input_path.cu are tested against both.
"""
from __future__ import annotations

import numpy as np


def _aa_weights(in_size, out_size):
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = scale if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    taps = []
    for i in range(out_size):
        center = f32(float(scale) * (i + 0.5))
        xmin = max(int(float(f32(center - support)) + 0.5), 0)
        xsize = min(int(float(f32(center + support)) + 0.5), in_size) - xmin
        w = np.zeros(max(xsize, 0), dtype=f32)
        for j in range(xsize):
            x = f32((float(f32(f32(j + xmin) - center)) + 0.5) * float(invscale))
            w[j] = max(f32(0.0), f32(1.0) - abs(x))
        tot = f32(w.sum(dtype=f32)) if xsize > 0 else f32(0)
        if tot != 0:
            w = (w / tot).astype(f32)
        taps.append((xmin, w))
    return taps


def resize_bilinear_aa(img, out_h, out_w):
    """img (..., h, w) float32 -> (..., out_h, out_w): torch interpolate(mode='bilinear', antialias=True)."""
    img = img.astype(np.float32)
    h, w = img.shape[-2:]
    tx, ty = _aa_weights(w, out_w), _aa_weights(h, out_h)
    tmp = np.zeros(img.shape[:-1] + (out_w,), dtype=np.float32)
    for ox, (x0, wt) in enumerate(tx):
        acc = np.zeros(img.shape[:-1], dtype=np.float32)
        for j, wj in enumerate(wt):
            acc = (acc + wj * img[..., x0 + j]).astype(np.float32)
        tmp[..., ox] = acc
    out = np.zeros(img.shape[:-2] + (out_h, out_w), dtype=np.float32)
    for oy, (y0, wt) in enumerate(ty):
        acc = np.zeros(tmp.shape[:-2] + (out_w,), dtype=np.float32)
        for j, wj in enumerate(wt):
            acc = (acc + wj * tmp[..., y0 + j, :]).astype(np.float32)
        out[..., oy, :] = acc
    return out


def resize_nearest(img, out_h, out_w):
    h, w = img.shape[-2:]

    def idx(n_in, n_out):
        if n_in == n_out:
            return np.arange(n_out)
        scale = np.float32(n_in) / np.float32(n_out)
        return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64), n_in - 1)
    return img[..., idx(h, out_h)[:, None], idx(w, out_w)[None, :]]


def clip_from_video(video, frame_start, frame_stride, num_frames, frame_height, frame_width, window, flip=False,
                    nearest=False):
    """video (F,H,W,C) uint8; window = (y0, x0, h, w) as tcow_b200.input_path.source_window computes it."""
    y0, x0, h, w = window
    inds = [frame_start + t * frame_stride for t in range(num_frames)]
    frames = video[inds][:, y0:y0 + h, x0:x0 + w]                       # (T,h,w,C)
    if flip:
        frames = frames[:, :, ::-1]
    chw = np.transpose(frames, (3, 0, 1, 2))                             # (C,T,h,w)
    if nearest:
        return resize_nearest(chw, frame_height, frame_width).astype(np.uint8)
    return resize_bilinear_aa((chw / 255.0).astype(np.float32), frame_height, frame_width)
