"""TEST INFRASTRUCTURE ONLY — plain-torch restatement of the reference's mask-tracking objective and metrics.

Follows loss.py:12-16 (bootstrap_warmup_loss), :19-31 (tversky_loss), :50-53 (bce_or_focal_loss), :55-81 (frame weights),
:83-148 (pixel weights), :164-225 (my_mask_loss), :236-318 + :352-354 (per-example terms and total) and
eval/metrics.py:9-113.  Pinned against the UNMODIFIED reference by tests/golden/loss_*.npz (oracle/make_golden_loss.py);
device-agnostic so the GPU tests can run it at sizes the fixtures do not cover.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def focal(x, y, alpha=0.25, gamma=2.0):                       # torchvision.ops.sigmoid_focal_loss(reduction='none')
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, y, reduction='none')
    p_t = p * y + (1 - p) * (1 - y)
    return (alpha * y + (1 - alpha) * (1 - y)) * ce * (1 - p_t) ** gamma


def tversky(x, y, alpha=1.0, beta=1.0, eps=0.1):
    if y.mean() >= 1e-6:
        p = torch.sigmoid(x)
        num = (p * y).sum()
        den = num + alpha * (p * (1 - y)).sum() + beta * ((1 - p) * y).sum()
        return 1.0 - num / (den + eps)
    return torch.zeros((), device=x.device)


def my_mask_loss(x, y, w, progress, apply_weights_for_aot, aot_loss=0.8, focal_loss=False):
    which = (w != 0).flatten(3).any(-1)                                        # (B,Q,T)
    which_full = which[..., None, None].expand_as(w)
    if not (which.any() and w.mean() >= 1e-4):
        return torch.zeros((), device=x.device)
    xs, ys, ws = x[which_full], y[which_full], w[which_full]
    l = focal(xs, ys) if focal_loss else F.binary_cross_entropy_with_logits(xs, ys, reduction='none')
    custom = (l * ws).mean()
    if aot_loss > 0.0:
        la = l * ws if apply_weights_for_aot else l
        frac = min(max(1.0 - progress * 8.5, 0.15), 1.0)
        boot = torch.topk(la.flatten(), k=int(frac * la.numel()))[0].mean()
        jac = boot if apply_weights_for_aot else tversky(xs, ys)
        loss = (boot + jac) / 2.0 * aot_loss + custom * (1.0 - aot_loss)
    else:
        loss = custom
    return loss * torch.sqrt(which_full.float().mean())


def frame_weights(sel_occl_fracs, query_time, occluded_weight=5.0):
    fw = (sel_occl_fracs[..., 0].float() * occluded_weight).clip(min=1.0)
    fw[-1, :, query_time] *= 0.2
    return fw


def pixel_weights(target, occl_ptr, class_balancing=True, hard_negative_factor=3.0):
    B, Q, T, H, W = target.shape
    pw = torch.ones_like(target, dtype=torch.float32)
    if class_balancing:
        pos, neg = target == 1.0, target == 0.0
        pf = (pos.sum() / pos.numel()).clip(min=0.05).item()
        nf = (neg.sum() / neg.numel()).clip(min=0.05).item()
        if pf > nf:
            pc, nc = np.power(nf / pf, 0.7), np.power(nf / pf, -0.3)
        else:
            pc, nc = np.power(pf / nf, -0.3), np.power(pf / nf, 0.7)
        pw[neg] *= nc
        pw[pos] *= pc
    pw[occl_ptr != 0] *= 2.0
    if hard_negative_factor > 1.0:
        band = int(np.sqrt(H * W) / 12.0)
        band += 1 - band % 2
        r = band // 2
        t = F.pad((target > 0).float().reshape(-1, 1, H, W), (r, r, r, r), mode='reflect')
        near = F.max_pool2d(t, band, stride=1).reshape(B, Q, T, H, W) > 0          # what gaussian_blur(...) > 0 selects
        near[target >= 0.5] = False
        pw[near] *= hard_negative_factor
    return pw


def seeker_loss(output_mask, target_mask, sel_occl_fracs, occl_ptr, query_time, progress, track_lw=1.0, occl_mask_lw=0.5,
                cont_mask_lw=0.5, occluded_weight=5.0, occl_cont_zero_weight=0.02, **kw):
    fw = frame_weights(sel_occl_fracs, query_time, occluded_weight)
    w0 = fw[..., None, None] * pixel_weights(target_mask[:, :, 0], occl_ptr[:, :, 0],
                                             kw.get('class_balancing', True), kw.get('hard_negative_factor', 3.0))
    lk = dict(aot_loss=kw.get('aot_loss', 0.8), focal_loss=kw.get('focal_loss', False))
    terms = {'track': my_mask_loss(output_mask[:, :, 0], target_mask[:, :, 0], w0, progress, False, **lk)}
    for key, ch in (('occl_mask', 1), ('cont_mask', 2)):
        present = target_mask[:, :, ch].any(-1).any(-1)[..., None, None].expand_as(target_mask[:, :, ch]).float()
        w = present * (1.0 - occl_cont_zero_weight) + occl_cont_zero_weight
        terms[key] = my_mask_loss(output_mask[:, :, ch], target_mask[:, :, ch], w, progress, True, **lk)
    total = terms['track'] * track_lw + terms['occl_mask'] * occl_mask_lw + terms['cont_mask'] * cont_mask_lw
    return total, terms


def metrics(output_mask, target_mask):
    """eval/metrics.py:17-100 for (B,Q,3,T,H,W) tensors."""
    o, t = output_mask > 0.0, target_mask > 0.5
    ta = t.sum((-1, -2)).cpu().numpy()
    ia = (o & t).sum((-1, -2)).cpu().numpy()
    ua = (o | t).sum((-1, -2)).cpu().numpy()
    B, Q, C, T = ta.shape
    lists = {k: [] for k in ('snitch', 'occl_mask', 'cont_mask', 'snitch_during_vis', 'snitch_during_occl', 'snitch_during_cont')}
    for b in range(B):
        for q in range(Q):
            for f in range(T):
                iou = [ia[b, q, c, f] / (ua[b, q, c, f] + 1e-7) for c in range(C)]
                if ta[b, q, 0, f] > 0:
                    lists['snitch'].append(iou[0])
                    if ta[b, q, 1, f] == 0:
                        lists['snitch_during_vis'].append(iou[0])
                    if ta[b, q, 1, f] > 0:
                        lists['snitch_during_occl'].append(iou[0])
                    if ta[b, q, 2, f] > 0:
                        lists['snitch_during_cont'].append(iou[0])
                if ta[b, q, 1, f] > 0:
                    lists['occl_mask'].append(iou[1])
                if ta[b, q, 2, f] > 0:
                    lists['cont_mask'].append(iou[2])
    out = {}
    for k, v in lists.items():
        out[f'mean_{k}_iou'] = float(np.mean(v)) if v else -1.0
        out[f'count_{k}_iou'] = len(v)
    return out
