"""Mask-loss terms of the training step on the device: the weighted BCE of loss.py:164-185 and the
soft-Jaccard / Tversky term of loss.py:19-31 from TWO bandwidth passes over the logits (sums, then the gradient) instead
of the ~25 elementwise kernels torch needs for the same expressions and their autograd (`mask_loss_terms`), and — second
half of this file — the reference's complete objective: per-pixel weights (loss.py:83-148), `my_mask_loss` with frame
selection, focal option, bootstrapped top-k BCE as an exact on-device radix select and the soft-Jaccard term
(loss.py:164-225), the flag loss (:150-162), the weighted total (:236-354) and the IoU metrics of eval/metrics.py:9-113
(csrc/loss_full.cu; no host synchronisation anywhere)."""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _MaskLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weights, alpha, beta, eps):
        if not logits.is_cuda:
            raise RuntimeError('tcow_b200.loss runs on a CUDA sm_100 device only (no CPU fallback)')
        x = logits.contiguous().float()
        y = target.contiguous().float()
        w = None if weights is None else weights.expand_as(logits).contiguous().float()
        n = x.numel()
        if n % 4:
            raise ValueError('mask_loss: the number of elements must be a multiple of 4')
        ws = torch.empty(int(_lib.load().tcow_mask_loss_workspace_floats()), device=x.device, dtype=torch.float32)
        sums = torch.empty(5, device=x.device, dtype=torch.float64)
        _lib.call('tcow_mask_loss_sums', x.data_ptr(), y.data_ptr(), 0 if w is None else w.data_ptr(), n, ws.data_ptr(),
                  sums.data_ptr(), _stream())
        bce = sums[0] / n
        num = sums[1]
        den = num + alpha * sums[2] + beta * sums[3]
        has_target = (sums[4] / n) >= 1e-6                          # loss.py:20: no target -> the Jaccard term is 0
        tversky = torch.where(has_target, 1.0 - num / (den + eps), torch.zeros_like(num))
        ctx.save_for_backward(x, y, w if w is not None else x.new_empty(0), sums)
        ctx.consts = (alpha, beta, eps, n, w is not None)
        return bce.float(), tversky.float()

    @staticmethod
    def backward(ctx, g_bce, g_tv):
        x, y, w, sums = ctx.saved_tensors
        alpha, beta, eps, n, has_w = ctx.consts
        num = sums[1]
        den = num + alpha * sums[2] + beta * sums[3] + eps
        has_target = ((sums[4] / n) >= 1e-6).to(torch.float64)
        # L = 1 - num/den;  d num/dx = p(1-p) y;  d den/dx = p(1-p) (y + alpha (1-y) - beta y)
        # dL/dx = p(1-p) [ -y/den + num/den^2 (alpha + (1 - alpha - beta) y) ]
        gt = g_tv.to(torch.float64) * has_target
        c1 = gt * (-1.0 / den + num / (den * den) * (1.0 - alpha - beta))
        c2 = gt * (num / (den * den) * alpha)
        coef = torch.stack([g_bce.to(torch.float64) / n, c1, c2]).float().contiguous()
        grad = torch.empty_like(x)
        _lib.call('tcow_mask_loss_grad', x.data_ptr(), y.data_ptr(), w.data_ptr() if has_w else 0, n, coef.data_ptr(),
                  grad.data_ptr(), _stream())
        return grad, None, None, None, None, None


def mask_loss_terms(output_mask_logits, target_mask, final_weights=None, alpha=1.0, beta=1.0, eps=0.1):
    """(mean(weights * bce_with_logits(x, y)), tversky_loss(x, y, alpha, beta, eps)) — loss.py:181-183 and :19-31."""
    return _MaskLoss.apply(output_mask_logits, target_mask, final_weights, float(alpha), float(beta), float(eps))


# ------------------------------------------------------------------------------------------------ the complete mask loss
def _grouped(t, name):
    """(B,Q,T,H,W) view -> (tensor, group_stride): G = B*Q groups of T contiguous frames, addressed in place when the
    strides allow it (a channel slice `mask[:, :, c]` of a contiguous (B,Q,3,T,H,W) tensor does)."""
    if t.dim() != 5:
        raise ValueError(f'{name}: expected a (B, Q, T, H, W) tensor, got {tuple(t.shape)}')
    B, Q, T, H, W = t.shape

    def stride_of(u):
        return u.stride(1) if Q > 1 else (u.stride(0) if B > 1 else T * H * W)
    ok = (t.dtype == torch.float32 and t.stride(4) == 1 and t.stride(3) == W and t.stride(2) == H * W
          and (B == 1 or Q == 1 or t.stride(0) == Q * t.stride(1)) and stride_of(t) % 4 == 0 and t.data_ptr() % 16 == 0)
    if not ok:
        t = t.float().contiguous()
    return t, stride_of(t)


def topk_fraction(progress):
    """Share of the pixels kept by the bootstrapped BCE at training progress in [0,1] (loss.py:197)."""
    return min(max(1.0 - progress * 8.5, 0.15), 1.0)


class _MyMaskLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weights, topk_frac, weighted_aot, aot_loss, focal):
        if not logits.is_cuda:
            raise RuntimeError('tcow_b200.loss runs on a CUDA sm_100 device only (no CPU fallback)')
        x, gx = _grouped(logits, 'my_mask_loss.logits')
        y, gy = _grouped(target, 'my_mask_loss.target')
        w, gw = _grouped(weights.expand_as(logits), 'my_mask_loss.weights')
        B, Q, T, H, W = x.shape
        if (H * W) % 4:
            raise ValueError('my_mask_loss: H*W must be a multiple of 4')
        lib = _lib.load()
        state = torch.empty(int(lib.tcow_mask_loss_state_bytes()), device=x.device, dtype=torch.uint8)
        sel = torch.empty(B * Q * T, device=x.device, dtype=torch.uint8)
        out = torch.empty((), device=x.device, dtype=torch.float32)
        _lib.call('tcow_mask_loss_forward', x.data_ptr(), gx, y.data_ptr(), gy, w.data_ptr(), gw, B * Q, T, H * W, int(focal),
                  int(weighted_aot), float(aot_loss), float(topk_frac), 1.0, 1.0, 0.1, sel.data_ptr(), state.data_ptr(),
                  out.data_ptr(), _stream())
        ctx.save_for_backward(x, y, w, sel, state)
        ctx.geo = (gx, gy, gw, B * Q, T, H * W, int(focal), int(weighted_aot), tuple(logits.shape))
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, y, w, sel, state = ctx.saved_tensors
        gx, gy, gw, G, T, hw, focal, weighted, shape = ctx.geo
        grad = torch.empty(shape, device=x.device, dtype=torch.float32)
        up = g_out.detach().reshape(1).float().contiguous()
        _lib.call('tcow_mask_loss_backward', x.data_ptr(), gx, y.data_ptr(), gy, w.data_ptr(), gw, G, T, hw, focal, weighted,
                  sel.data_ptr(), state.data_ptr(), up.data_ptr(), grad.data_ptr(), T * hw, 0, _stream())
        return grad, None, None, None, None, None, None


def my_mask_loss(output_mask_logits, target_mask, final_weights, progress, apply_weights_for_aot, aot_loss=0.8,
                 focal_loss=False):
    """loss.py:164-225 `MyLosses.my_mask_loss` for (B,Q,T,H,W) logits / targets / weights, on the device without a host
    synchronisation: frame selection, weighted BCE (or focal), bootstrapped top-k BCE, soft Jaccard, sqrt scaling.
    `aot_loss` / `focal_loss` are the train_args of args.py:198-202."""
    return _MyMaskLoss.apply(output_mask_logits, target_mask, final_weights, topk_fraction(float(progress)),
                             bool(apply_weights_for_aot), float(aot_loss), bool(focal_loss))


def mask_track_frame_weights(sel_occl_fracs, query_time, occluded_weight=5.0):
    """loss.py:55-81: (B,Q,T,3) soft occlusion fractions -> (B,Q,T) frame weights.  The query-frame discount is applied to
    the LAST batch element only, as the reference's loop variable leaves it (loss.py:79)."""
    fw = (sel_occl_fracs[..., 0].float() * float(occluded_weight)).clip(min=1.0)
    fw[-1, :, query_time] *= 0.2
    return fw


def class_balance_corrections(target_mask):
    """loss.py:101-122: (pos_corr, neg_corr) as a device fp32[2] from the share of target == 1 / == 0 pixels."""
    t, gt = _grouped(target_mask, 'class_balance_corrections.target')
    B, Q, T, H, W = t.shape
    counts = torch.empty(2, device=t.device, dtype=torch.int64)
    _lib.call('tcow_loss_class_counts', t.data_ptr(), gt, B * Q, T, H * W, counts.data_ptr(), _stream())
    frac = (counts / t.numel()).clip(min=0.05).double()            # fp32 division then .item() -> python float in the reference
    pos, neg = frac[0], frac[1]
    ratio = torch.where(pos > neg, neg / pos, pos / neg)
    minority_up, majority_down = ratio ** -0.3, ratio ** 0.7
    pos_corr = torch.where(pos > neg, majority_down, minority_up)
    neg_corr = torch.where(pos > neg, minority_up, majority_down)
    return torch.stack([pos_corr, neg_corr]).float()


def mask_track_pixel_weights(target_mask, snitch_occl_by_ptr=None, frame_weights=None, class_balancing=True,
                             hard_negative_factor=3.0, no_hard_negatives=False):
    """loss.py:83-148 times the frame weights (:268): (B,Q,T,H,W) float target, uint8 occlusion pointers -> weights."""
    t, gt = _grouped(target_mask, 'mask_track_pixel_weights.target')
    B, Q, T, H, W = t.shape
    occl, go = None, 0
    if snitch_occl_by_ptr is not None:
        occl = snitch_occl_by_ptr if snitch_occl_by_ptr.dtype == torch.uint8 else (snitch_occl_by_ptr != 0).to(torch.uint8)
        if not (occl.stride(4) == 1 and occl.stride(3) == W and occl.stride(2) == H * W
                and (B == 1 or Q == 1 or occl.stride(0) == Q * occl.stride(1))):
            occl = occl.contiguous()
        go = occl.stride(1) if Q > 1 else (occl.stride(0) if B > 1 else T * H * W)
    corr = class_balance_corrections(t) if class_balancing else None
    hnf = float(hard_negative_factor) if (hard_negative_factor > 1.0 and not no_hard_negatives) else 1.0
    band = int((H * W) ** 0.5 / 12.0)
    band += 1 - (band % 2)
    tmp = torch.empty(B * Q * T * H * W, device=t.device, dtype=torch.uint8) if hnf > 1.0 else None
    fw = None if frame_weights is None else frame_weights.float().contiguous()
    out = torch.empty((B, Q, T, H, W), device=t.device, dtype=torch.float32)
    _lib.call('tcow_loss_pixel_weights', t.data_ptr(), gt, 0 if occl is None else occl.data_ptr(), go,
              0 if fw is None else fw.data_ptr(), 0 if corr is None else corr.data_ptr(), B * Q, T, H, W, hnf, band,
              0 if tmp is None else tmp.data_ptr(), out.data_ptr(), _stream())
    return out


def occlusion_flag_loss(output_flag, target_flag):
    """loss.py:150-162: BCE over the flags whose target is not 2 (out of frame)."""
    keep = target_flag != 2
    return torch.nn.functional.binary_cross_entropy_with_logits(output_flag[keep].float(), target_flag[keep].float())


def seeker_mask_track_loss(output_mask, target_mask, sel_occl_fracs, snitch_occl_by_ptr, query_time, progress,
                           track_lw=1.0, occl_mask_lw=0.5, cont_mask_lw=0.5, occluded_weight=5.0, occl_cont_zero_weight=0.02,
                           class_balancing=True, focal_loss=False, aot_loss=0.8, hard_negative_factor=3.0):
    """The seeker's training objective for one (sub-)batch: loss.py:236-318 per_example_mask_track + the weighted total of
    :352-354 with the defaults of args.py:182-206.  output_mask / target_mask (B,Q,3,T,H,W), sel_occl_fracs (B,Q,T,3),
    snitch_occl_by_ptr (B,Q,1,T,H,W) uint8.  Returns (total, dict of the three terms)."""
    terms = {}
    kw = dict(aot_loss=aot_loss, focal_loss=focal_loss)
    if track_lw > 0.0:
        fw = mask_track_frame_weights(sel_occl_fracs, query_time, occluded_weight)
        w = mask_track_pixel_weights(target_mask[:, :, 0], snitch_occl_by_ptr[:, :, 0], fw, class_balancing,
                                     hard_negative_factor)
        terms['track'] = my_mask_loss(output_mask[:, :, 0], target_mask[:, :, 0], w, progress, False, **kw)
    for key, ch, lw in (('occl_mask', 1, occl_mask_lw), ('cont_mask', 2, cont_mask_lw)):
        if lw > 0.0:
            present = (target_mask[:, :, ch] != 0).flatten(-2).any(-1).float()                     # (B,Q,T)
            fw = present * (1.0 - occl_cont_zero_weight) + occl_cont_zero_weight
            w = fw[..., None, None].expand_as(target_mask[:, :, ch])
            terms[key] = my_mask_loss(output_mask[:, :, ch], target_mask[:, :, ch], w, progress, True, **kw)
    total = sum(terms[k] * lw for k, lw in (('track', track_lw), ('occl_mask', occl_mask_lw), ('cont_mask', cont_mask_lw))
                if k in terms)
    return total, terms


def mask_track_metrics(output_mask, target_mask):
    """eval/metrics.py:9-113 calculate_metrics_mask_track for (B,Q,3,T,H,W) logits / targets: the areas come from one pass
    of tcow_mask_iou_areas, the per-frame IoU bookkeeping of :51-82 (a Python triple loop over numpy arrays in the
    reference) is a handful of masked means over the (B,Q,3,T) area table, still on the device."""
    from . import ops
    a = ops.mask_iou_areas(output_mask.float().contiguous(), target_mask.float().contiguous()).double()   # (B,Q,3,T,3)
    gt, inter, union = a[..., 0], a[..., 1], a[..., 2]
    iou = inter / (union + 1e-7)
    C = gt.shape[2]
    has = gt > 0
    out = {}

    def put(name, sel, values):
        n = sel.sum()
        out['mean_' + name] = torch.where(n > 0, (values * sel).sum() / n.clamp(min=1), torch.full_like(n, -1.0, dtype=torch.float64)).float()
        out['count_' + name] = n.to(torch.int32)

    zero = torch.zeros_like(has[:, :, 0])
    put('snitch_iou', has[:, :, 0], iou[:, :, 0])
    put('occl_mask_iou', has[:, :, 1] if C >= 2 else zero, iou[:, :, 1] if C >= 2 else iou[:, :, 0])
    put('cont_mask_iou', has[:, :, 2] if C >= 3 else zero, iou[:, :, 2] if C >= 3 else iou[:, :, 0])
    put('snitch_during_vis_iou', has[:, :, 0] & ~has[:, :, 1] if C >= 2 else zero, iou[:, :, 0])
    put('snitch_during_occl_iou', has[:, :, 0] & has[:, :, 1] if C >= 2 else zero, iou[:, :, 0])
    put('snitch_during_cont_iou', has[:, :, 0] & has[:, :, 2] if C >= 3 else zero, iou[:, :, 0])
    return out
