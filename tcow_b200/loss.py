"""Mask-loss terms of the training step on the device: the weighted BCE of loss.py:164-185 and the
soft-Jaccard / Tversky term of loss.py:19-31 from TWO bandwidth passes over the logits (sums, then the gradient) instead
of the ~25 elementwise kernels torch needs for the same expressions and their autograd.  The reference's top-k
bootstrapping (loss.py:12-16) stays in torch (a global selection), fed by the same logits."""
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
