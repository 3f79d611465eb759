"""The reference's complete mask-tracking objective and metrics.  CPU: the torch restatement
(oracle/loss_oracle.py) against fixtures produced by the reference's own MyLosses / calculate_metrics_mask_track
(oracle/make_golden_loss.py).  GPU: the CUDA implementation (tcow_b200/loss.py, csrc/loss_full.cu) against both."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import loss_oracle, make_golden_loss as mgl

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('loss_') and f.endswith('.npz'))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return json.loads(bytes(z['meta']).decode()), z


def kwargs_of(meta):
    return dict(occl_cont_zero_weight=meta.get('occl_cont_zero_weight', 0.02), class_balancing=meta.get('class_balancing', True),
                focal_loss=meta.get('focal_loss', False), aot_loss=meta.get('aot_loss', 0.8),
                hard_negative_factor=meta.get('hard_negative_factor', 3.0))


def test_fixture_set_is_complete():
    assert len(CASES) == 5


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_losses(name):
    meta, z = load(name)
    out, tgt, fracs, occl = mgl.make_inputs(meta)
    out = out.clone().requires_grad_(True)
    kw = kwargs_of(meta)
    fw = loss_oracle.frame_weights(fracs, meta['query_time'])
    assert np.array_equal(fw.numpy(), z['frame_weights'])
    pw = loss_oracle.pixel_weights(tgt[:, :, 0], occl[:, :, 0], kw['class_balancing'], kw['hard_negative_factor'])
    assert np.array_equal(pw.numpy(), z['pixel_weights'])          # incl. the dilation == (gaussian_blur > 0) claim
    total, terms = loss_oracle.seeker_loss(out, tgt, fracs, occl, meta['query_time'], meta['progress'], **kw)
    for k in ('track', 'occl_mask', 'cont_mask'):
        assert abs(float(terms[k]) - float(z[k])) <= 1e-6 * max(1.0, abs(float(z[k]))), k
    total.backward()
    assert np.abs(out.grad.numpy() - z['grad']).max() <= 1e-7 + 1e-5 * np.abs(z['grad']).max()
    m = loss_oracle.metrics(out.detach(), tgt)
    for k, v in m.items():
        assert abs(v - float(z['metric_' + k])) <= 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_loss_matches_reference_fixture(name):
    from tcow_b200 import loss as L
    meta, z = load(name)
    out, tgt, fracs, occl = (t.cuda() for t in mgl.make_inputs(meta))
    out = out.clone().requires_grad_(True)
    kw = kwargs_of(meta)
    fw = L.mask_track_frame_weights(fracs, meta['query_time'])
    assert np.array_equal(fw.cpu().numpy(), z['frame_weights'])
    pw = L.mask_track_pixel_weights(tgt[:, :, 0], occl[:, :, 0], None, kw['class_balancing'], kw['hard_negative_factor'])
    assert np.array_equal(pw.cpu().numpy(), z['pixel_weights'])
    sw = L.mask_track_pixel_weights(tgt[:, :, 0], occl[:, :, 0], fw, kw['class_balancing'], kw['hard_negative_factor'])
    assert np.array_equal(sw.cpu().numpy(), z['snitch_weights'])
    total, terms = L.seeker_mask_track_loss(out, tgt, fracs, occl, meta['query_time'], meta['progress'], **kw)
    for k in ('track', 'occl_mask', 'cont_mask'):
        assert abs(float(terms[k]) - float(z[k])) <= 2e-6 * max(1.0, abs(float(z[k]))), (k, float(terms[k]), float(z[k]))
    assert abs(float(total) - float(z['total'])) <= 4e-6 * max(1.0, abs(float(z['total'])))
    total.backward()
    err = np.abs(out.grad.cpu().numpy() - z['grad']).max()
    assert err <= 1e-7 + 2e-5 * np.abs(z['grad']).max(), err
    m = L.mask_track_metrics(out.detach(), tgt)
    for k, v in m.items():
        assert abs(float(v) - float(z['metric_' + k])) <= 1e-6, k


@pytest.mark.gpu
def test_cuda_loss_full_size_against_oracle_on_device():
    """Training-step size (2 videos x 3 queries, T=30, 240x320 = 13.8 M pixels per channel): exact top-k of the
    bootstrapped BCE by radix select vs torch.topk, all three channels, gradient included."""
    from tcow_b200 import loss as L
    c = dict(seed=9, B=2, Q=3, T=30, H=240, W=320, progress=0.07, query_time=0)
    out, tgt, fracs, occl = (t.cuda() for t in mgl.make_inputs(c))
    a = out.clone().requires_grad_(True)
    b = out.clone().requires_grad_(True)
    total, terms = L.seeker_mask_track_loss(a, tgt, fracs, occl, 0, c['progress'])
    ototal, oterms = loss_oracle.seeker_loss(b, tgt, fracs, occl, 0, c['progress'])
    for k in terms:
        assert abs(float(terms[k]) - float(oterms[k])) <= 1e-5 * abs(float(oterms[k])), (k, float(terms[k]), float(oterms[k]))
    total.backward()
    ototal.backward()
    err = (a.grad - b.grad).abs().max().item()
    assert err <= 1e-9 + 2e-5 * b.grad.abs().max().item(), err
    # ties: constant logits make every per-pixel loss equal -> the k selected share the gradient evenly (same total)
    x0 = torch.zeros(1, 1, 2, 16, 16, device='cuda', requires_grad=True)
    y0 = torch.zeros(1, 1, 2, 16, 16, device='cuda')
    w0 = torch.ones(1, 1, 2, 16, 16, device='cuda')
    l0 = L.my_mask_loss(x0, y0, w0, 0.05, False, aot_loss=1.0)
    l0.backward()
    assert abs(float(l0) - 0.5 * (np.log(2.0) + 0.0)) < 1e-6          # bootstrap = ln 2, jaccard = 0 (no target)
    assert abs(float(x0.grad.sum()) - 0.5 * 0.5) < 1e-6                # d(mean of k equal values)/dx summed = dl/dx = p - y


@pytest.mark.gpu
def test_cuda_loss_zero_weights_branch():
    from tcow_b200 import loss as L
    x = torch.randn(1, 2, 3, 16, 16, device='cuda', requires_grad=True)
    y = (torch.rand(1, 2, 3, 16, 16, device='cuda') > 0.5).float()
    l = L.my_mask_loss(x, y, torch.zeros_like(y), 0.1, True)           # loss.py:221: nothing selected -> 0, no gradient
    l.backward()
    assert float(l) == 0.0 and float(x.grad.abs().max()) == 0.0
