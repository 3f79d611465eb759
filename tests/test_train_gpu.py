"""Training-step parity on the GPU (BASELINE configs[3]): the hand-written backward (tcow_b200/train_engine.py, through
the C ABI) against the fp32 autograd of the oracle on the same seeded weights, clips and targets, and against the
gradient fixtures produced by the UNMODIFIED reference (oracle/make_golden_grads.py).

Tolerance, per parameter tensor: cosine >= 0.999 and relative L2 error <= 2e-2 for the
weight matrices; 1-D tensors (biases, LayerNorm affines — sums of bf16-rounded gradients over all tokens) are held to
cosine >= 0.998 and rel-L2 <= 5e-2."""
import pytest
import torch

from conftest import cached_state_dict
from oracle import make_golden_grads as mgg
from test_train_host import GRAD_CASES, check_against_fixture, load_drop_path, load_grad_golden

import tcow_b200
from tcow_b200 import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def build(logger, meta, **over):
    kw = dict(meta['ref_kwargs'])
    kw.update(over)
    net = tcow_b200.Seeker(logger, **kw)
    net.load_state_dict(cached_state_dict(meta['weight_seed'], meta['T'], meta['Hf'], meta['Wf'], kw['flag_channels']))
    net = net.to(DEV).train()
    if meta.get('pretrained_norm'):
        net.seeker.tracker_backbone.pretrained = True
    return net


def case_data(meta):
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    rgb, q = synth.make_batch(meta['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    tm, tf = synth.make_targets(meta['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    return rgb, q, tm, tf


def compare(grads, ref, where=''):
    """Per-tensor cosine / rel-L2 of {name: grad} against the fp32 reference gradients; returns the worst offenders."""
    bad = []
    worst_cos, worst_rel = 1.0, 0.0
    for name, r in ref.items():
        g = grads[name].detach().float().cpu()
        assert g.shape == r.shape, name
        rn = r.norm().item()
        if rn < 1e-12:
            if g.norm().item() > 1e-9:
                bad.append((name, 'expected zero gradient', g.norm().item()))
            continue
        cos = torch.nn.functional.cosine_similarity(g.reshape(1, -1).double(), r.reshape(1, -1).double()).item()
        rel = ((g - r).norm() / rn).item()
        one_d = r.dim() == 1 or 'norm' in name
        cos_min, rel_max = (0.998, 5e-2) if one_d else (0.999, 2e-2)
        worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
        if cos < cos_min or rel > rel_max:
            bad.append((name, round(cos, 5), round(rel, 4)))
    return bad, worst_cos, worst_rel


def our_grads(net, rgb, q, tm, tf):
    net.zero_grad(set_to_none=True)
    mask, flags = net(rgb.to(DEV), q.to(DEV))
    loss = synth.training_loss(mask, flags, tm.to(DEV), None if flags is None else tf.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    # parameters outside the graph keep .grad = None, exactly as under the reference's autograd: the final norm when
    # norm_embeddings is off (vision_tf.py:152-153) — an optimizer with weight decay must skip it (train.py:233)
    unused = {n for n, p in net.named_parameters() if p.grad is None}
    final_norm = {synth.BACKBONE_PREFIX + 'norm.weight', synth.BACKBONE_PREFIX + 'norm.bias'}
    assert unused == (set() if net.seeker.norm_embeddings else final_norm), unused
    grads = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in net.named_parameters()}
    return float(loss.detach()), grads, mask.detach()


@pytest.mark.parametrize('name', GRAD_CASES)
def test_backward_matches_oracle_autograd_and_reference_fixture(name, logger):
    meta, norms, samples, projs = load_grad_golden(name)
    net = build(logger, meta)
    dp = load_drop_path(name, meta)          # the Bernoulli draws the reference made (stochastic-depth fixtures)
    net.seeker.train_engine().drop_path_override = dp
    rgb, q, tm, tf = case_data(meta)
    loss, grads, _ = our_grads(net, rgb, q, tm, tf)
    assert all(g is not None for g in grads.values())
    assert abs(loss - meta['loss']) < 5e-3
    sd = cached_state_dict(meta['weight_seed'], meta['T'], meta['Hf'], meta['Wf'])
    _, ref = mgg.oracle_grads(sd, meta, rgb, q, tm, tf, dp)
    bad, wc, wr = compare(grads, ref)
    assert not bad, f'{len(bad)} tensors out of tolerance (worst cos {wc:.5f}, rel {wr:.4f}): {bad[:8]}'
    # and against what the reference itself produced (norms, sampled entries, projections per tensor)
    dev, where = check_against_fixture(grads, meta, norms, samples, projs, rel_tol=5e-2)
    assert dev <= 1.0, f'gradient deviates from the reference fixture at {where}: {dev:.3f} x tolerance'


def test_unmerged_temporal_projection_backward(logger):
    meta, *_ = load_grad_golden('grad_small_causal1')
    net = build(logger, meta)
    net.seeker.train_engine().merge_temporal_proj = False
    rgb, q, tm, tf = case_data(meta)
    _, grads, _ = our_grads(net, rgb, q, tm, tf)
    sd = cached_state_dict(meta['weight_seed'], meta['T'], meta['Hf'], meta['Wf'])
    _, ref = mgg.oracle_grads(sd, meta, rgb, q, tm, tf)
    bad, wc, wr = compare(grads, ref)
    assert not bad, f'{bad[:8]} (worst cos {wc:.5f}, rel {wr:.4f})'


def test_training_forward_equals_inference_forward(logger):
    """The saved-activation forward and the inference plan are the same function (same kernels, same order)."""
    meta, *_ = load_grad_golden('grad_small_causal1')
    net = build(logger, meta)
    rgb, q, tm, tf = case_data(meta)
    _, _, mask_train = our_grads(net, rgb, q, tm, tf)
    with torch.no_grad():
        mask_inf, _ = net(rgb.to(DEV), q.to(DEV))
    # same plan, but the inference path embeds the patches with the one-kernel form (accumulate, then add the embeddings)
    # and the training path with gather + embed_init + reduce-add GEMM (embeddings first): different fp32 rounding in the
    # first layer, amplified by 12 bf16 blocks to ~3e-3 — two equally valid bf16 evaluations of the same function
    assert (mask_train - mask_inf).abs().max().item() <= 5e-3


def test_query_loop_then_one_backward_and_accumulation(logger):
    """pipeline.py:134-182 calls the seeker once per query and back-propagates through all calls at once; the batched
    forward_queries must give the same gradients; a second backward accumulates into .grad like autograd does."""
    meta, *_ = load_grad_golden('grad_small_causal1')
    net = build(logger, meta)
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    rgb, q0 = synth.make_batch([0], num_frames=T, frame_height=Hf, frame_width=Wf)
    _, q1 = synth.make_batch([1], num_frames=T, frame_height=Hf, frame_width=Wf)
    tm, tf = synth.make_targets([0, 1], num_frames=T, frame_height=Hf, frame_width=Wf)
    rgb, q0, q1, tm, tf = (t.to(DEV) for t in (rgb, q0, q1, tm, tf))
    net.zero_grad(set_to_none=True)
    outs = [net(rgb, qq) for qq in (q0, q1)]                      # two forwards alive at once
    loss = synth.training_loss(torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs]), tm, tf)
    loss.backward()
    g_loop = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    mask, flags = net.forward_queries(rgb, torch.stack([q0, q1], 1))
    synth.training_loss(mask[0], flags[0], tm, tf).backward()
    g_batched = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    assert set(g_batched) == set(g_loop) and len(g_loop) == 249
    for n in g_loop:
        d = (g_loop[n] - g_batched[n]).norm().item()
        assert d <= 2e-2 * g_loop[n].norm().item() + 1e-9, n
    mask, flags = net.forward_queries(rgb, torch.stack([q0, q1], 1))
    synth.training_loss(mask[0], flags[0], tm, tf).backward()   # accumulates
    for n, p in net.named_parameters():
        if n in g_batched:
            assert (p.grad - 2 * g_batched[n]).norm().item() <= 1e-3 * g_batched[n].norm().item() + 1e-9, n


def test_optimizer_step_reduces_loss_and_repacks_weights(logger):
    meta, *_ = load_grad_golden('grad_small_causal1')
    net = build(logger, meta)
    rgb, q, tm, tf = case_data(meta)
    opt = torch.optim.SGD(net.parameters(), lr=2e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        mask, flags = net(rgb.to(DEV), q.to(DEV))
        loss = synth.training_loss(mask, flags, tm.to(DEV), tf.to(DEV))
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0] - 1e-3, losses


def test_stochastic_depth_draws(logger):
    """train() with drop_path_rate > 0 draws its own masks: outputs vary between calls with the expected frequency
    structure (block 0 never drops); eval() is deterministic and equals the drop-free forward."""
    meta, *_ = load_grad_golden('grad_small_causal1')
    net = build(logger, meta, drop_path_rate=0.5)
    rgb, q, tm, tf = case_data(meta)
    torch.manual_seed(3)
    a, _ = net(rgb.to(DEV), q.to(DEV))
    b, _ = net(rgb.to(DEV), q.to(DEV))
    assert a.requires_grad and (a - b).abs().max().item() > 1e-3
    net.eval()
    c, _ = net(rgb.to(DEV), q.to(DEV))
    d, _ = net(rgb.to(DEV), q.to(DEV))
    assert torch.equal(c, d)
    with torch.no_grad():
        e, _ = net(rgb.to(DEV), q.to(DEV))
    assert (c - e).abs().max().item() <= 5e-3      # training plan vs inference plan: see test_training_forward_equals_inference_forward
    scales = net.seeker.train_engine()._drop_path_scales(net.seeker.train(), 12, 2, 6, 4, 48, 50, 50, True, 1, torch.device(DEV))
    assert scales[0] is None and all(s is not None for s in scales[1:])
    last = scales[-1]['rs_t']
    assert set(last.unique().tolist()) <= {0.0, 2.0}          # keep = 0.5 in the last block: survivors scaled by 2


def test_backward_full_size_vs_oracle_autograd(logger):
    """BASELINE's full shape (T=30, 240x320, 301 tokens per frame, causal): one sample, gradients of all 251 tensors
    against the fp32 autograd of the oracle run on the host cores (about half a minute)."""
    import os
    T, Hf, Wf = 30, 240, 320
    meta = dict(T=T, Hf=Hf, Wf=Wf, causal=1, weight_seed=901, samples=[3],
                ref_kwargs=dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                                tracker_pretrained=False, attention_type='divided_space_time', patch_size=16,
                                causal_attention=1, norm_embeddings=False, drop_path_rate=0.0, network_depth=12,
                                track_map_stride=4, track_map_resize='bilinear', query_channels=1, output_channels=3,
                                flag_channels=3))
    net = build(logger, meta)
    rgb, q, tm, tf = case_data(meta)
    loss, grads, _ = our_grads(net, rgb, q, tm, tf)
    torch.set_num_threads(os.cpu_count() or 8)
    sd = cached_state_dict(901, T, Hf, Wf)
    oloss, ref = mgg.oracle_grads(sd, meta, rgb, q, tm, tf)
    assert abs(loss - oloss) < 5e-3
    bad, wc, wr = compare(grads, ref)
    assert not bad, f'{len(bad)} tensors out of tolerance (worst cos {wc:.5f}, rel {wr:.4f}): {bad[:8]}'


def test_fused_mask_loss_matches_torch_restatement():
    """tcow_b200.loss.mask_loss_terms against the reference formulas (loss.py:19-31 tversky_loss, :181-183 weighted BCE),
    values and gradients; also the no-target branch (loss.py:20)."""
    from tcow_b200 import loss as tl
    g = torch.Generator(device=DEV).manual_seed(5)
    shape = (2, 3, 3, 4, 32, 48)
    x = (torch.randn(shape, device=DEV, generator=g) * 3).requires_grad_(True)
    y = (torch.rand(shape, device=DEV, generator=g) > 0.8).float()
    w = torch.rand(shape, device=DEV, generator=g) * 2

    def ref(x, y, w):
        bce = (torch.nn.functional.binary_cross_entropy_with_logits(x, y, reduction='none') * w).mean()
        if y.mean() >= 1e-6:
            p = torch.sigmoid(x)
            num = (p * y).sum()
            den = num + 1.0 * (p * (1 - y)).sum() + 1.0 * ((1 - p) * y).sum()
            tv = 1.0 - num / (den + 0.1)
        else:
            tv = torch.tensor(0.0, device=x.device)
        return bce, tv

    for yy, ww in ((y, w), (y, None), (torch.zeros_like(y), w)):
        x.grad = None
        b1, t1 = tl.mask_loss_terms(x, yy, ww)
        (0.7 * b1 + 1.3 * t1).backward()
        g1 = x.grad.clone()
        x.grad = None
        b2, t2 = ref(x, yy, ww if ww is not None else torch.ones_like(x))
        (0.7 * b2 + 1.3 * t2).backward()
        assert abs(b1.item() - b2.item()) <= 1e-5 * max(1.0, abs(b2.item()))
        assert abs(t1.item() - t2.item()) <= 1e-5
        assert (g1 - x.grad).abs().max().item() <= 2e-6 * max(1.0, x.grad.abs().max().item() * 1e3)
        assert ((g1 - x.grad).norm() / x.grad.norm()).item() <= 1e-4


@pytest.mark.parametrize('cfg', [dict(T=1, Hf=32, Wf=32, samples=[9], causal=1),
                                 dict(T=3, Hf=16, Wf=48, samples=[4], causal=0, flag_channels=0, track_map_resize='nearest'),
                                 dict(T=7, Hf=48, Wf=16, samples=[1, 2, 3], causal=-1)],
                         ids=['single_frame', 'no_flags_nearest', 'no_cls_three_samples'])
def test_backward_edge_configs_vs_oracle_autograd(cfg, logger):
    """Edge shapes of the training step (one frame, one patch row, no flag head, no cls token) against the oracle autograd."""
    T, Hf, Wf = cfg['T'], cfg['Hf'], cfg['Wf']
    fc = cfg.get('flag_channels', 3)
    meta = dict(cfg, weight_seed=901, flag_channels=fc,
                ref_kwargs=dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                                tracker_pretrained=False, attention_type='divided_space_time', patch_size=16,
                                causal_attention=cfg['causal'], norm_embeddings=False, drop_path_rate=0.0, network_depth=12,
                                track_map_stride=4, track_map_resize=cfg.get('track_map_resize', 'bilinear'),
                                query_channels=1, output_channels=3, flag_channels=fc))
    net = build(logger, meta)
    rgb, q = synth.make_batch(cfg['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    tm, tf = synth.make_targets(cfg['samples'], num_frames=T, frame_height=Hf, frame_width=Wf, flag_channels=fc)
    loss, grads, _ = our_grads(net, rgb, q, tm, tf)
    sd = cached_state_dict(901, T, Hf, Wf, fc)
    oloss, ref = mgg.oracle_grads(sd, meta, rgb, q, tm, tf)
    assert abs(loss - oloss) < 5e-3
    bad, wc, wr = compare(grads, ref)
    assert not bad, f'{len(bad)} tensors out of tolerance (worst cos {wc:.5f}, rel {wr:.4f}): {bad[:8]}'
