"""End-to-end parity of the CUDA Seeker (through Seeker.forward -> ctypes -> libtcow_b200.so) against
(a) the golden vectors produced by the UNMODIFIED reference and (b) the fp32 oracle on the same inputs.
Tolerance (BASELINE.json north_star): max |delta logit| <= 1e-2 for bf16 vs the fp32 reference, binarised
mask (logit > 0, eval/metrics.py:18) IoU >= 0.99."""
import pytest
import torch

import tcow_b200
from conftest import cached_state_dict, golden_names, load_golden
from oracle import seeker_oracle
from tcow_b200 import synth

pytestmark = pytest.mark.gpu
TOL_LOGIT = 1e-2
TOL_IOU = 0.99


DEV = 'cuda:0'


def build(logger, meta, **over):
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    fc = meta.get('flag_channels', 3)
    kw = dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf, tracker_pretrained=False,
              attention_type='divided_space_time', patch_size=16, causal_attention=meta['causal'],
              norm_embeddings=meta.get('norm_embeddings', False), drop_path_rate=0.1, network_depth=12,
              track_map_stride=4, track_map_resize=meta.get('track_map_resize', 'bilinear'), query_channels=1,
              output_channels=3, flag_channels=fc)
    kw.update(over)
    net = tcow_b200.Seeker(logger, **kw)
    net.load_state_dict(cached_state_dict(meta.get('weight_seed', 901), T, Hf, Wf, fc), strict=True)
    net = net.cuda().eval()
    if meta.get('pretrained_norm'):
        net.seeker.tracker_backbone.pretrained = True
    return net


def iou(a, b):
    a, b = a > 0, b > 0
    union = (a | b).sum().item()
    return 1.0 if union == 0 else (a & b).sum().item() / union


@pytest.mark.parametrize('name', [n for n in golden_names() if not n.startswith('full_')])
def test_matches_reference_golden(name, logger):
    meta, gmask, gflags = load_golden(name)
    net = build(logger, meta)
    rgb, q = synth.make_batch(meta['samples'], num_frames=meta['T'], frame_height=meta['Hf'],
                              frame_width=meta['Wf'], query_frame=meta.get('query_frame', 0))
    with torch.no_grad():
        mask, flags = net(rgb.cuda(), q.cuda())
    assert mask.dtype == torch.float32 and tuple(mask.shape) == (len(meta['samples']), 3, meta['T'], meta['Hf'], meta['Wf'])
    ly, lx = meta.get('lattice', (1, 1))
    m = mask.cpu()[:, :, :, ::ly, ::lx]
    err = (m - gmask).abs().max().item()
    assert err <= TOL_LOGIT, f'{name}: max |dlogit| {err:.3e}'
    for s in range(m.shape[0]):
        assert iou(m[s], gmask[s]) >= TOL_IOU
    if gflags is None:
        assert flags is None
    else:
        assert tuple(flags.shape) == tuple(gflags.shape)
        assert (flags.cpu() - gflags).abs().max().item() <= 2e-2


def test_full_size_batch_matches_reference_golden(logger):
    """North-star shape: T=30, 240x320, causal=1, two clips in one batch."""
    meta, gmask, gflags = load_golden('full_causal1')
    net = build(logger, meta)
    rgb, q = synth.make_batch(meta['samples'], num_frames=30, frame_height=240, frame_width=320)
    with torch.no_grad():
        mask, flags = net(rgb.cuda(), q.cuda())
    ly, lx = meta['lattice']
    m = mask.cpu()[:, :, :, ::ly, ::lx]
    err = (m - gmask).abs().max().item()
    ious = [iou(m[s], gmask[s]) for s in range(m.shape[0])]
    print(f'full-size: max|dlogit| {err:.4e}  IoU {ious}  flags err {(flags.cpu() - gflags).abs().max().item():.4e}')
    assert err <= TOL_LOGIT
    assert min(ious) >= TOL_IOU
    assert (flags.cpu() - gflags).abs().max().item() <= 2e-2


def test_full_size_vs_oracle_dense(logger):
    """Every pixel of one full-size clip against the fp32 oracle (CPU), not just the golden lattice."""
    meta, _, _ = load_golden('full_causal1')
    net = build(logger, meta)
    rgb, q = synth.make_batch([5], num_frames=30, frame_height=240, frame_width=320, query_frame=0)
    with torch.no_grad():
        mask, flags = net(rgb.cuda(), q.cuda())
        omask, oflags = seeker_oracle.seeker_forward(cached_state_dict(901, 30, 240, 320), rgb, q, causal_attention=1)
    err = (mask.cpu() - omask).abs().max().item()
    print(f'dense: max|dlogit| {err:.4e} IoU {iou(mask.cpu(), omask):.5f}')
    assert err <= TOL_LOGIT and iou(mask.cpu(), omask) >= TOL_IOU
    assert (flags.cpu() - oflags).abs().max().item() <= 2e-2


def test_unmerged_temporal_projection_path(logger):
    """proj and temporal_fc as two GEMMs (training-style) agrees with the reference too."""
    meta, gmask, _ = load_golden('mid_causal1')
    net = build(logger, meta)
    net.seeker.engine().merge_temporal_proj = False
    rgb, q = synth.make_batch(meta['samples'], num_frames=meta['T'], frame_height=meta['Hf'], frame_width=meta['Wf'],
                              query_frame=meta['query_frame'])
    with torch.no_grad():
        mask, _ = net(rgb.cuda(), q.cuda())
    ly, lx = meta['lattice']
    assert (mask.cpu()[:, :, :, ::ly, ::lx] - gmask).abs().max().item() <= TOL_LOGIT


def test_causality_bit_exact(logger):
    """causal_attention=1: frames >= t0 cannot influence outputs before t0 — bit-identical, as in the reference."""
    meta = dict(T=6, Hf=32, Wf=32, causal=1)
    net = build(logger, meta)
    rgb, q = synth.make_batch([11], num_frames=6, frame_height=32, frame_width=32)
    rgb2 = rgb.clone()
    rgb2[:, :, 4:] = torch.rand_like(rgb2[:, :, 4:])
    with torch.no_grad():
        m1, f1 = net(rgb.cuda(), q.cuda())
        m2, f2 = net(rgb2.cuda(), q.cuda())
    assert torch.equal(m1[:, :, :4], m2[:, :, :4]) and torch.equal(f1[:, :4], f2[:, :4])
    assert (m1[:, :, 4:] - m2[:, :, 4:]).abs().max().item() > 1e-3


def test_batch_invariance_and_chunking(logger):
    """A clip's output does not depend on its batch neighbours nor on the engine's chunk size."""
    meta = dict(T=4, Hf=32, Wf=48, causal=1)
    net = build(logger, meta)
    rgb, q = synth.make_batch([0, 1, 2], num_frames=4, frame_height=32, frame_width=48)
    with torch.no_grad():
        m3, f3 = net(rgb.cuda(), q.cuda())
        m1, f1 = net(rgb[1:2].cuda(), q[1:2].cuda())
        net.seeker.engine().max_chunk = 2
        mc, fc = net(rgb.cuda(), q.cuda())
    assert torch.equal(m3[1:2], m1) and torch.equal(f3[1:2], f1)
    assert torch.equal(m3, mc) and torch.equal(f3, fc)


def test_inputs_not_modified_and_dtype_cast(logger):
    meta = dict(T=4, Hf=32, Wf=48, causal=1, pretrained_norm=True)
    net = build(logger, meta)
    rgb, q = synth.make_batch([3], num_frames=4, frame_height=32, frame_width=48)
    r, qq = rgb.cuda(), q.cuda()
    r0 = r.clone()
    with torch.no_grad():
        m32, _ = net(r, qq)
        m64, _ = net(r.double(), qq.bool())          # mask_tracker.py:103-104 casts any dtype to fp32
    assert torch.equal(r, r0)
    assert torch.equal(m32, m64)


def test_forward_queries_equals_per_query_loop(logger):
    """forward_queries == the reference pipeline's loop over queries on shared frames (pipeline.py:134-182)."""
    meta = dict(T=4, Hf=32, Wf=48, causal=1)
    net = build(logger, meta)
    rgb, _ = synth.make_batch([0, 1], num_frames=4, frame_height=32, frame_width=48)
    qs = torch.stack([torch.stack([synth.make_clip(100 + 3 * b + k, 4, 32, 48)[1] for k in range(3)]) for b in range(2)])
    with torch.no_grad():
        m, f = net.forward_queries(rgb.cuda(), qs.cuda())                      # (2,3,3,T,H,W), (2,3,T,3)
        net.seeker.engine().max_chunk = 4                                        # chunk boundary inside a video
        m4, f4 = net.forward_queries(rgb.cuda(), qs.cuda())
        loop = [net(rgb.cuda(), qs[:, k].cuda()) for k in range(3)]
    assert tuple(m.shape) == (2, 3, 3, 4, 32, 48) and tuple(f.shape) == (2, 3, 4, 3)
    for k in range(3):
        assert torch.equal(m[:, k], loop[k][0]) and torch.equal(f[:, k], loop[k][1])
    assert torch.equal(m, m4) and torch.equal(f, f4)


def test_weight_update_invalidates_packed_cache(logger):
    meta = dict(T=4, Hf=32, Wf=48, causal=1)
    net = build(logger, meta)
    rgb, q = synth.make_batch([3], num_frames=4, frame_height=32, frame_width=48)
    with torch.no_grad():
        m1, _ = net(rgb.cuda(), q.cuda())
        net.seeker.tracker_post_linear.bias.add_(1.0)
        m2, _ = net(rgb.cuda(), q.cuda())
    assert (m2 - m1 - 1.0).abs().max().item() < 1e-4


def test_errors_match_reference_types(logger):
    meta = dict(T=4, Hf=32, Wf=48, causal=1)
    net = build(logger, meta)
    rgb, q = synth.make_batch([3], num_frames=4, frame_height=32, frame_width=48)
    with torch.no_grad():
        with pytest.raises(AssertionError):
            net(rgb[:, :, :3].cuda(), q[:, :, :3].cuda())           # vision_tf.py:96  assert T == self.T
        with pytest.raises(AssertionError):
            net(rgb.cuda(), torch.cat([q, q], 1).cuda())            # mask_tracker.py:105
    mask, flags = net(rgb.cuda(), q.cuda())                         # grad mode: the training path (train_engine.py)
    assert mask.requires_grad and flags.requires_grad
    with pytest.raises(RuntimeError, match='inference plan'):       # the inference engine itself refuses gradient mode
        net.seeker.engine().forward(net.seeker, rgb.cuda(), q.cuda())


def test_sweep_runner_matches_per_item_forwards(logger):
    """The clip-grouped sweep (forward_queries + on-device IoU areas) gives, per item, what a plain per-sample forward and
    the torch restatement of eval/metrics.py:18-41 give."""
    from tcow_b200 import ops, sweep
    T, Hf, Wf, F = 4, 32, 48, 12
    net = build(logger, dict(T=T, Hf=Hf, Wf=Wf, causal=1, weight_seed=901))
    items = sweep.plan_sweep(num_videos=2, num_queries=3, num_video_frames=F, num_frames=T, query_idx=1)
    assert len(items) == 2 * 3 * 3
    g = torch.Generator().manual_seed(3)
    vids = {v: torch.rand(3, F, Hf, Wf, generator=g) for v in range(2)}
    tgt = (torch.rand(3, F, Hf, Wf, generator=g) > 0.5).float()

    def get_query(v, q):
        m = torch.zeros(Hf, Wf)
        m[4 * q:4 * q + 8, 6 * q:6 * q + 10] = 1
        return m

    res = sweep.run_sweep(net, items, vids.__getitem__, get_query, lambda v, q: tgt, T, torch.device(DEV), clips_per_pass=2)
    assert sorted(res) == sorted((i.video, i.query, i.frame_start, i.frame_stride) for i in items)
    it = items[7]
    idx = torch.arange(T) * it.frame_stride + it.frame_start
    q = torch.zeros(1, 1, T, Hf, Wf)
    q[0, 0, 0] = get_query(it.video, it.query)
    with torch.no_grad():
        mask, flags = net(vids[it.video][:, idx][None].to(DEV), q.to(DEV))
    pred, gt = (mask[0] > 0).cpu(), tgt[:, idx] > 0.5
    iou = (pred & gt).sum((-1, -2)).float() / ((pred | gt).sum((-1, -2)).float() + 1e-7)
    row = res[(it.video, it.query, it.frame_start, it.frame_stride)]
    assert abs(row['mean_snitch_iou'] - iou[0].mean().item()) < 2e-2      # a few boundary pixels may flip (bf16)
    assert row['count_snitch_iou'] == T
    a = ops.mask_iou_areas(mask.contiguous(), tgt[:, idx][None].to(DEV).contiguous()).cpu()[0]
    assert torch.equal(a[..., 0], gt.sum((-1, -2)).float()) and torch.equal(a[..., 1], (pred & gt).sum((-1, -2)).float())
    assert torch.equal(a[..., 2], (pred | gt).sum((-1, -2)).float())


def test_timed_bench_batch_matches_reference_golden(logger):
    """The exact batch bench.py times (BASELINE configs[1]: 8 clips in one pass, M = 72 000 rows, CTA-pair GEMMs): clips 0
    and 7 against the unmodified reference's outputs on the same clips (tests/golden/full_bench_b8.npz)."""
    meta, gmask, gflags = load_golden('full_bench_b8')
    net = build(logger, meta)
    rgb, q = synth.bench_clips(0, meta['bench_batch'], 30, 240, 320)
    with torch.no_grad():
        mask, flags = net(rgb.cuda(), q.cuda())
    ly, lx = meta['lattice']
    m = mask[meta['samples']].cpu()[:, :, :, ::ly, ::lx]
    err = (m - gmask).abs().max().item()
    ious = [iou(m[s], gmask[s]) for s in range(m.shape[0])]
    print(f'bench batch: max|dlogit| {err:.4e}  IoU {ious}')
    assert err <= TOL_LOGIT and min(ious) >= TOL_IOU
    assert (flags[meta['samples']].cpu() - gflags).abs().max().item() <= 2e-2


def test_uint8_inputs_with_device_side_scaling(logger):
    """decoder-style uint8 frames / query masks go straight into the gather kernel.  frame_scale=1/255
    reproduces the host-side `rgb / 255.0` of data/data_plugin.py:174; without it uint8 is the plain cast of
    mask_tracker.py:103 (bit-identical to passing .float())."""
    meta, gmask, _ = load_golden('mid_causal1')
    net = build(logger, meta)
    rgb, q = synth.make_batch(meta['samples'], num_frames=meta['T'], frame_height=meta['Hf'], frame_width=meta['Wf'],
                              query_frame=meta['query_frame'])
    rgb8 = (rgb * 255).round().to(torch.uint8)
    with torch.no_grad():
        m8, f8 = net(rgb8.cuda(), q.to(torch.uint8).cuda(), frame_scale=1.0 / 255.0)
        mf, ff = net((rgb8.float() / 255.0).cuda(), q.cuda())
        mq, fq = net.forward_queries(rgb8.cuda(), q.to(torch.uint8).cuda()[:, None], frame_scale=1.0 / 255.0)
        raw8, _ = net(rgb8.cuda(), q.cuda())
        rawf, _ = net(rgb8.float().cuda(), q.cuda())
    assert (m8 - mf).abs().max().item() <= 2e-3 and (f8 - ff).abs().max().item() <= 2e-3   # x*(1/255) vs x/255 roundings
    assert torch.equal(mq[:, 0], m8) and torch.equal(fq[:, 0], f8)
    assert torch.equal(raw8, rawf)
    ly, lx = meta['lattice']
    # quantising the frames to 8 bits moves the logits by less than the bf16 budget: still within tolerance of the golden
    assert (m8.cpu()[:, :, :, ::ly, ::lx] - gmask).abs().max().item() <= TOL_LOGIT
