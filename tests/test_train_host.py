"""CPU-side checks of the training path: the oracle's autograd against the reference's gradient fixtures
(oracle/make_golden_grads.py), the flat gradient layout and its mapping back onto the reference parameters, and the
bucketed gradient all-reduce (gloo, world size 2)."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, cached_state_dict
from oracle import make_golden_grads as mgg
from tcow_b200 import ddp, synth
from tcow_b200.train_engine import _GradLayout

GRAD_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith('grad_') and f.endswith('.npz'))


def load_grad_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    return meta, z['norms'], z['samples'], z['projs']


def load_drop_path(name, meta):
    """The stochastic-depth keep masks the reference drew for this fixture (None when DropPath was off)."""
    if 'drop_keep' not in meta:
        return None
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    out = []
    for i, keep in enumerate(meta['drop_keep']):
        out.append(None if keep is None else
                   dict(keep=keep, **{w: torch.from_numpy(z[f'dp{i}_{w}'].astype(np.float32)) for w in 'tsm'}))
    return out


def check_against_fixture(grads, meta, norms, samples, projs, rel_tol, abs_floor=1e-9):
    """grads: {full reference parameter name: tensor}.  Returns the worst normalised deviation (<= 1 passes)."""
    worst = (0.0, '')
    for i, name in enumerate(meta['names']):
        n, s, p = mgg.summarize(name, grads[name].detach().cpu().float())
        scale = max(norms[i], abs_floor)
        dev = max(abs(n - norms[i]) / scale,
                  float(np.abs(s - samples[i]).max()) / (scale / np.sqrt(max(grads[name].numel(), 1)) * 30 + abs_floor),
                  float(np.abs(p - projs[i]).max()) / (scale * 6 + abs_floor)) / rel_tol
        if dev > worst[0]:
            worst = (dev, name)
    return worst


@pytest.mark.parametrize('name', ['grad_small_causal1', 'grad_small_causal0', 'grad_small_droppath1'])
def test_oracle_autograd_matches_reference_gradients(name):
    meta, norms, samples, projs = load_grad_golden(name)
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    sd = cached_state_dict(meta['weight_seed'], T, Hf, Wf, meta.get('flag_channels', 3))
    rgb, q = synth.make_batch(meta['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    tm, tf = synth.make_targets(meta['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    loss, grads = mgg.oracle_grads(sd, meta, rgb, q, tm, tf, load_drop_path(name, meta))
    assert abs(loss - meta['loss']) < 1e-5
    dev, where = check_against_fixture(grads, meta, norms, samples, projs, rel_tol=1e-3)
    assert dev <= 1.0, f'oracle autograd deviates from the reference fixture at {where}: {dev:.3f} x tolerance'


def test_grad_layout_is_disjoint_aligned_and_ordered():
    lay = _GradLayout(depth=12, D=768, Kp=1024, n_pos=301, T=30, n_pad=64, merged=True)
    spans = sorted((off, off + int(np.prod(shape))) for off, shape in lay.slots.values())
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:]))            # no overlap
    assert all(off % 64 == 0 for off, _ in lay.slots.values())                      # 256-byte aligned (TMA reduce)
    # ranges complete in execution order of the backward: head, block 11 .. block 0, embeddings; together they tile
    # the whole buffer
    order = [lay.head_range] + lay.block_ranges + [lay.embed_range]
    assert order[0][0] == 0 and order[-1][1] == lay.total
    assert all(a[1] == b[0] for a, b in zip(order, order[1:]))
    assert lay.slots['b11.fc2_w'][0] == lay.block_ranges[0][0]
    n_params = sum(int(np.prod(s)) for _, s in lay.slots.values())
    # merged temporal projection: one 768x768 (+bias) per block instead of two; pooled head 64x768 instead of 771x768
    # (+ a second 768-bias slot per block for the part of the merged bias that sits outside DropPath)
    assert n_params == 122145027 - 12 * (768 * 768 + 768) + 12 * 768 - (771 * 768 + 771) + (64 * 768 + 64) - 768 - 768


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lay = _GradLayout(depth=2, D=64, Kp=128, n_pos=7, T=4, n_pad=64, merged=True)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(lay.total, generator=g)
    mine = flat.clone()
    sync = ddp.GradSync(average=True, bucket_bytes=64 << 10)
    sync.begin(flat)
    for lo, hi in [lay.head_range] + lay.block_ranges + [lay.embed_range]:
        sync.ready(lo, hi)
    sync.finish()
    other = torch.randn(lay.total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok = torch.allclose(flat, (mine + other) / 2, atol=1e-6)
    covered = sorted(sync.ranges)
    tiles = covered[0][0] == 0 and covered[-1][1] == lay.total and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # parameters start identical after a broadcast
    lin = torch.nn.Linear(4, 3)
    ddp.broadcast_parameters(lin, src=0)
    w = lin.weight.detach().clone()
    dist.all_reduce(w)
    same = torch.allclose(w, 2 * lin.weight.detach())
    q.put((rank, ok, tiles, len(sync.ranges), same))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_bucketed_gradient_allreduce():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(2)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok, tiles, nranges, same in res:
        assert ok and tiles and same
        assert nranges >= 2          # more than one bucket: the exchange overlaps the backward


def _small_tracker(logger, causal=1, drop_path_rate=0.3):
    import tcow_b200
    T, Hf, Wf = 4, 32, 48
    net = tcow_b200.Seeker(logger, num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                           tracker_pretrained=False, causal_attention=causal, patch_size=16, drop_path_rate=drop_path_rate)
    net.load_state_dict(cached_state_dict(901, T, Hf, Wf))
    return net.seeker


def test_unpack_is_the_adjoint_of_the_weight_packing(logger):
    """The packed->reference gradient mapping (_unpack: un-merge of temporal_fc o proj, un-fold of the pooled head, conv
    bias / cls / pos_embed bookkeeping) against torch.autograd through the same packing written as a differentiable
    function, on the CPU."""
    from tcow_b200.train_engine import SeekerTrainEngine
    mod = _small_tracker(logger)
    eng = SeekerTrainEngine(mod)
    pk = eng._pack(mod, torch.device('cpu'))
    bb = mod.tracker_backbone.timesformer.model
    D, N, T, P = 768, 6, 4, 16
    lay = _GradLayout(len(pk.blocks), D, 4 * P * P, N + 1, T, pk.n_pad, True)
    g = torch.Generator().manual_seed(4)
    flat = torch.randn(lay.total, generator=g)
    # head rows past the used ones never receive gradient in a real backward
    n_used = 3 * 16 + 3
    lay.view(flat, 'head_w')[n_used:] = 0
    lay.view(flat, 'head_b')[n_used:] = 0
    dropped = [i % 2 == 1 for i in range(len(pk.blocks))]          # odd blocks: stochastic depth active (two bias gradients)
    got = eng._unpack(mod, pk, lay, flat, True, dropped)

    params = {n: p.detach().clone().double().requires_grad_(True) for n, p in mod.named_parameters()}
    pre = 'tracker_backbone.timesformer.model.'
    total = 0.0
    gv = lambda name: lay.view(flat, name).double()
    for i in range(len(pk.blocks)):
        q = pre + f'blocks.{i}.'
        Wp, bp = params[q + 'temporal_attn.proj.weight'], params[q + 'temporal_attn.proj.bias']
        Wf, bf = params[q + 'temporal_fc.weight'], params[q + 'temporal_fc.bias']
        total = total + (gv(f'b{i}.t_out_w') * (Wf @ Wp)).sum() + (gv(f'b{i}.t_out_b') * (Wf @ bp)).sum()
        total = total + ((gv(f'b{i}.t_out_b2') if dropped[i] else gv(f'b{i}.t_out_b')) * bf).sum()
        for ours, theirs in (('t_qkv', 'temporal_attn.qkv'), ('fc2', 'mlp.fc2'), ('n1', 'norm1')):
            w_slot, b_slot = (ours + '_g', ours + '_b') if ours == 'n1' else (ours + '_w', ours + '_b')
            total = total + (gv(f'b{i}.{w_slot}') * params[q + theirs + '.weight']).sum()
            total = total + (gv(f'b{i}.{b_slot}') * params[q + theirs + '.bias']).sum()
    Wt, bt = params['tracker_post_linear.weight'], params['tracker_post_linear.bias']
    Wpool = Wt.reshape(3, 4, 4, 4, 4, D).mean((2, 4)).reshape(48, D)
    bpool = bt.reshape(3, 4, 4, 4, 4).mean((2, 4)).reshape(48)
    total = total + (gv('head_w')[:48] * Wpool).sum() + (gv('head_b')[:48] * bpool).sum()
    total = total + (gv('head_w')[48:51] * params['flag_post_linear.weight']).sum()
    total = total + (gv('head_b')[48:51] * params['flag_post_linear.bias']).sum()
    # embeddings: X[(b,n,t)] = conv + conv_bias + pos[1+n] + time[t]; X[M+b] = cls + pos[0]  (B = 1 here)
    total = total + (gv('patch_w') * params[pre + 'patch_embed.proj.weight'].reshape(D, -1)).sum()
    total = total + (gv('time').sum(0) * params[pre + 'patch_embed.proj.bias']).sum()
    total = total + (gv('pos') * params[pre + 'pos_embed'][0]).sum() + (gv('time') * params[pre + 'time_embed'][0]).sum()
    total = total + (gv('pos')[0] * params[pre + 'cls_token'][0, 0]).sum()
    total.backward()
    checked = 0
    for name, p in params.items():
        if p.grad is None:
            continue
        ref = p.grad.float()
        rel = ((got[name].float() - ref).norm() / ref.norm().clamp_min(1e-20)).item()
        assert got[name].shape == ref.shape and rel < 1e-5, (name, rel)
        checked += 1
    assert checked >= 12 * 10 + 9


@pytest.mark.parametrize('causal', [1, 0, 3])
def test_drop_path_row_scales_follow_the_reference_granularity(causal, logger):
    """Stochastic depth draws one Bernoulli per (b h w) sequence / (b t) frame / sample b (vit.py:170-172, 181-186, 216;
    vit_utils.py:139-164) and the cls rows follow vit.py:191-198."""
    from tcow_b200.train_engine import SeekerTrainEngine
    mod = _small_tracker(logger, causal=causal).train()
    eng = SeekerTrainEngine(mod)
    B, N, T = 2, 6, 4
    M, use_cls = B * N * T, causal in (0, 1)
    R = M + B
    Rs = R if use_cls else M
    keep = 0.75
    ov = [None] + [dict(keep=keep, t=(torch.arange(B * N) % 3 != 0).float(), s=(torch.arange(B * T) % 2 == 0).float(),
                        m=torch.tensor([1.0, 0.0])) for _ in range(11)]
    eng.drop_path_override = ov
    sc = eng._drop_path_scales(mod, 12, B, N, T, M, R, Rs, use_cls, causal, torch.device('cpu'))
    assert sc[0] is None
    d = sc[5]
    rows = torch.arange(M)
    b, n, t = rows // (N * T), (rows // T) % N, rows % T
    assert torch.equal(d['rs_t'], ov[5]['t'][b * N + n] / keep)                     # temporal: per (b, n) sequence
    assert torch.equal(d['rs_s'][:M], ov[5]['s'][b * T + t] / keep)                  # spatial: per (b, t) frame
    assert torch.equal(d['rs_m'][:M], ov[5]['m'][b] / keep) and torch.equal(d['rs_m'][M:], ov[5]['m'] / keep)
    if causal == 1:      # cls residual = frame 0's output only
        assert torch.equal(d['rs_s'][M:], ov[5]['s'].view(B, T)[:, 0] / keep) and torch.equal(d['bs_s'], d['rs_s'])
    elif causal == 0:    # mean over frames taken after DropPath: the matmul part is pre-weighted, the bias sees the mean scale
        assert torch.equal(d['rs_s'][M:], torch.ones(B))
        assert torch.allclose(d['bs_s'][M:], (ov[5]['s'] / keep).view(B, T).mean(1))
    else:
        assert d['rs_s'].numel() == M
    # drawn masks: block i keeps with probability 1 - linspace(0, rate, 12)[i]
    eng.drop_path_override = None
    torch.manual_seed(0)
    drawn = eng._drop_path_scales(mod, 12, 64, N, T, 64 * N * T, 64 * N * T + 64, 64 * N * T + 64, True, 1, torch.device('cpu'))
    assert drawn[0] is None
    frac_kept = (drawn[11]['rs_t'] > 0).float().mean().item()
    vals = drawn[11]['rs_t'].unique()
    assert abs(frac_kept - 0.7) < 0.08 and all(min(abs(v), abs(v - 1.0 / 0.7)) < 1e-5 for v in vals.tolist())
