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
