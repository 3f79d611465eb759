"""TEST INFRASTRUCTURE ONLY — gradient fixtures from the UNMODIFIED reference's autograd.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_grads
For every case: seeded state dict -> reference ``Seeker`` in train() mode with drop_path_rate=0 (stochastic depth draws
from the global RNG) -> synth.training_loss -> loss.backward().  A full gradient is 122 M floats, so
the fixture keeps, per parameter tensor: its L2 norm, 256 entries at seeded indices and 4 seeded random projections.
The oracle restatement differentiated by torch.autograd is checked against the same run here (printed) and in
tests/test_oracle.py; the CUDA backward is then compared with the oracle's FULL gradients and with these fixtures.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tcow_b200 import synth  # noqa: E402
from oracle import ref_import, seeker_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
WEIGHT_SEED = 901
NSAMP, NPROJ = 256, 4

CASES = [
    dict(name='grad_small_causal1', T=4, Hf=32, Wf=48, samples=[0, 1], causal=1),
    dict(name='grad_small_causal0', T=5, Hf=48, Wf=32, samples=[2], causal=0),
    dict(name='grad_small_causal3', T=6, Hf=32, Wf=32, samples=[4], causal=3),
    dict(name='grad_small_norm_nearest', T=4, Hf=32, Wf=48, samples=[6, 7], causal=1, norm_embeddings=True,
         pretrained_norm=True, track_map_resize='nearest'),
    # stochastic depth ON in the reference (train mode, global RNG seeded): the Bernoulli draws of DropPath are
    # replayed from the same seed and stored, so that the oracle and the CUDA path can be given the same masks
    dict(name='grad_small_droppath1', T=4, Hf=32, Wf=48, samples=[0, 1, 3], causal=1, drop_path_rate=0.6, rng_seed=7),
    dict(name='grad_small_droppath0', T=5, Hf=32, Wf=32, samples=[2, 5], causal=0, drop_path_rate=0.6, rng_seed=11),
]


def replay_drop_path(c, B, N, T, depth=12):
    """The keep masks the reference draws (vit_utils.py:150-152: floor(keep + rand)) for a forward run right after
    torch.manual_seed(c['rng_seed']): per block i >= 1 (block 0 is nn.Identity, vit.py:149), in execution order
    temporal (B*N,1,1), spatial (B*T,1,1), mlp (B,1,1)."""
    rate = c.get('drop_path_rate', 0.0)
    if rate <= 0:
        return None
    torch.manual_seed(c['rng_seed'])
    rates = [x.item() for x in torch.linspace(0, rate, depth)]
    out = []
    for i in range(depth):
        if rates[i] <= 0:
            out.append(None)
            continue
        keep = 1 - rates[i]
        d = {'keep': keep}
        for which, n in (('t', B * N), ('s', B * T), ('m', B)):
            d[which] = (keep + torch.rand((n, 1, 1), dtype=torch.float32)).floor_().reshape(-1)
        out.append(d)
    return out


def summarize(name, grad):
    """(norm, sampled entries, projections) of one gradient tensor — seeded by the parameter name."""
    g = grad.detach().reshape(-1).to(torch.float64)
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b'\0'), 'little') % (2 ** 31) + g.numel() % 9973
    gen = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, g.numel(), (NSAMP,), generator=gen)
    proj = torch.stack([(g * torch.randn(g.numel(), generator=gen, dtype=torch.float64)).sum() for _ in range(NPROJ)])
    return float(g.norm()), g[idx].to(torch.float32).numpy(), proj.to(torch.float32).numpy()


def oracle_grads(sd, c, rgb, q, tm, tf, drop_path=None):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mask, flags = seeker_oracle.seeker_forward(
        leaves, rgb, q, causal_attention=c['causal'], norm_embeddings=c.get('norm_embeddings', False),
        pretrained_norm=c.get('pretrained_norm', False), track_map_resize=c.get('track_map_resize', 'bilinear'),
        flag_channels=c.get('flag_channels', 3), drop_path=drop_path)
    loss = synth.training_loss(mask, flags, tm, tf)
    loss.backward()
    return float(loss.detach()), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


def run_case(c):
    T, Hf, Wf = c['T'], c['Hf'], c['Wf']
    fc = c.get('flag_channels', 3)
    sd = synth.make_state_dict(WEIGHT_SEED, num_frames=T, frame_height=Hf, frame_width=Wf, flag_channels=fc)
    kwargs = dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                  tracker_pretrained=False, attention_type='divided_space_time', patch_size=16,
                  causal_attention=c['causal'], norm_embeddings=c.get('norm_embeddings', False),
                  drop_path_rate=c.get('drop_path_rate', 0.0), network_depth=12, track_map_stride=4,
                  track_map_resize=c.get('track_map_resize', 'bilinear'), query_channels=1,
                  output_channels=3, flag_channels=fc)
    net = ref_import.build_reference(sd, **kwargs).train()
    if c.get('pretrained_norm'):
        net.seeker.tracker_backbone.pretrained = True
    rgb, q = synth.make_batch(c['samples'], num_frames=T, frame_height=Hf, frame_width=Wf)
    tm, tf = synth.make_targets(c['samples'], num_frames=T, frame_height=Hf, frame_width=Wf, flag_channels=fc)
    if 'rng_seed' in c:
        torch.manual_seed(c['rng_seed'])
    mask, flags = net(rgb.clone(), q.clone())
    loss = synth.training_loss(mask, flags, tm, tf)
    loss.backward()
    ref = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}
    B, N = rgb.shape[0], (Hf // 16) * (Wf // 16)
    dp = replay_drop_path(c, B, N, T)
    oloss, og = oracle_grads(sd, c, rgb, q, tm, tf, dp)
    worst = 0.0
    for k, g in ref.items():
        rel = ((og[k] - g).norm() / g.norm().clamp_min(1e-30)).item()
        worst = max(worst, rel)
    print(f"{c['name']:26s} loss ref {float(loss):.6f} oracle {oloss:.6f}  worst per-tensor rel-L2 (oracle vs ref) {worst:.2e}",
          flush=True)
    assert abs(float(loss) - oloss) < 1e-5 and worst < 1e-3, 'oracle autograd disagrees with the reference'
    names = sorted(ref.keys())
    norms, samples, projs = [], [], []
    for k in names:
        n, s, p = summarize(k, ref[k])
        norms.append(n); samples.append(s); projs.append(p)
    meta = dict(c)
    meta.update(weight_seed=WEIGHT_SEED, ref_kwargs=kwargs, torch=torch.__version__, names=names, loss=float(loss))
    extra = {}
    if dp is not None:
        meta['drop_keep'] = [None if d is None else d['keep'] for d in dp]
        for i, d in enumerate(dp):
            if d is not None:
                for which in 'tsm':
                    extra[f'dp{i}_{which}'] = d[which].numpy().astype(np.uint8)
        print('   dropped sequences per block (t/s/m):',
              [None if d is None else tuple(int((1 - d[w]).sum()) for w in 'tsm') for d in dp])
    np.savez_compressed(os.path.join(OUT, c['name'] + '.npz'), norms=np.array(norms, dtype=np.float64),
                        samples=np.stack(samples), projs=np.stack(projs),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **extra)


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for c in CASES:
        if only and c['name'] not in only:
            continue
        run_case(c)
