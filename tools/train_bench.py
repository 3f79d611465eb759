"""Training-step timing (BASELINE configs[3] shape: T=30 240x320, B=2 videos x 3 queries per GPU): forward + backward
(+ optional SGD step), CUDA events, with a per-kernel-class breakdown from the engine's profile hooks."""
import argparse
import collections
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tcow_b200  # noqa: E402
from tcow_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--videos', type=int, default=2)
ap.add_argument('--queries', type=int, default=3)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--profile', action='store_true')
ap.add_argument('--T', type=int, default=30)
ap.add_argument('--H', type=int, default=240)
ap.add_argument('--W', type=int, default=320)
args = ap.parse_args()
T, Hf, Wf = args.T, args.H, args.W
dev = 'cuda:0'
sd = synth.make_state_dict(901, num_frames=T, frame_height=Hf, frame_width=Wf)
net = tcow_b200.Seeker(logging.getLogger('bench'), num_total_frames=T, num_visible_frames=T, frame_height=Hf,
                       frame_width=Wf, tracker_pretrained=False, causal_attention=1, patch_size=16, drop_path_rate=0.0)
net.load_state_dict(sd)
net = net.to(dev).train()
V, Q = args.videos, args.queries
rgb = torch.rand(V, 3, T, Hf, Wf, device=dev)
q = torch.zeros(V, Q, 1, T, Hf, Wf, device=dev)
q[:, :, 0, 0, 10:50, 20:60] = 1
tm = (torch.rand(V * Q, 3, T, Hf, Wf, device=dev) > 0.7).float()
tf = (torch.rand(V * Q, T, 3, device=dev) > 0.5).float()
eng = net.seeker.train_engine()


def step():
    net.zero_grad(set_to_none=True)
    mask, flags = net.forward_queries(rgb, q)
    loss = synth.training_loss(mask.flatten(0, 1), flags.flatten(0, 1), tm, tf)
    loss.backward()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
B = V * Q
flops_fwd = 2302.61e9 * B
print(f'train step: {ms:.2f} ms for {B} samples  ({B / ms * 1e3:.1f} samples/s; fwd+bwd ~3x fwd FLOPs -> '
      f'{3 * flops_fwd / ms / 1e9:.0f} TFLOP/s algorithmic)  loss {float(loss.detach()):.4f}  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
if args.profile:
    eng.profile = []
    step()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for kind, fl, nb, a, b in eng.profile:
        d = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
        d[0] += 1; d[1] += a.elapsed_time(b); d[2] += fl; d[3] += nb
    tot = sum(d[1] for d in agg.values())
    print(f'profiled kernels: {tot:.2f} ms')
    for kind, (n, t, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'  {kind:18s} n={n:4d} {t:8.3f} ms {100 * t / tot:5.1f}%  {fl / t / 1e9 if fl else 0:8.1f} TFLOP/s {nb / t / 1e6 if nb else 0:8.1f} GB/s')
    eng.profile = None
