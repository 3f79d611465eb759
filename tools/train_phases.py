"""Where a training step's time goes outside the engine's kernels: CUDA-event timing of the phases of one step."""
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tcow_b200  # noqa: E402
from tcow_b200 import synth  # noqa: E402

T, Hf, Wf, V, Q = 30, 240, 320, 2, 3
dev = 'cuda:0'
net = tcow_b200.Seeker(logging.getLogger('p'), num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                       tracker_pretrained=False, causal_attention=1, patch_size=16, drop_path_rate=0.1)
net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=Hf, frame_width=Wf))
net = net.to(dev).train()
rgb = torch.rand(V, 3, T, Hf, Wf, device=dev)
q = torch.zeros(V, Q, 1, T, Hf, Wf, device=dev)
q[:, :, 0, 0, 10:50, 20:60] = 1
tm = (torch.rand(V * Q, 3, T, Hf, Wf, device=dev) > 0.7).float()
tf = (torch.rand(V * Q, T, 3, device=dev) > 0.5).float()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
eng = net.seeker.train_engine()
orig_fwd, orig_bwd, orig_unpack, orig_pack = eng.forward, eng.backward, eng._unpack, eng._pack
marks = []


def ev(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def wrap(fn, name):
    def f(*a, **k):
        ev(name + ':begin')
        r = fn(*a, **k)
        ev(name + ':end')
        return r
    return f


eng.forward, eng.backward, eng._unpack, eng._pack = wrap(orig_fwd, 'engine.forward'), wrap(orig_bwd, 'engine.backward'), \
    wrap(orig_unpack, 'unpack'), wrap(orig_pack, 'pack')
for it in range(4):
    marks.clear()
    ev('step:begin')
    opt.zero_grad(set_to_none=True)
    mask, flags = net.forward_queries(rgb, q)
    ev('loss:begin')
    loss = synth.training_loss(mask.flatten(0, 1), flags.flatten(0, 1), tm, tf)
    ev('loss:end')
    loss.backward()
    ev('backward:end')
    opt.step()
    ev('step:end')
    torch.cuda.synchronize()
t0 = marks[0][1]
prev = t0
for name, e in marks:
    print(f'{name:24s} t={t0.elapsed_time(e):8.3f} ms   (+{prev.elapsed_time(e):7.3f})')
    prev = e
