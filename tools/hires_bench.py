"""BASELINE configs[4]: high-resolution stress — T=60, 480x640 (1200 patches per frame, 72 000 tokens per clip),
non-causal temporal attention (cls mean path), bf16, batch 1 per pass.  Prints clips/s and the kernel breakdown."""
import collections
import json
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tcow_b200  # noqa: E402
from tcow_b200 import synth  # noqa: E402

T, Hf, Wf = 60, 480, 640
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = 'cuda:0'
net = tcow_b200.Seeker(logging.getLogger('hires'), num_total_frames=T, num_visible_frames=T, frame_height=Hf,
                       frame_width=Wf, tracker_pretrained=False, causal_attention=0, patch_size=16)
net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=Hf, frame_width=Wf))
net = net.to(dev).eval()
rgb, q = synth.make_batch([0], num_frames=T, frame_height=Hf, frame_width=Wf)
rgb, q = rgb.to(dev), q.to(dev)
eng = net.seeker.engine()
with torch.no_grad():
    for _ in range(3):
        net(rgb, q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        net(rgb, q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.profile = []
    net(rgb, q)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for kind, fl, nb, a, b in eng.profile:
    d = agg.setdefault(kind, [0, 0.0, 0.0])
    d[0] += 1; d[1] += a.elapsed_time(b); d[2] += fl
FLOP = 20878.31e9   # SURVEY.md §8d, equals FlopCounterMode on the reference
print(json.dumps({'metric': 'seeker_fwd_clips_per_s_hires', 'value': 1e3 / ms, 'unit': 'clips/s', 'ms_per_clip': ms,
                  'config': 'T=60 480x640 causal_attention=0 bf16 B=1 (BASELINE configs[4])',
                  'tflops_algorithmic': FLOP / ms / 1e9, 'frac_of_burst_peak_1659.9': FLOP / ms / 1e9 / 1659.9,
                  'breakdown_ms': {k: round(v[1], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}))
