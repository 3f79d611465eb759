"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
For every case: build the seeded 251-tensor state dict (tcow_b200.synth), load it
into the reference ``Seeker`` (model/seeker.py:17), run its forward on seeded clips,
and store the outputs.  The oracle restatement is checked against the same run
(printed), and tests/test_oracle.py re-checks it against the stored vectors.
Full-size outputs are stored on a strided pixel lattice to keep fixtures small.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tcow_b200 import synth  # noqa: E402
from oracle import ref_import, seeker_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    # name, T, Hf, Wf, samples, causal, extra
    dict(name='small_causal1', T=4, Hf=32, Wf=48, samples=[0, 1], causal=1),
    dict(name='small_causal0', T=5, Hf=48, Wf=32, samples=[2], causal=0),
    dict(name='small_causal2', T=3, Hf=32, Wf=32, samples=[3], causal=2),
    dict(name='small_causal3', T=6, Hf=32, Wf=32, samples=[4], causal=3),
    dict(name='small_causalm1', T=4, Hf=32, Wf=32, samples=[5], causal=-1),
    dict(name='small_norm_pretr', T=4, Hf=32, Wf=48, samples=[6], causal=1, norm_embeddings=True,
         pretrained_norm=True),
    dict(name='small_nearest_noflags', T=4, Hf=32, Wf=48, samples=[7], causal=1,
         track_map_resize='nearest', flag_channels=0),
    dict(name='mid_causal1', T=30, Hf=64, Wf=96, samples=[8, 9], causal=1, query_frame=3, lattice=(3, 5)),
    dict(name='full_causal1', T=30, Hf=240, Wf=320, samples=[0, 1], causal=1, lattice=(7, 9)),
    # more than 304 tokens per frame and more than 32 frames: the long-sequence kernels (config 5 in miniature)
    dict(name='long_causal0', T=34, Hf=256, Wf=320, samples=[12], causal=0, lattice=(5, 7)),
    # BASELINE configs[4] at its real shape: S=1201 tokens per frame, cls mean over 60 frames (vit.py:195); ~2 min on CPU
    dict(name='hires_causal0', T=60, Hf=480, Wf=640, samples=[20], causal=0, lattice=(9, 11)),
    # clips 0 and 7 of the batch bench.py times on rank 0 (synth.bench_clips): the bench asserts them after the timed loop
    dict(name='full_bench_b8', T=30, Hf=240, Wf=320, samples=[0, 7], bench_batch=8, causal=1, lattice=(7, 9)),
]
WEIGHT_SEED = 901


def run_case(c):
    T, Hf, Wf = c['T'], c['Hf'], c['Wf']
    flag_channels = c.get('flag_channels', 3)
    sd = synth.make_state_dict(WEIGHT_SEED, num_frames=T, frame_height=Hf, frame_width=Wf,
                               flag_channels=flag_channels)
    kwargs = dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf,
                  tracker_pretrained=False, attention_type='divided_space_time', patch_size=16,
                  causal_attention=c['causal'], norm_embeddings=c.get('norm_embeddings', False),
                  drop_path_rate=0.1, network_depth=12, track_map_stride=4,
                  track_map_resize=c.get('track_map_resize', 'bilinear'), query_channels=1,
                  output_channels=3, flag_channels=flag_channels)
    net = ref_import.build_reference(sd, **kwargs)
    if c.get('pretrained_norm'):
        net.seeker.tracker_backbone.pretrained = True      # SURVEY §8(c) trap 5
    if 'bench_batch' in c:
        rgb, q = synth.bench_clips(0, c['bench_batch'], T, Hf, Wf)
        rgb, q = rgb[c['samples']], q[c['samples']]
    else:
        rgb, q = synth.make_batch(c['samples'], num_frames=T, frame_height=Hf, frame_width=Wf,
                                  query_frame=c.get('query_frame', 0))
    t0 = time.time()
    with torch.no_grad():
        mask, flags = net(rgb.clone(), q.clone())
    t_ref = time.time() - t0
    t0 = time.time()
    with torch.no_grad():
        omask, oflags = seeker_oracle.seeker_forward(
            sd, rgb, q, causal_attention=c['causal'], norm_embeddings=c.get('norm_embeddings', False),
            pretrained_norm=c.get('pretrained_norm', False),
            track_map_resize=c.get('track_map_resize', 'bilinear'), flag_channels=flag_channels)
    t_or = time.time() - t0
    em = (mask - omask).abs().max().item()
    ef = (flags - oflags).abs().max().item() if flags is not None else 0.0
    print(f"{c['name']:24s} ref {t_ref:6.2f}s oracle {t_or:6.2f}s  |mask|max {mask.abs().max():.3f} "
          f"std {mask.std():.3f}  oracle-vs-ref mask {em:.2e} flags {ef:.2e}", flush=True)
    assert em < 5e-5 and ef < 5e-5, 'oracle restatement disagrees with the reference'
    ly, lx = c.get('lattice', (1, 1))
    meta = {k: v for k, v in c.items()}
    meta.update(weight_seed=WEIGHT_SEED, ref_kwargs=kwargs, torch=torch.__version__,
                mask_absmax=float(mask.abs().max()), mask_std=float(mask.std()))
    arrs = dict(mask=mask[:, :, :, ::ly, ::lx].numpy().astype(np.float32),
                meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    if flags is not None:
        arrs['flags'] = flags.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, c['name'] + '.npz'), **arrs)


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for c in CASES:
        if only and c['name'] not in only:
            continue
        run_case(c)
