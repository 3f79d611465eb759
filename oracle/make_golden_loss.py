"""TEST INFRASTRUCTURE ONLY — pin the loss / metrics restatement and the CUDA loss against the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden_loss
Seeded synthetic logits / targets / occlusion data go through the reference's own `MyLosses` (loss.py) and
`calculate_metrics_mask_track` (eval/metrics.py): frame weights, pixel weights, `my_mask_loss` for the three channels at
several training-progress values (top-k fractions), with and without focal loss, the weighted total and its gradient with
respect to the logits.  Stored in tests/golden/loss_*.npz; the inputs are rebuilt from the seed by `make_inputs`.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    dict(name='loss_default_early', seed=1, B=2, Q=2, T=4, H=24, W=32, progress=0.02, query_time=0),
    dict(name='loss_default_late', seed=2, B=2, Q=2, T=4, H=24, W=32, progress=0.6, query_time=1),
    dict(name='loss_focal', seed=3, B=1, Q=3, T=3, H=32, W=48, progress=0.05, query_time=0, focal_loss=True),
    dict(name='loss_no_aot_no_balance', seed=4, B=2, Q=1, T=3, H=24, W=32, progress=0.3, query_time=2, aot_loss=0.0,
         class_balancing=False, hard_negative_factor=1.0),
    # channels without any occluder / container anywhere and occl_cont_zero_weight = 0: frames are dropped (which_frames)
    dict(name='loss_frame_selection', seed=5, B=2, Q=2, T=5, H=24, W=32, progress=0.1, query_time=0, occl_cont_zero_weight=0.0,
         empty_frames=True),
]


def make_inputs(c):
    """Deterministic stand-ins for the Kubric tensors of pipeline.py:184-226: logits, {0,1} targets (snitch / frontmost
    occluder / outermost container), soft occlusion fractions, occlusion pointers."""
    g = torch.Generator().manual_seed(1000 + c['seed'])
    B, Q, T, H, W = c['B'], c['Q'], c['T'], c['H'], c['W']
    output_mask = torch.randn(B, Q, 3, T, H, W, generator=g) * 2.0
    target = torch.zeros(B, Q, 3, T, H, W)
    for b in range(B):
        for q in range(Q):
            for ch in range(3):
                for t in range(T):
                    if c.get('empty_frames') and ch > 0 and (t + b + q) % 2 == 0:
                        continue
                    h, w = H // (3 + ch), W // (3 + ch)
                    y0 = int(torch.randint(0, H - h + 1, (1,), generator=g))
                    x0 = int(torch.randint(0, W - w + 1, (1,), generator=g))
                    target[b, q, ch, t, y0:y0 + h, x0:x0 + w] = 1.0
    sel_occl_fracs = torch.rand(B, Q, T, 3, generator=g)
    occl_ptr = ((torch.rand(B, Q, 1, T, H, W, generator=g) > 0.7) & (target[:, :, 0:1] > 0.5)).to(torch.uint8) * 3
    return output_mask, target, sel_occl_fracs, occl_ptr


def train_args(c):
    return argparse.Namespace(track_lw=1.0, occl_mask_lw=0.5, cont_mask_lw=0.5, occluded_weight=5.0,
                              occl_cont_zero_weight=c.get('occl_cont_zero_weight', 0.02),
                              class_balancing=c.get('class_balancing', True), focal_loss=c.get('focal_loss', False),
                              aot_loss=c.get('aot_loss', 0.8), hard_negative_factor=c.get('hard_negative_factor', 3.0))


def run_case(c, loss_mod):
    args = train_args(c)
    L = loss_mod.MyLosses(args, logging.getLogger('x'), 'train')
    out, tgt, fracs, occl = make_inputs(c)
    out = out.clone().requires_grad_(True)
    data_retval = {'source_name': ['kubric'],
                   'kubric_retval': {'pv_rgb_tf': torch.zeros(c['B'], 3, c['T'], c['H'], c['W']),
                                     'traject_retval_tf': {'query_time': torch.tensor([c['query_time']] * c['B'])}}}
    model_retval = {'target_mask': tgt, 'output_mask': out, 'sel_occl_fracs': fracs, 'snitch_occl_by_ptr': occl}
    r = L.per_example_mask_track(data_retval, model_retval, c['progress'], False)
    total = r['track'] * args.track_lw + r['occl_mask'] * args.occl_mask_lw + r['cont_mask'] * args.cont_mask_lw
    total.backward()
    fw = L.get_mask_track_frame_weights(fracs, c['query_time'])
    pw = L.get_mask_track_pixel_weights(fracs, tgt[:, :, 0], occl[:, :, 0])
    arrs = dict(track=float(r['track']), occl_mask=float(r['occl_mask']), cont_mask=float(r['cont_mask']), total=float(total),
                grad=out.grad.numpy().astype(np.float32), frame_weights=fw.numpy(), pixel_weights=pw.numpy(),
                snitch_weights=model_retval['snitch_weights'].numpy())
    for k, v in r['metrics'].items():
        arrs['metric_' + k] = np.array(v.item())
    arrs['meta'] = np.frombuffer(json.dumps(dict(c, torch=torch.__version__)).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, c['name'] + '.npz'), **arrs)
    print(f"{c['name']:26s} track {arrs['track']:.6f} occl {arrs['occl_mask']:.6f} cont {arrs['cont_mask']:.6f} "
          f"total {arrs['total']:.6f} |grad| {np.abs(arrs['grad']).sum():.4e}", flush=True)


if __name__ == '__main__':
    ref_import.import_reference_seeker()          # registers the stub modules `from __init__ import *` needs
    REF = ref_import.REF
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path[:0] = [REF, REF + '/eval', REF + '/utils', REF + '/data', REF + '/model']
    try:
        import loss as ref_loss
    finally:
        os.chdir(cwd)
    os.makedirs(OUT, exist_ok=True)
    for c in CASES:
        run_case(c, ref_loss)
