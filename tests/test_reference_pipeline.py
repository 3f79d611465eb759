"""The reference's own evaluation driver running on the drop-in module (VERDICT r01 "missing" #5).

`eval/inference.py:60-93 perform_inference` builds `pipeline.MyTrainPipeline` (pipeline.py:15-47), runs
`forward -> forward_plugin` (pipeline.py:49-82, :200-239: `self.networks['seeker'](seeker_input, seeker_query_mask)`),
the metrics of loss.py / eval/metrics.py and `process_entire_batch`, then moves everything to the host.  Here that code —
imported unmodified from /root/reference — drives `tcow_b200.Seeker` and, for comparison, the reference Seeker with the same
weights.  There is no GPU in the build container, so the drop-in's arithmetic is supplied by the fp32 oracle through the
engine seam (`QueryMaskTracker._engine`); everything the reference touches — constructor kwargs, ModuleDict registration,
`.eval()` / `set_grad_enabled`, the forward signature and return types, state-dict loading — is the product code.  Skipped
where /root/reference does not exist (the GPU box)."""
import argparse
import logging
import os
import sys

import pytest
import torch

import tcow_b200
from conftest import cached_state_dict
from oracle import ref_import, seeker_oracle
from tcow_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ref_import.REF, 'model')),
                                reason='reference tree not available (GPU box)')

T, HF, WF = 4, 32, 48
KW = dict(num_total_frames=T, num_visible_frames=T, frame_height=HF, frame_width=WF, tracker_pretrained=False,
          attention_type='divided_space_time', patch_size=16, causal_attention=1, norm_embeddings=False, drop_path_rate=0.1,
          network_depth=12, track_map_stride=4, track_map_resize='bilinear', query_channels=1, output_channels=3,
          flag_channels=3)


class OracleEngine:
    """Stands in for SeekerEngine on a machine without a GPU: same call signature, arithmetic by the fp32 oracle."""

    def forward(self, mod, input_frames, query_mask, queries_per_video=1, frame_scale=1.0):
        sd = {'seeker.' + k: v for k, v in mod.state_dict().items()}
        frames = input_frames.float() * frame_scale
        if queries_per_video > 1:
            frames = frames.repeat_interleave(queries_per_video, 0)
        return seeker_oracle.seeker_forward(sd, frames, query_mask, causal_attention=int(mod.causal_attention),
                                            norm_embeddings=mod.norm_embeddings, pretrained_norm=mod.tracker_backbone.pretrained,
                                            track_map_resize=mod.track_map_resize, flag_channels=mod.flag_channels)


class Logger(logging.Logger):
    def __init__(self):
        super().__init__('pipeline_test')
        self.scalars = {}

    def report_scalar(self, key, value, **kw):
        self.scalars[key] = value


@pytest.fixture(scope='module')
def reference_modules():
    ref_import.import_reference_seeker()
    REF = ref_import.REF
    cwd = os.getcwd()
    os.chdir(REF)
    saved = list(sys.path)
    sys.path[:0] = [REF, REF + '/eval', REF + '/utils', REF + '/data', REF + '/model']
    try:
        import inference
        import pipeline
        yield inference, pipeline
    finally:
        os.chdir(cwd)
        sys.path[:] = saved


def make_batch():
    rgb, q = synth.make_batch([0, 1], num_frames=T, frame_height=HF, frame_width=WF)
    tm, _ = synth.make_targets([0, 1], num_frames=T, frame_height=HF, frame_width=WF)
    return {'source_name': ['plugin', 'plugin'], 'within_batch_idx': torch.arange(2), 'pv_rgb_tf': rgb,
            'pv_query_tf': q.to(torch.uint8), 'pv_target_tf': tm.to(torch.int8)}


def test_perform_inference_runs_on_the_drop_in(reference_modules):
    inference, pipeline = reference_modules
    logger = Logger()
    sd = cached_state_dict(901, T, HF, WF)
    ours = tcow_b200.Seeker(logger, **KW)
    ours.load_state_dict(sd)
    ours.seeker._engine = OracleEngine()
    ref = ref_import.build_reference(sd, **KW)
    train_args = argparse.Namespace(num_frames=T, num_queries=1, batch_size=2)
    all_args = {'train': train_args, 'test': argparse.Namespace(num_queries=1)}
    out = {}
    for name, net in (('ours', ours), ('ref', ref)):
        r = inference.perform_inference(make_batch(), {'seeker': net}, torch.device('cpu'), logger, all_args, 0)
        assert torch.is_grad_enabled() is False          # set_phase('test') (pipeline.py:44-47)
        torch.set_grad_enabled(True)
        out[name] = r
        assert not net.training
    mo, mr = out['ours']['model_retval'], out['ref']['model_retval']
    assert set(mo) == set(mr) == {'seeker_input', 'seeker_query_mask', 'target_mask', 'output_mask', 'output_flags'}
    assert mo['output_mask'].shape == (2, 3, T, HF, WF) and mo['output_flags'].shape == (2, T, 3)
    assert abs(mo['output_mask'] - mr['output_mask']).max() < 5e-6 and abs(mo['output_flags'] - mr['output_flags']).max() < 5e-6
    lo, lr = out['ours']['loss_retval']['metrics'], out['ref']['loss_retval']['metrics']
    assert set(lo) == set(lr) and len(lo) == 12
    for k in lo:
        assert abs(float(lo[k]) - float(lr[k])) < 1e-6, k


def test_train_phase_of_the_reference_pipeline_reaches_the_training_engine(reference_modules):
    """set_phase('train') (pipeline.py:37-41) puts the drop-in in train mode with gradients on: the forward dispatches to
    the hand-written training plan (which needs the GPU — here it must fail loudly, not fall back)."""
    _, pipeline = reference_modules
    logger = Logger()
    ours = tcow_b200.Seeker(logger, **KW)
    p = pipeline.MyTrainPipeline(argparse.Namespace(num_frames=T, num_queries=1), logger, {'seeker': ours}, torch.device('cpu'))
    p.set_phase('train')
    try:
        assert ours.training and torch.is_grad_enabled()
        assert len(list(p.parameters())) == 251              # what train.py:241 hands to the optimizer
        with pytest.raises(RuntimeError, match='CUDA'):
            p.forward_plugin(make_batch())
    finally:
        torch.set_grad_enabled(True)
