"""Drop-in for model/mask_tracker.py: QueryMaskTracker with the reference's constructor, parameters and
forward contract (model/mask_tracker.py:24-28, :92-142); the arithmetic runs on sm_100a CUDA kernels."""
from __future__ import annotations

import torch

from . import vision_tf
from .engine import SeekerEngine
from .train_engine import SeekerFunction, SeekerTrainEngine


def _parse_tracker_pretrained(value):
    """(use_pretrained, path) from a bool or the string forms args.py:150 allows (mask_tracker.py:52-67)."""
    if isinstance(value, bool):
        return value, ''
    if not isinstance(value, str):
        raise ValueError(f'Invalid tracker_pretrained value: {value}.')
    if value.lower() in ('1', 'y', 'yes', 't', 'true'):
        return True, ''
    return (False, '') if len(value) <= 5 else (True, value)


class QueryMaskTracker(torch.nn.Module):

    def __init__(self, logger, num_total_frames=24, num_visible_frames=16, frame_height=224, frame_width=288,
                 tracker_pretrained=False, attention_type='divided_space_time', patch_size=16,
                 causal_attention=False, norm_embeddings=False, drop_path_rate=0.1, network_depth=12,
                 track_map_stride=4, track_map_resize='bilinear', query_channels=1, output_channels=3,
                 flag_channels=3):
        super().__init__()
        self.logger = logger
        # every constructor argument becomes an attribute of the same name, as callers of the reference expect
        # (mask_tracker.py:30-50); tracker_pretrained is normalised to (bool, path) by the rule of mask_tracker.py:52-67
        for name, value in dict(num_total_frames=num_total_frames, num_visible_frames=num_visible_frames,
                                frame_height=frame_height, frame_width=frame_width, attention_type=attention_type,
                                patch_size=patch_size, causal_attention=causal_attention,
                                norm_embeddings=norm_embeddings, drop_path_rate=drop_path_rate,
                                network_depth=network_depth, track_map_stride=track_map_stride,
                                track_map_resize=track_map_resize, query_channels=query_channels,
                                output_channels=output_channels, flag_channels=flag_channels).items():
            setattr(self, name, value)
        self.input_channels = 3 + query_channels
        self.tracker_pretrained, self.pretrained_path = _parse_tracker_pretrained(tracker_pretrained)
        logger.info(f'(QueryMaskTracker) tracker_pretrained: {self.tracker_pretrained} '
                    f'pretrained_path: {self.pretrained_path}')
        if query_channels != 1:
            raise NotImplementedError('tcow_b200 supports query_channels=1 (train.py:202 hard-codes it)')
        if frame_height % patch_size or frame_width % patch_size:
            raise AssertionError('frame size must be a multiple of the patch size (mask_tracker.py:89-90)')

        self.tracker_backbone = vision_tf.MyDenseTimeSformerBackbone(
            logger, num_frames=num_total_frames, frame_height=frame_height, frame_width=frame_width,
            patch_dim=patch_size, in_channels=self.input_channels, pretrained=self.tracker_pretrained,
            pretrained_path=self.pretrained_path, attention_type=attention_type, causal_attention=causal_attention,
            norm_embeddings=norm_embeddings, drop_path_rate=drop_path_rate, network_depth=network_depth)
        self.use_feature_dim = self.tracker_backbone.output_feature_dim
        # parameter names below are part of the checkpoint contract (251-tensor state dict)
        self.tracker_post_linear = torch.nn.Linear(self.use_feature_dim, output_channels * patch_size * patch_size)
        if flag_channels > 0:
            self.flag_post_linear = torch.nn.Linear(self.use_feature_dim, flag_channels)
        self._engine = None  # built lazily on the parameters' device; not part of the state dict
        self._train_engine = None

    def engine(self):
        if self._engine is None:
            self._engine = SeekerEngine(self)
        return self._engine

    def train_engine(self):
        if self._train_engine is None:
            self._train_engine = SeekerTrainEngine(self)
        return self._train_engine

    def _wants_grad(self):
        if not torch.is_grad_enabled():
            return False
        if any(p.requires_grad for p in self.parameters()):
            return True
        # nn.DataParallel replicas (train.py:223) report no parameters() (replicate() keeps the broadcast copies in
        # _former_parameters): the autograd node of train_engine.py could not hand gradients back to the wrapped module's
        # leaves, and the inference plan would return outputs without a grad_fn — fail here, not in loss.backward().
        replica = getattr(self, '_is_replica', False) and any(
            t is not None and t.requires_grad for m in self.modules() for t in getattr(m, '_former_parameters', {}).values())
        if replica and self.training:
            raise RuntimeError(
                'tcow_b200.Seeker was called in gradient mode as an nn.DataParallel replica (train.py:222-223). Training '
                'runs one process per GPU: launch with torchrun, call tcow_b200.ddp.broadcast_parameters(net) and '
                'tcow_b200.ddp.attach(net) instead of wrapping the module in nn.DataParallel (INTEGRATION.md).')
        return False

    def _forward_train(self, input_frames, query_mask, queries_per_video):
        # autograd node over the hand-written backward (train_engine.py); gradients flow to the parameters only —
        # the reference never differentiates w.r.t. its inputs either (train.py:93-101)
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        (mask, flags) = SeekerFunction.apply(self.train_engine(), self, names, queries_per_video, input_frames,
                                             query_mask, *params)
        return (mask, flags if self.flag_channels > 0 else None)

    def forward(self, input_frames, query_mask, frame_scale=1.0):
        '''
        :param input_frames (B, 3, T, Hf, Wf) tensor.
        :param query_mask (B, 1, T, Hf, Wf) tensor.
        :param frame_scale (float): Extension over the reference (default = its behaviour): factor applied to the RGB
            values on the device; 1/255 lets a caller pass the decoder's uint8 frames and skip the host-side
            `rgb / 255.0` of data/data_plugin.py:174 (4x less host->device traffic).
        :return (output_mask (B, C, T, Hf, Wf) fp32 logits, output_flags (B, T, F) fp32 or None).
        '''
        assert query_mask.shape[1] == 1                         # mask_tracker.py:105
        if self._wants_grad():
            if frame_scale != 1.0:
                input_frames = input_frames.to(torch.float32) * frame_scale
            return self._forward_train(input_frames, query_mask, 1)
        return self.engine().forward(self, input_frames, query_mask, frame_scale=frame_scale)

    def forward_queries(self, input_frames, query_masks, frame_scale=1.0):
        '''
        All Qs queries of every clip in one pass — what pipeline.py:134-182 obtains by calling forward() Qs times
        on the same frames and stacking.  The RGB frames are uploaded / read once per clip, not once per query.
        :param input_frames (B, 3, T, Hf, Wf) tensor.
        :param query_masks (B, Qs, 1, T, Hf, Wf) tensor.
        :return (output_mask (B, Qs, C, T, Hf, Wf), output_flags (B, Qs, T, F) or None).
        '''
        assert query_masks.dim() == 6 and query_masks.shape[2] == 1
        (B, Qs) = query_masks.shape[:2]
        assert input_frames.shape[0] == B
        flat_q = query_masks.reshape(B * Qs, *query_masks.shape[2:])
        if self._wants_grad():
            if frame_scale != 1.0:
                input_frames = input_frames.to(torch.float32) * frame_scale
            (mask, flags) = self._forward_train(input_frames, flat_q, Qs)
        else:
            (mask, flags) = self.engine().forward(self, input_frames, flat_q, queries_per_video=Qs,
                                                  frame_scale=frame_scale)
        mask = mask.reshape(B, Qs, *mask.shape[1:])
        if flags is not None:
            flags = flags.reshape(B, Qs, *flags.shape[1:])
        return (mask, flags)
