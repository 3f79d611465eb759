"""Drop-in for model/vision_tf.py: DenseTimeSformer / MyDenseTimeSformerBackbone.

Same constructor arguments and attributes as the reference (model/vision_tf.py:32-66, :172-183); the
``forward`` of the reference (:68-169) is executed by ``tcow_b200.engine.SeekerEngine`` on the GPU as
part of the fused Seeker plan, so calling this module on its own returns the dense features from that
engine (used by tests); ``QueryMaskTracker`` calls the engine end to end instead.
"""
from __future__ import annotations

import torch

from .vit import TimeSformer

TIMESFORMER_MEAN = (0.45, 0.45, 0.45)
TIMESFORMER_STD = (0.225, 0.225, 0.225)


class DenseTimeSformer(torch.nn.Module):

    def __init__(self, logger, pretrained, pretrained_path, frame_height, frame_width, patch_dim, in_channels,
                 num_frames, attention_type, causal_attention, norm_embeddings, drop_path_rate, network_depth):
        super().__init__()
        self.logger = logger
        self.pretrained = pretrained
        self.Hf, self.Wf = frame_height, frame_width
        self.Ho, self.Wo = frame_height // patch_dim, frame_width // patch_dim
        self.ho = self.wo = patch_dim
        self.Ci = in_channels
        self.T = num_frames
        self.attention_type = attention_type
        self.causal_attention = causal_attention
        self.norm_embeddings = norm_embeddings
        self.drop_path_rate = drop_path_rate
        self.network_depth = network_depth
        self.timesformer = TimeSformer(
            img_size=(self.Hf, self.Wf), patch_size=patch_dim, num_frames=self.T,
            attention_type=attention_type, causal_attention=causal_attention, drop_path_rate=drop_path_rate,
            network_depth=network_depth, pretrained=pretrained, pretrained_model=pretrained_path,
            in_chans=self.Ci)
        self.output_feature_dim = self.timesformer.model.embed_dim

    def forward(self, input_pixels, extra_token_in):
        raise RuntimeError('DenseTimeSformer.forward is fused into the Seeker CUDA plan; call Seeker / '
                           'QueryMaskTracker (tcow_b200.engine runs vision_tf.py:68-169 on the GPU)')


class MyDenseTimeSformerBackbone(DenseTimeSformer):

    def __init__(self, logger, num_frames=16, frame_height=224, frame_width=288, patch_dim=16, in_channels=3,
                 pretrained=False, pretrained_path='', attention_type='divided_space_time',
                 causal_attention=False, norm_embeddings=False, drop_path_rate=0.1, network_depth=12):
        super().__init__(logger, pretrained, pretrained_path, frame_height, frame_width, patch_dim, in_channels,
                         num_frames, attention_type, causal_attention, norm_embeddings, drop_path_rate,
                         network_depth)
