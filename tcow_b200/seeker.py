"""Drop-in for model/seeker.py: ``Seeker(logger, **seeker_args)`` with ``forward(input_frames, query_mask)``
(model/seeker.py:17-25).  State-dict keys carry the same ``seeker.`` prefix as the reference."""
from __future__ import annotations

import torch

from . import mask_tracker


class Seeker(torch.nn.Module):

    def __init__(self, logger, **kwargs):
        super().__init__()
        self.logger = logger
        self.seeker = mask_tracker.QueryMaskTracker(logger, **kwargs)

    def forward(self, *args, **kwargs):
        return self.seeker(*args, **kwargs)

    def forward_queries(self, input_frames, query_masks, frame_scale=1.0):
        '''(B,3,T,Hf,Wf), (B,Qs,1,T,Hf,Wf) -> ((B,Qs,C,T,Hf,Wf), (B,Qs,T,F)): the per-query loop of
        pipeline.py:134-182 as one batched pass over shared frames.'''
        return self.seeker.forward_queries(input_frames, query_masks, frame_scale=frame_scale)
