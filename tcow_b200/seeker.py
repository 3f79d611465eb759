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

    def forward(self, *args):
        return self.seeker(*args)
