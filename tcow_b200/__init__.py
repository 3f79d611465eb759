"""tcow_b200 — B200-native (sm_100a) drop-in for the Seeker forward of basilevh/tcow.

Public surface mirrors the reference modules of the hot path:
``Seeker`` (model/seeker.py), ``QueryMaskTracker`` (model/mask_tracker.py),
``MyDenseTimeSformerBackbone`` (model/vision_tf.py).  Importing the package does not need a GPU;
running a forward does, and there is no fallback.
"""
from .seeker import Seeker
from .mask_tracker import QueryMaskTracker
from .vision_tf import DenseTimeSformer, MyDenseTimeSformerBackbone

__all__ = ['Seeker', 'QueryMaskTracker', 'DenseTimeSformer', 'MyDenseTimeSformerBackbone']
