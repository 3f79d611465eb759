"""tcow_b200 — B200-native (sm_100a) drop-in for the Seeker network of basilevh/tcow (inference and training).

Public surface mirrors the reference modules of the hot path:
``Seeker`` (model/seeker.py), ``QueryMaskTracker`` (model/mask_tracker.py),
``MyDenseTimeSformerBackbone`` (model/vision_tf.py).  Importing the package does not need a GPU;
running a forward does, and there is no fallback.  ``tcow_b200.ddp`` (data-parallel gradient exchange),
``tcow_b200.sweep`` (sharded evaluation sweep) and ``tcow_b200.checkpoint`` (reference checkpoint format) are the callers'
side of the path.
"""
from .seeker import Seeker
from .mask_tracker import QueryMaskTracker
from .vision_tf import DenseTimeSformer, MyDenseTimeSformerBackbone

__all__ = ['Seeker', 'QueryMaskTracker', 'DenseTimeSformer', 'MyDenseTimeSformerBackbone']
