import json
import logging
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 (B200) device')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    mask = torch.from_numpy(z['mask'])
    flags = torch.from_numpy(z['flags']) if 'flags' in z.files else None
    return meta, mask, flags


def golden_names():
    """Forward-pass fixtures (oracle/make_golden.py); gradient / loss / input-path / inflate fixtures have their own tests."""
    other = ('grad_', 'pretrained_', 'loss_', 'input_', 'metrics_')
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and not f.startswith(other))


@pytest.fixture(scope='session')
def logger():
    return logging.getLogger('tcow_b200_test')


_SD_CACHE = {}


def cached_state_dict(seed, T, Hf, Wf, flag_channels=3):
    from tcow_b200 import synth
    key = (seed, T, Hf, Wf, flag_channels)
    if key not in _SD_CACHE:
        if len(_SD_CACHE) > 2:
            _SD_CACHE.clear()
        _SD_CACHE[key] = synth.make_state_dict(seed, num_frames=T, frame_height=Hf, frame_width=Wf,
                                               flag_channels=flag_channels)
    return _SD_CACHE[key]
