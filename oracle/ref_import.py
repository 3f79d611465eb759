"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference Seeker from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to pin the oracle and generate ``tests/golden`` fixtures.
Recipe from SURVEY.md §8(c): the reference's ``from __init__ import *`` pulls in
plotting/logging packages that are absent here and unused by the forward math, so
they are registered as empty stub modules; imports resolve from cwd.
"""
from __future__ import annotations

import logging
import os
import sys
import types

REF = os.environ.get('TCOW_REF', '/root/reference')


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def import_reference_seeker():
    if not os.path.isdir(os.path.join(REF, 'model')):
        raise FileNotFoundError(f'reference tree not found at {REF}')

    class Registry:
        def __init__(self, name):
            self._d = {}

        def register(self, obj=None):
            if obj is None:
                return lambda o: self.register(o)
            self._d[obj.__name__] = obj
            return obj

        def get(self, name):
            return self._d[name]

    for n in ['fvcore', 'fvcore.common', 'fvcore.nn', 'simplejson', 'timm', 'imageio',
              'matplotlib', 'matplotlib.colors', 'matplotlib.pyplot', 'seaborn']:
        if n not in sys.modules:
            _stub(n)
    _stub('fvcore.common.registry', Registry=Registry)
    _stub('fvcore.nn.weight_init', c2_msra_fill=lambda *a, **k: None, c2_xavier_fill=lambda *a, **k: None)
    _stub('fvcore.common.file_io', PathManager=object)
    _stub('lovely_numpy', lo=lambda *a, **k: None)
    _stub('lovely_tensors', monkey_patch=lambda *a, **k: None)
    cwd = os.getcwd()
    os.chdir(REF)
    saved = list(sys.path)
    sys.path[:0] = [REF, REF + '/model', REF + '/third_party/TimeSformer']
    try:
        # The reference uses generic top-level names; make sure ours do not shadow them.
        for n in ['seeker', 'mask_tracker', 'vision_tf', 'resnet', '__init__']:
            sys.modules.pop(n, None)
        import seeker  # noqa: the reference's model/seeker.py
        return seeker.Seeker
    finally:
        os.chdir(cwd)
        sys.path[:] = saved


def build_reference(state_dict=None, **seeker_kwargs):
    Seeker = import_reference_seeker()
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        net = Seeker(logging.getLogger('tcow_ref'), **seeker_kwargs)
    finally:
        os.chdir(cwd)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net.eval()
