"""The oracle restatement against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import cached_state_dict, golden_names, load_golden
from oracle import seeker_oracle
from tcow_b200 import synth

SMALL = [n for n in golden_names() if n.startswith(('small_', 'mid_', 'long_'))]


@pytest.mark.parametrize('name', SMALL)
def test_oracle_matches_reference_golden(name):
    meta, gmask, gflags = load_golden(name)
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    fc = meta.get('flag_channels', 3)
    sd = cached_state_dict(meta['weight_seed'], T, Hf, Wf, fc)
    rgb, q = synth.make_batch(meta['samples'], num_frames=T, frame_height=Hf, frame_width=Wf,
                              query_frame=meta.get('query_frame', 0))
    with torch.no_grad():
        mask, flags = seeker_oracle.seeker_forward(
            sd, rgb, q, causal_attention=meta['causal'], norm_embeddings=meta.get('norm_embeddings', False),
            pretrained_norm=meta.get('pretrained_norm', False),
            track_map_resize=meta.get('track_map_resize', 'bilinear'), flag_channels=fc)
    ly, lx = meta.get('lattice', (1, 1))
    assert (mask[:, :, :, ::ly, ::lx] - gmask).abs().max().item() < 2e-5
    if gflags is None:
        assert flags is None
    else:
        assert (flags - gflags).abs().max().item() < 2e-5


def test_oracle_full_size_one_sample():
    """Full north-star shape (T=30, 240x320), one sample, against the reference's golden lattice."""
    meta, gmask, gflags = load_golden('full_causal1')
    T, Hf, Wf = meta['T'], meta['Hf'], meta['Wf']
    sd = cached_state_dict(meta['weight_seed'], T, Hf, Wf)
    rgb, q = synth.make_batch(meta['samples'][:1], num_frames=T, frame_height=Hf, frame_width=Wf)
    with torch.no_grad():
        mask, flags = seeker_oracle.seeker_forward(sd, rgb, q, causal_attention=1)
    ly, lx = meta['lattice']
    assert (mask[:, :, :, ::ly, ::lx] - gmask[:1]).abs().max().item() < 2e-5
    assert (flags - gflags[:1]).abs().max().item() < 2e-5


def test_oracle_causality_property():
    """causal_attention=1: changing frames >= t0 leaves output frames < t0 bit-identical (vit.py:115-121)."""
    T, Hf, Wf, t0 = 6, 32, 32, 4
    sd = cached_state_dict(901, T, Hf, Wf)
    rgb, q = synth.make_batch([11], num_frames=T, frame_height=Hf, frame_width=Wf)
    rgb2 = rgb.clone()
    rgb2[:, :, t0:] = torch.rand_like(rgb2[:, :, t0:])
    with torch.no_grad():
        m1, f1 = seeker_oracle.seeker_forward(sd, rgb, q, causal_attention=1)
        m2, f2 = seeker_oracle.seeker_forward(sd, rgb2, q, causal_attention=1)
    assert torch.equal(m1[:, :, :t0], m2[:, :, :t0]) and torch.equal(f1[:, :t0], f2[:, :t0])
    assert (m1[:, :, t0:] - m2[:, :, t0:]).abs().max() > 1e-3
