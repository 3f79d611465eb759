"""Host-side logic that needs no GPU: state-dict layout, constructor contract, C-ABI symbols, weight folding."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

import tcow_b200
from tcow_b200 import _lib, synth
from tcow_b200.engine import SeekerEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(num_total_frames=4, num_visible_frames=4, frame_height=32, frame_width=48, tracker_pretrained=False,
          attention_type='divided_space_time', patch_size=16, causal_attention=1, norm_embeddings=False,
          drop_path_rate=0.1, network_depth=12, track_map_stride=4, track_map_resize='bilinear',
          query_channels=1, output_channels=3, flag_channels=3)


@pytest.fixture(scope='module')
def net(logger):
    return tcow_b200.Seeker(logger, **KW)


def test_state_dict_layout_matches_reference(net):
    """251 tensors with the names/shapes of checkpoint['net_seeker'] (train.py:281)."""
    want = synth.state_dict_shapes(num_frames=4, frame_height=32, frame_width=48)
    got = net.state_dict()
    assert len(got) == 251
    assert list(got.keys()) == list(want.keys())
    for k, shp in want.items():
        assert tuple(got[k].shape) == tuple(shp), k
        assert got[k].dtype == torch.float32
    assert sum(p.numel() for p in net.parameters()) == sum(torch.Size(s).numel() for s in want.values())


def test_param_count_north_star(logger):
    shapes = synth.state_dict_shapes(num_frames=30, frame_height=240, frame_width=320)
    assert sum(torch.Size(s).numel() for s in shapes.values()) == 122145027


def test_reference_init_semantics(net):
    """vit.py:264-306: time_embed and every temporal_fc are zero, biases zero, LN (1,0)."""
    bb = net.seeker.tracker_backbone.timesformer.model
    assert bb.time_embed.abs().max() == 0
    for blk in bb.blocks:
        assert blk.temporal_fc.weight.abs().max() == 0 and blk.temporal_fc.bias.abs().max() == 0
        assert blk.attn.qkv.bias.abs().max() == 0
        assert torch.all(blk.norm1.weight == 1) and torch.all(blk.norm1.bias == 0)
        assert 0.018 < blk.mlp.fc1.weight.std() < 0.022 and blk.mlp.fc1.weight.abs().max() <= 2.0
    assert net.seeker.tracker_post_linear.weight.abs().max() <= 768 ** -0.5 + 1e-6  # torch default Linear init


def test_state_dict_roundtrip(net, logger):
    sd = synth.make_state_dict(7, num_frames=4, frame_height=32, frame_width=48)
    net2 = tcow_b200.Seeker(logger, **KW)
    net2.load_state_dict(sd, strict=True)
    for k, v in net2.state_dict().items():
        assert torch.equal(v, sd[k])


def test_constructor_contract(logger):
    with pytest.raises(ValueError):
        tcow_b200.Seeker(logger, **{**KW, 'tracker_pretrained': 3.5})       # mask_tracker.py:67
    with pytest.raises(ValueError):
        tcow_b200.Seeker(logger, **{**KW, 'network_depth': 13})             # vit.py:449
    with pytest.raises(AssertionError):
        tcow_b200.Seeker(logger, **{**KW, 'frame_height': 40})              # mask_tracker.py:89
    with pytest.raises(AssertionError):
        tcow_b200.Seeker(logger, **{**KW, 'attention_type': 'bogus'})       # vit.py:133
    m = tcow_b200.Seeker(logger, **{**KW, 'tracker_pretrained': 'no'})      # short string -> False
    assert m.seeker.tracker_pretrained is False
    m = tcow_b200.Seeker(logger, **{**KW, 'flag_channels': 0})
    assert not hasattr(m.seeker, 'flag_post_linear') and len(m.state_dict()) == 249


def test_cpu_forward_fails_loudly(net):
    rgb, q = synth.make_batch([0], num_frames=4, frame_height=32, frame_width=48)
    with torch.no_grad(), pytest.raises(RuntimeError, match='CUDA'):
        net(rgb, q)
    with pytest.raises(AssertionError):
        net(rgb, torch.cat([q, q], 1))                                      # mask_tracker.py:105


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/tcow_b200.h declares."""
    hdr = open(os.path.join(ROOT, 'include', 'tcow_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(tcow_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} not exported'
    assert _lib.load().tcow_abi_version() == 1


def test_library_rejects_bad_arguments_without_gpu():
    with pytest.raises(ValueError, match='multiples of 64'):
        _lib.call('tcow_gemm_bf16', 16, 64, 16, 64, None, 16, 100, 128, 100, 64, 0, None)
    with pytest.raises(ValueError):
        _lib.call('tcow_attn_temporal', 16, 2304, 16, 768, 10, 65, 12, 0, None)
    with pytest.raises(ValueError):
        _lib.call('tcow_layernorm_bf16', 16, 16, 16, 16, 8, 100, 1e-6, None)


def test_head_fold_equals_pool_then_linear(net):
    """avg_pool2d folded into tracker_post_linear == reference order (mask_tracker.py:113-122)."""
    torch.manual_seed(0)
    mod = net.seeker
    with torch.no_grad():
        mod.tracker_post_linear.weight.normal_(0, 0.05)
        mod.tracker_post_linear.bias.normal_(0, 0.05)
        mod.flag_post_linear.weight.normal_(0, 0.05)
    eng = SeekerEngine(mod)
    pk = eng._pack(mod, torch.device('cpu'))
    assert pk.pp == 4 and pk.stride == 4 and pk.flag_col0 == 48 and pk.n_pad == 64
    feat = torch.randn(5, 768)
    ref = F.linear(feat, mod.tracker_post_linear.weight, mod.tracker_post_linear.bias).reshape(5, 3, 16, 16)
    ref = F.avg_pool2d(ref, 4, 4).reshape(5, 48)
    Wh = pk.head_w.float()
    got = feat.to(torch.bfloat16).float() @ Wh.t() + pk.head_b
    assert (got[:, :48] - ref).abs().max() < 0.05
    # exact check in fp32 (undo the bf16 rounding of the packed copy)
    Wt = mod.tracker_post_linear.weight.reshape(3, 4, 4, 4, 4, 768).mean((2, 4)).reshape(48, 768)
    assert (feat @ Wt.t() + pk.head_b[:48] - ref).abs().max() < 1e-5
    assert torch.equal(pk.head_w[48:51].float(), mod.flag_post_linear.weight.to(torch.bfloat16).float())
    assert pk.head_w[51:].abs().max() == 0


def test_merged_temporal_projection_is_algebraically_exact(net):
    torch.manual_seed(1)
    mod = net.seeker
    blk = mod.tracker_backbone.timesformer.model.blocks[3]
    with torch.no_grad():
        blk.temporal_fc.weight.normal_(0, 0.02); blk.temporal_fc.bias.normal_(0, 0.02)
        blk.temporal_attn.proj.bias.normal_(0, 0.02)
    o = torch.randn(7, 768, dtype=torch.float64)
    ref = F.linear(F.linear(o, blk.temporal_attn.proj.weight.double(), blk.temporal_attn.proj.bias.double()),
                   blk.temporal_fc.weight.double(), blk.temporal_fc.bias.double())
    W = blk.temporal_fc.weight.double() @ blk.temporal_attn.proj.weight.double()
    b = blk.temporal_fc.weight.double() @ blk.temporal_attn.proj.bias.double() + blk.temporal_fc.bias.double()
    assert (F.linear(o, W, b) - ref).abs().max() < 1e-12
    pk = SeekerEngine(mod)._pack(mod, torch.device('cpu'))
    assert (pk.blocks[3].t_out[0].double() - W).abs().max() <= 2.0 ** -8 * W.abs().max()   # one bf16 rounding


def test_checkpoint_round_trip_in_reference_format(tmp_path, logger):
    """train.py:269-304 layout -> tcow_b200.checkpoint.load_networks (eval/inference.py:19-57 signature)."""
    import argparse

    from tcow_b200 import checkpoint
    T, Hf, Wf = 4, 32, 48
    seeker_args = dict(num_total_frames=T, num_visible_frames=T, frame_height=Hf, frame_width=Wf, tracker_pretrained='1',
                       attention_type='divided_space_time', patch_size=16, causal_attention=1, norm_embeddings=False,
                       drop_path_rate=0.1, network_depth=12, track_map_stride=4, track_map_resize='bilinear',
                       query_channels=1, output_channels=3, flag_channels=3)
    sd = synth.make_state_dict(901, num_frames=T, frame_height=Hf, frame_width=Wf)
    net = checkpoint.build_seeker(logger, seeker_args, sd)
    assert net.seeker.tracker_backbone.pretrained is True          # RGB normalisation flag without any download
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    train_args = argparse.Namespace(name='unit', num_frames=T, learn_rate=1e-4)   # a Namespace: needs weights_only=False
    path = checkpoint.save_model_checkpoint(str(tmp_path), 7, train_args, {'dset': 1}, seeker_args, {'seeker': net},
                                            {'seeker': opt})
    raw = torch.load(path, map_location='cpu', weights_only=False)
    assert set(raw) >= {'epoch', 'train_args', 'dset_args', 'seeker_args', 'net_seeker', 'optim_seeker'}
    assert len(raw['net_seeker']) == 251
    (networks, targs, dargs, margs, epoch) = checkpoint.load_networks(str(tmp_path), 'cpu', logger)
    assert epoch == 7 and targs.name == 'unit' and dargs == {'dset': 1} and margs['seeker'] == seeker_args
    got = networks['seeker'].state_dict()
    assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert networks['seeker'].seeker.tracker_backbone.pretrained is True
    assert open(tmp_path / 'checkpoint_epoch.txt').read().strip() == '7'
    off = dict(seeker_args, tracker_pretrained='0')
    assert checkpoint.build_seeker(logger, off).seeker.tracker_backbone.pretrained is False


def test_pretrained_inflate_matches_reference_fixture(tmp_path, logger):
    """tracker_pretrained=<local file>: helpers.py:100-202 (conv channels tiled 3 -> 4 and rescaled, pos_embed resampled
    196 -> 24 patches, temporal attention / temporal_norm1 copied from the spatial ones) — digest of every backbone tensor
    equals what the unmodified reference built from the same file (oracle/make_golden_pretrained.py)."""
    import json

    import numpy as np

    from oracle import make_golden_pretrained as mgp
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'pretrained_inflate.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    path = str(tmp_path / 'vit.pth')
    torch.save(mgp.fake_imagenet_vit(meta['seed']), path)
    net = tcow_b200.Seeker(logger, tracker_pretrained=path, **meta['kwargs'])
    assert net.seeker.tracker_pretrained is True and net.seeker.tracker_backbone.pretrained is True
    got = mgp.backbone_digests(net.state_dict())
    want = {k.replace('/', '.'): z[k] for k in z.files if k != 'meta'}
    assert set(got) == set(want) and len(got) == 247
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    # the projection may be stored as a linear layer (vit.py:381-390): same result
    torch.save(mgp.fake_imagenet_vit(meta['seed'], linear_patch_proj=True), path)
    net2 = tcow_b200.Seeker(logger, tracker_pretrained=path, **meta['kwargs'])
    assert torch.equal(net2.state_dict()[mgp.PREFIX + 'patch_embed.proj.weight'],
                       net.state_dict()[mgp.PREFIX + 'patch_embed.proj.weight'])


def test_pretrained_without_a_file_fails_loudly(logger, monkeypatch, tmp_path):
    monkeypatch.delenv('TCOW_PRETRAINED_VIT', raising=False)
    monkeypatch.setattr(torch.hub, 'get_dir', lambda: str(tmp_path))
    with pytest.raises(RuntimeError, match='offline'):
        tcow_b200.Seeker(logger, **{**KW, 'tracker_pretrained': '1'})       # args.py:150 default, no network here
    with pytest.raises(FileNotFoundError):
        tcow_b200.Seeker(logger, **{**KW, 'tracker_pretrained': str(tmp_path / 'missing_file.pth')})


def _as_data_parallel_replica(module):
    """What torch.nn.parallel.replicate leaves on a device: no _parameters, broadcast copies in _former_parameters."""
    memo = {}
    for m in module.modules():
        r = m._replicate_for_data_parallel()
        r._former_parameters = {}
        memo[m] = r
    for m, r in memo.items():
        for k, child in m._modules.items():
            r._modules[k] = memo[child] if child is not None else None
        for k, p in m._parameters.items():
            if p is not None:
                t = p * 1.0                      # non-leaf copy that still requires grad, like Broadcast.apply's output
                setattr(r, k, t)
                r._former_parameters[k] = t
    return memo[module]


def test_data_parallel_replica_in_grad_mode_names_the_ddp_route(net):
    rep = _as_data_parallel_replica(net.seeker)
    assert len(list(rep.parameters())) == 0
    rgb, q = synth.make_batch([0], num_frames=4, frame_height=32, frame_width=48)
    rep.train()
    with pytest.raises(RuntimeError, match='ddp.attach'):
        rep(rgb, q)
    rep.eval()                                   # inference replicas are fine (they fail later only for lack of a GPU)
    with pytest.raises(RuntimeError, match='CUDA'):
        rep(rgb, q)
    # the weight-cache stamp sees the replica's tensors (a stale stamp would mean stale packed weights)
    assert len(SeekerEngine._stamp(rep)) == len(SeekerEngine._stamp(net.seeker)) > 0


def test_ddp_broadcast_bumps_the_version_stamp(net):
    """ddp.broadcast_parameters writes through p.detach() (shares the version counter), not p.data."""
    p = next(net.parameters())
    v0 = p._version
    with torch.no_grad():
        p.detach().copy_(p.detach() + 0)
    assert p._version == v0 + 1
    src = open(os.path.join(ROOT, 'tcow_b200', 'ddp.py')).read()
    assert 'dist.broadcast(p.data' not in src
