"""Checkpoint compatibility with the reference.

The reference writes ``checkpoint.pth`` in train.py:269-304 — a dict with ``epoch``, ``train_args`` (an
argparse.Namespace), ``dset_args``, ``seeker_args``, ``net_seeker`` (the 251-tensor state dict), ``optim_*`` and
``lr_sched_*`` — and reads it back in eval/inference.py:19-57.  ``load_networks`` below has that function's signature
and return value and builds the B200 drop-in instead of the stock module:

* torch >= 2.6 defaults to ``weights_only=True``, which rejects the pickled Namespace: load with ``weights_only=False``
  (checkpoints are trusted local files, as in the reference);
* ``seeker_args['tracker_pretrained']`` is normally truthy ('1', args.py:150).  In the reference that makes the
  constructor download ImageNet weights (vit.py:462-464) which ``load_state_dict`` then overwrites; its only lasting
  effect is the forward-time RGB normalisation keyed off ``backbone.pretrained`` (vision_tf.py:81-89).  Here the module is
  constructed without the download and the flag is set afterwards — same forward, no network access.
"""
from __future__ import annotations

import os

import torch

from . import seeker as seeker_mod
from .mask_tracker import _parse_tracker_pretrained as _parse_pretrained


def build_seeker(logger, seeker_args, state_dict=None, device=None):
    """Seeker(**seeker_args) without the pretrained download; restores the RGB-normalisation flag afterwards."""
    args = dict(seeker_args)
    pretrained, _ = _parse_pretrained(args.get('tracker_pretrained', False))
    args['tracker_pretrained'] = False
    net = seeker_mod.Seeker(logger, **args)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    net.seeker.tracker_pretrained = pretrained
    net.seeker.tracker_backbone.pretrained = pretrained
    net.seeker.tracker_backbone.timesformer.pretrained = pretrained
    if device is not None:
        net = net.to(device)
    return net


def resolve_checkpoint_file(path, epoch=-1):
    """A checkpoint directory holds `checkpoint.pth` (latest) and `model_<epoch>.pth` snapshots (train.py:269-304)."""
    if not os.path.exists(path):
        raise AssertionError(f'checkpoint path does not exist: {path}')       # the reference asserts (inference.py:28)
    if not os.path.isdir(path):
        return path
    return os.path.join(path, 'checkpoint.pth' if epoch < 0 else f'model_{epoch}.pth')


def load_networks(checkpoint_path, device, logger, epoch=-1):
    """Same call and same 5-tuple as eval/inference.py:19-57 — `(networks, train_args, dset_args, model_args, epoch)` —
    with the B200 Seeker in `networks['seeker']`.  `epoch >= 0` picks that snapshot of a checkpoint directory."""
    say = print if logger is None else logger.info
    file = resolve_checkpoint_file(checkpoint_path, epoch)
    say(f'Loading weights from: {file}')
    ckpt = torch.load(file, map_location='cpu', weights_only=False)        # pickled argparse.Namespace inside
    model_args = {'seeker': ckpt['seeker_args']}
    networks = {'seeker': build_seeker(logger, model_args['seeker'], ckpt['net_seeker'], device)}
    say(f"=> Loaded epoch (1-based): {ckpt['epoch'] + 1}")
    return networks, ckpt['train_args'], ckpt['dset_args'], model_args, ckpt['epoch']


def save_model_checkpoint(checkpoint_dir, epoch, train_args, dset_args, seeker_args, networks, optimizers=None,
                          lr_schedulers=None):
    """Writes the layout the reference reads back (train.py:269-296): `checkpoint.pth` with the argument records, one
    `net_<name>` / `optim_<name>` / `lr_sched_<name>` state dict per entry, plus `checkpoint_epoch.txt`."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    record = dict(epoch=epoch, train_args=train_args, dset_args=dset_args, seeker_args=seeker_args)
    for prefix, group in (('net_', networks), ('optim_', optimizers or {}), ('lr_sched_', lr_schedulers or {})):
        record.update({prefix + name: obj.state_dict() for name, obj in group.items()})
    path = os.path.join(checkpoint_dir, 'checkpoint.pth')
    torch.save(record, path)
    with open(os.path.join(checkpoint_dir, 'checkpoint_epoch.txt'), 'w') as f:
        f.write(f'{int(epoch)}\n')
    return path
