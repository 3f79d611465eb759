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

import numpy as np
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


def load_networks(checkpoint_path, device, logger, epoch=-1):
    '''
    Drop-in for eval/inference.py:19-57.
    :param checkpoint_path (str): Path to model checkpoint folder or file.
    :param epoch (int): If >= 0, desired checkpoint epoch to load.
    :return (networks, train_args, dset_args, model_args, epoch).
    '''
    print_fn = logger.info if logger is not None else print
    assert os.path.exists(checkpoint_path)
    if os.path.isdir(checkpoint_path):
        model_fn = f'model_{epoch}.pth' if epoch >= 0 else 'checkpoint.pth'
        checkpoint_path = os.path.join(checkpoint_path, model_fn)
    print_fn('Loading weights from: ' + checkpoint_path)
    checkpoint = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
    train_args = checkpoint['train_args']
    train_dset_args = checkpoint['dset_args']
    seeker_args = checkpoint['seeker_args']
    model_args = {'seeker': seeker_args}
    seeker_net = build_seeker(logger, seeker_args, checkpoint['net_seeker'], device)
    networks = {'seeker': seeker_net}
    epoch = checkpoint['epoch']
    print_fn('=> Loaded epoch (1-based): ' + str(epoch + 1))
    return (networks, train_args, train_dset_args, model_args, epoch)


def save_model_checkpoint(checkpoint_dir, epoch, train_args, dset_args, seeker_args, networks, optimizers=None,
                          lr_schedulers=None):
    """Writes what train.py:269-296 writes (checkpoint.pth + checkpoint_epoch.txt), so the reference can read it back."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    checkpoint = {'epoch': epoch, 'train_args': train_args, 'dset_args': dset_args, 'seeker_args': seeker_args}
    for (k, v) in networks.items():
        checkpoint['net_' + k] = v.state_dict()
    for (k, v) in (optimizers or {}).items():
        checkpoint['optim_' + k] = v.state_dict()
    for (k, v) in (lr_schedulers or {}).items():
        checkpoint['lr_sched_' + k] = v.state_dict()
    path = os.path.join(checkpoint_dir, 'checkpoint.pth')
    torch.save(checkpoint, path)
    np.savetxt(os.path.join(checkpoint_dir, 'checkpoint_epoch.txt'), np.array([epoch], dtype=np.int32), fmt='%d')
    return path
