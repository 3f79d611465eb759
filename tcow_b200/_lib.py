"""ctypes binding of the C ABI in include/tcow_b200.h (libtcow_b200.so, built by `make` /
``__graft_entry__.build()``).  There is no fallback: a missing library or a non-sm_100 device raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCOW_B200_LIB selects an alternative build of the same ABI (A/B experiments); default is the in-tree library.
LIB_PATH = os.environ.get('TCOW_B200_LIB') or os.path.join(_HERE, 'libtcow_b200.so')

TCOW_ERR_ARG, TCOW_ERR_CUDA, TCOW_ERR_ARCH = -1, -2, -3
EPI_BF16, EPI_BF16_GELU, EPI_F32_STORE, EPI_F32_ADD = 0, 1, 2, 3
EPI_BF16_GELU_AUX, EPI_BF16_DGELU = 5, 6

# name -> argtypes (restype is int unless noted); mirrors include/tcow_b200.h one to one.
SIGNATURES = {
    'tcow_abi_version': [],
    'tcow_last_error': [],
    'tcow_check_device': [],
    'tcow_gemm_bf16': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                       c_int, c_void_p],
    'tcow_layernorm_bf16': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    'tcow_attn_temporal': [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p],
    'tcow_attn_spatial': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int,
                          c_int64, c_void_p],
    'tcow_cls_merge': [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_int, c_void_p],
    'tcow_patch_gather': [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_void_p],
    'tcow_patch_gather_typed': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                c_int, c_int, c_void_p],
    'tcow_clip_from_video_u8': [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p],
    'tcow_mask_clip_from_video_u8': [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_void_p, c_void_p],
    'tcow_patch_embed_fused': [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p],
    'tcow_embed_init': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    'tcow_mask_upsample': [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_void_p],
    'tcow_flag_mean': [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    'tcow_mask_iou_areas': [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    'tcow_mask_loss_workspace_floats': [],
    'tcow_mask_loss_sums': [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    'tcow_mask_loss_grad': [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    'tcow_mask_loss_state_bytes': [],
    'tcow_mask_loss_forward': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_int, c_int,
                               c_double, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p],
    'tcow_mask_loss_backward': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    'tcow_loss_class_counts': [c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p],
    'tcow_loss_pixel_weights': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                c_int, c_void_p, c_void_p, c_void_p],
    # ---- training step
    'tcow_gemm_bf16_aux': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                           c_int, c_int, c_int, c_int, c_void_p],
    'tcow_gemm_bf16_add_scaled': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int, c_int, c_int, c_void_p],
    'tcow_scale_rows_bf16': [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    'tcow_gemm_bf16_wgrad': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p],
    'tcow_gemm_bf16_wgrad_bias': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int,
                                  c_void_p],
    'tcow_gemm_bf16_wgrad_sched': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_void_p],
    'tcow_train_workspace_floats': [c_int],
    'tcow_layernorm_bf16_train': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                  c_void_p],
    'tcow_layernorm_bwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_int, c_int, c_int, c_void_p],
    'tcow_layernorm_bwd_scaled': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    'tcow_colsum_bf16': [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    'tcow_embed_bwd': [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    'tcow_mask_head_bwd': [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_int, c_int, c_int, c_void_p],
    'tcow_attn_spatial_train': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_int64, c_void_p],
    'tcow_attn_temporal_bwd': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int,
                               c_int, c_int, c_int, c_void_p],
    'tcow_attn_spatial_bwd': [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                              c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p],
    'tcow_cls_merge_bwd': [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_void_p],
}

_lib = None


class TcowError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TcowError(f'{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); '
                            'tcow_b200 has no fallback path')
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = (c_char_p if name == 'tcow_last_error' else
                          c_int64 if name in ('tcow_train_workspace_floats', 'tcow_mask_loss_workspace_floats',
                                              'tcow_mask_loss_state_bytes') else c_int)
        _lib = lib
    return _lib


def call(name, *args):
    """Invoke an entry point; map error codes onto the exception types the reference raises."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.tcow_last_error().decode(errors='replace')
        if rc == TCOW_ERR_ARG:
            raise ValueError(f'{name}: {msg}')
        raise TcowError(f'{name} failed ({rc}): {msg}')
    return rc
