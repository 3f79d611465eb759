"""Training step of the Seeker (forward with saved activations + hand-written backward) on sm_100a kernels.

The reference trains through torch.autograd over model/mask_tracker.py:92-142 -> model/vision_tf.py:68-169 ->
third_party/TimeSformer/timesformer/models/vit.py:155-217 (train.py:93-101: loss.backward(); optimizer.step()).
Here the same function is differentiated by hand: `SeekerTrainEngine.forward` runs the plan of engine.py with
training variants of the kernels (LayerNorm saves xhat/rstd, fc1 saves its pre-activation, spatial attention saves its
log-sum-exp) and `backward` walks it in reverse through the adjoint kernels of include/tcow_b200.h:

  dX  of every nn.Linear : the forward tcgen05 GEMM on the transposed bf16 weight (dgelu fused for fc1)
  dW  of every nn.Linear : tcow_gemm_bf16_wgrad (tcgen05, MN-major operands — activations are never transposed)
  db                     : tcow_colsum_bf16
  LayerNorm              : tcow_layernorm_bwd, fused with the fp32 residual-gradient accumulate + its bf16 copy
  attention              : tcow_attn_temporal_bwd / tcow_attn_spatial_bwd (recompute P from the saved lse)
  embeddings, mask head  : tcow_embed_bwd, tcow_mask_head_bwd

Gradient layout: all gradients of one backward live in ONE flat fp32 buffer in "packed" parameter space (merged
temporal projection, pooled head), block by block in reverse execution order, so that a data-parallel run can
all-reduce finished blocks on a side stream while earlier blocks are still being differentiated (ddp.py).  The packed
gradients are mapped back onto the reference's 251 parameters by small fp32 weight-space products.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from .ops import (EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU_AUX, EPI_F32_ADD, EPI_F32_STORE)

HEADS = 12


class _W:
    pass


def _mm(a, b):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return a @ b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


class _GradLayout:
    """Offsets of every packed gradient inside the flat buffer; blocks last-to-first (= completion order)."""

    def __init__(self, depth, D, Kp, n_pos, T, n_pad, merged):
        self.slots = {}
        self.block_ranges = []
        off = 0

        def add(name, *shape):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            self.slots[name] = (off, tuple(shape))
            off += (n + 63) // 64 * 64          # 256-byte aligned slots (TMA reduce-add needs 16 B)
        add('head_w', n_pad, D); add('head_b', n_pad); add('norm_g', D); add('norm_b', D)
        self.head_range = (0, off)
        for i in reversed(range(depth)):
            o0 = off
            p = f'b{i}.'
            add(p + 'fc2_w', D, 4 * D); add(p + 'fc2_b', D); add(p + 'fc1_w', 4 * D, D); add(p + 'fc1_b', 4 * D)
            add(p + 'n2_g', D); add(p + 'n2_b', D)
            add(p + 's_proj_w', D, D); add(p + 's_proj_b', D); add(p + 's_qkv_w', 3 * D, D); add(p + 's_qkv_b', 3 * D)
            add(p + 'n1_g', D); add(p + 'n1_b', D)
            if merged:
                add(p + 't_out_w', D, D); add(p + 't_out_b', D); add(p + 't_out_b2', D)
            else:
                add(p + 't_fc_w', D, D); add(p + 't_fc_b', D); add(p + 't_proj_w', D, D); add(p + 't_proj_b', D)
            add(p + 't_qkv_w', 3 * D, D); add(p + 't_qkv_b', 3 * D); add(p + 'tn1_g', D); add(p + 'tn1_b', D)
            self.block_ranges.append((o0, off))
        o0 = off
        add('patch_w', D, Kp); add('pos', n_pos, D); add('time', T, D)
        self.embed_range = (o0, off)
        self.total = off

    def view(self, flat, name):
        off, shape = self.slots[name]
        n = 1
        for s in shape:
            n *= s
        return flat[off:off + n].view(shape)


class _Saved:
    """Activations one forward call keeps for its backward (owned by the autograd node, freed with it)."""
    pass


class SeekerTrainEngine:
    def __init__(self, tracker, merge_temporal_proj=True):
        self.merge_temporal_proj = merge_temporal_proj
        self._scratch = {}
        self.launches = 0
        self.profile = None
        self.grad_sync = None        # optional ddp.GradSync: all-reduces finished ranges of the flat gradient buffer
        self._tickets = {}           # device index -> int32[2] ticket counters of the weight-gradient GEMM scheduler
        # tests inject fixed stochastic-depth keep masks here: list (one per block) of None or
        # dict(t=[B*N], s=[B*T], m=[B]) 0/1 tensors; None = draw them (vit_utils.py:139-164)
        self.drop_path_override = None
        self.last_flat = None

    # ------------------------------------------------------------------ plumbing
    def _launch(self, kind, fn, *args, flops=0.0, nbytes=0.0):
        if self.profile is None:
            fn(*args)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args)
            e1.record()
            self.profile.append((kind, flops, nbytes, e0, e1))
        self.launches += 1

    def _gemm(self, kind, a, w, bias, out, epi):
        M, K = a.shape
        N = w.shape[0]
        self._launch(kind, ops.gemm, a, w, bias, out, epi, flops=2.0 * M * N * K,
                     nbytes=2.0 * (M * K + N * K) + out.element_size() * M * N * (2 if epi == EPI_F32_ADD else 1))

    def _wgrad(self, kind, dy, x, dw, db=None):
        """dw += dy^T x; with db also db += dy.sum(0) (the layer's bias gradient, from the same pass over dy)."""
        R, N1 = dy.shape
        self._launch(kind, ops.gemm_wgrad, dy, x, dw, db, self._wgrad_tickets(dy.device), flops=2.0 * R * N1 * x.shape[1],
                     nbytes=2.0 * R * (N1 + x.shape[1]) + 4.0 * dw.numel())

    def _wgrad_tickets(self, device):
        """Ticket counters of the weight-gradient GEMM's dynamic unit scheduler (ops.gemm_wgrad `sched`).  It matters when a
        gradient exchange runs beside the backward (ddp.attach): NCCL's all-reduce kernels then share the SMs, and static
        striping leaves a tail behind the SMs they occupy.  On one GPU it measured neutral (60.4 vs 59.9 ms per step), so it
        is simply always on; TCOW_WGRAD_SCHED=0 goes back to static striping."""
        if os.environ.get('TCOW_WGRAD_SCHED', '1') == '0':
            return None
        t = self._tickets.get(device.index)
        if t is None:
            t = self._tickets[device.index] = torch.zeros(2, device=device, dtype=torch.int32)
        return t

    # ------------------------------------------------------------------ weights (repacked every step: they change)
    def _pack(self, mod, device):
        bb = mod.tracker_backbone.timesformer.model
        D = bb.embed_dim
        P = mod.patch_size
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32)
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        bft = lambda t: t.t().to(torch.bfloat16).contiguous()
        pk = _W()
        pk.patch_w = bf(f32(bb.patch_embed.proj.weight).reshape(D, -1))
        pk.patch_b = f32(bb.patch_embed.proj.bias).contiguous()
        pk.pos = f32(bb.pos_embed).reshape(-1, D).contiguous()
        pk.time = f32(bb.time_embed).reshape(-1, D).contiguous()
        pk.cls = f32(bb.cls_token).reshape(D).contiguous()
        # Per weight type, the 12 blocks are packed together: one stack, one bf16 cast and one transposing cast per type
        # (about 40 small launches per step instead of 300).
        blocks = list(bb.blocks)
        nb = len(blocks)

        def packed_linear(get):
            W = torch.stack([f32(get(blk).weight) for blk in blocks])                     # [nb, out, in] fp32
            Wb = W.to(torch.bfloat16)
            Wt = torch.empty((nb, W.shape[2], W.shape[1]), device=device, dtype=torch.bfloat16)
            Wt.copy_(W.transpose(1, 2))                                                    # cast + transpose in one pass
            bias = [f32(get(blk).bias).contiguous() for blk in blocks]
            return [(Wb[i], bias[i], Wt[i]) for i in range(nb)]

        t_qkv = packed_linear(lambda blk: blk.temporal_attn.qkv)
        s_qkv = packed_linear(lambda blk: blk.attn.qkv)
        s_proj = packed_linear(lambda blk: blk.attn.proj)
        fc1 = packed_linear(lambda blk: blk.mlp.fc1)
        fc2 = packed_linear(lambda blk: blk.mlp.fc2)
        Wp_all = torch.stack([f32(blk.temporal_attn.proj.weight) for blk in blocks])
        bp_all = torch.stack([f32(blk.temporal_attn.proj.bias) for blk in blocks])
        Wf_all = torch.stack([f32(blk.temporal_fc.weight) for blk in blocks])
        bf_all = torch.stack([f32(blk.temporal_fc.bias) for blk in blocks])
        if self.merge_temporal_proj:   # fc(proj(o)) = o (Wf Wp)^T + (Wf bp + bf)   (vit.py:111 -> :174, no nonlinearity)
            Wm_all = _mm(Wf_all, Wp_all)                                                   # batched fp32 product
            b1_all = _mm(Wf_all, bp_all[:, :, None])[:, :, 0].contiguous()                 # the bias part inside DropPath
            bm_all = b1_all + bf_all
            Wm_b = Wm_all.to(torch.bfloat16)
            Wm_t = torch.empty_like(Wm_b)
            Wm_t.copy_(Wm_all.transpose(1, 2))
        else:
            t_proj = packed_linear(lambda blk: blk.temporal_attn.proj)
            t_fc = packed_linear(lambda blk: blk.temporal_fc)
        pk.blocks = []
        pk.raw_t_all = (Wp_all, bp_all, Wf_all, bf_all)
        for i, blk in enumerate(blocks):
            w = _W()
            ln = lambda m: (f32(m.weight).contiguous(), f32(m.bias).contiguous())
            w.tn1, w.n1, w.n2 = ln(blk.temporal_norm1), ln(blk.norm1), ln(blk.norm2)
            w.t_qkv, w.s_qkv, w.s_proj, w.fc1, w.fc2 = t_qkv[i], s_qkv[i], s_proj[i], fc1[i], fc2[i]
            w.raw_t = (Wp_all[i], bp_all[i], Wf_all[i], bf_all[i])
            if self.merge_temporal_proj:
                w.t_out = (Wm_b[i], bm_all[i], Wm_t[i], b1_all[i], bf_all[i])
            else:
                w.t_proj, w.t_fc = t_proj[i], t_fc[i]
            pk.blocks.append(w)
        pk.norm = (f32(bb.norm.weight).contiguous(), f32(bb.norm.bias).contiguous())
        C, s = mod.output_channels, max(int(mod.track_map_stride), 1)
        if P % s != 0:
            raise NotImplementedError(f'track_map_stride={s} must divide patch_size={P}')
        pp = P // s
        Wt, bt = f32(mod.tracker_post_linear.weight), f32(mod.tracker_post_linear.bias)
        rows = [Wt.reshape(C, pp, s, pp, s, D).mean((2, 4)).reshape(C * pp * pp, D)]
        biases = [bt.reshape(C, pp, s, pp, s).mean((2, 4)).reshape(-1)]
        F = mod.flag_channels
        if F > 0:
            rows.append(f32(mod.flag_post_linear.weight))
            biases.append(f32(mod.flag_post_linear.bias))
        n_used = C * pp * pp + max(F, 0)
        n_pad = (n_used + 63) // 64 * 64
        Wh = torch.zeros(n_pad, D, device=device)
        bh = torch.zeros(n_pad, device=device)
        Wh[:n_used] = torch.cat(rows, 0)
        bh[:n_used] = torch.cat(biases, 0)
        pk.head = (bf(Wh), bh, bft(Wh))
        pk.pp, pk.stride, pk.flag_col0, pk.n_pad = pp, s, C * pp * pp, n_pad
        return pk

    def _scratch_for(self, device, R, D, n_pad):
        key = (device.index, R, D, n_pad)
        sc = self._scratch.get(key)
        if sc is None:
            e = lambda shape, dt: torch.empty(shape, device=device, dtype=dt)
            sc = dict(X=e((R, D), torch.float32), G=e((R, D), torch.float32), Gb=e((R, D), torch.bfloat16),
                      dA=e((R, D), torch.bfloat16), dO=e((R, D), torch.bfloat16), Gs=e((R, D), torch.bfloat16), dQKV=e((R, 3 * D), torch.bfloat16),
                      dZ=e((R, 4 * D), torch.bfloat16), LOW=e((R, n_pad), torch.float32),
                      dLOW=e((R, n_pad), torch.bfloat16), ws=e((ops.train_workspace_floats(4 * D),), torch.float32))
            if len(self._scratch) >= 2:
                self._scratch.clear()
            self._scratch[key] = sc
        return sc

    # ------------------------------------------------------------------ stochastic depth
    def _drop_path_scales(self, mod, depth, B, N, T, M, R, Rs, use_cls, causal, device):
        """Per-block row scales of the three residual branches (DropPath, vit_utils.py:139-164): the reference draws one
        Bernoulli(keep) per leading-dim entry of the branch — (b h w) sequences for the temporal branch, (b t) frames for
        the spatial one, b samples for the MLP (vit.py:170-172, 181-186, 216) — and scales survivors by 1/keep; rates are
        linspace(0, drop_path_rate, depth) (vit.py:272-273), block 0 has none."""
        rate = float(mod.drop_path_rate)
        if not (mod.training and rate > 0.0) and self.drop_path_override is None:
            return None
        rates = torch.linspace(0, rate, depth).tolist()
        out = []
        for i in range(depth):
            ov = None if self.drop_path_override is None else self.drop_path_override[i]
            if ov is None and (self.drop_path_override is not None or rates[i] == 0.0):
                out.append(None)
                continue
            keep = 1.0 - rates[i]
            if ov is not None:
                keep = float(ov.get('keep', keep))
                mt, ms, mm = (ov[k].to(device=device, dtype=torch.float32) for k in ('t', 's', 'm'))
            else:
                mt, ms, mm = (torch.floor(keep + torch.rand(n, device=device)) for n in (B * N, B * T, B))
            st, ss, sm = mt / keep, ms / keep, mm / keep
            rs_t = st.repeat_interleave(T).contiguous()
            rs_s = ss.view(B, 1, T).expand(B, N, T).reshape(M)
            bs_s = rs_s
            if use_cls:
                if causal == 1:      # only frame 0's cls output is used (vit.py:198)
                    c = ss.view(B, T)[:, 0]
                    rs_s, bs_s = torch.cat([rs_s, c]), torch.cat([rs_s, c])
                else:                # the mean over frames is taken after DropPath: the matmul part is pre-weighted
                    rs_s, bs_s = torch.cat([rs_s, torch.ones(B, device=device)]), torch.cat([rs_s, ss.view(B, T).mean(1)])
            rs_m = torch.cat([sm.repeat_interleave(N * T), sm])
            pad = lambda v: torch.cat([v, v.new_zeros(R - v.numel())]).contiguous() if v.numel() < R else v.contiguous()
            out.append(dict(rs_t=rs_t, rs_s=rs_s.contiguous(), bs_s=bs_s.contiguous(), rs_m=rs_m.contiguous(), ss=ss,
                            rs_t_pad=pad(rs_t), rs_s_pad=pad(rs_s)))
        return out

    # ------------------------------------------------------------------ forward (saves what backward needs)
    def forward(self, mod, input_frames, query_mask, queries_per_video=1):
        if not input_frames.is_cuda:
            raise RuntimeError('tcow_b200 runs on a CUDA sm_100 device only; move the module and inputs to the GPU '
                               '(there is no CPU fallback)')
        device = input_frames.device
        bbm = mod.tracker_backbone
        V, Cin, T, Hf, Wf = input_frames.shape
        B = V * queries_per_video
        if Cin != 3:
            raise RuntimeError(f'expected 3 RGB channels (+1 query channel = in_chans 4), got {Cin}')
        if tuple(query_mask.shape) != (B, 1, T, Hf, Wf):
            raise RuntimeError(f'query_mask shape {tuple(query_mask.shape)} does not match frames {tuple(input_frames.shape)}')
        assert T == bbm.T                                   # vision_tf.py:96
        assert Hf == bbm.Hf and Wf == bbm.Wf                # vision_tf.py:97
        P, D = mod.patch_size, bbm.output_feature_dim
        Ho, Wo = Hf // P, Wf // P
        N = Ho * Wo
        causal = int(mod.causal_attention)
        if causal in (0, 1):
            use_cls = True
        elif causal >= 2 or causal == -1:
            use_cls = False
        else:
            raise ValueError(f'unsupported causal_attention={causal}')
        causal_diag = -1 if causal <= 0 else (0 if causal <= 2 else causal - 2)   # vit.py:93-99
        if mod.track_map_resize not in ('bilinear', 'nearest'):
            raise ValueError(f'unsupported track_map_resize={mod.track_map_resize!r}')
        if N + (1 if use_cls else 0) > 304:
            raise NotImplementedError('training supports up to 304 tokens per frame (the spatial attention backward '
                                      'keeps one frame in shared memory)')

        with torch.cuda.device(device), torch.no_grad():
            _lib.call('tcow_check_device')
            pk = self._pack(mod, device)
            frames = input_frames.to(torch.float32).contiguous()
            query = query_mask.to(torch.float32).contiguous()
            M = B * N * T
            R = M + B
            Rs = R if use_cls else M
            Kp = 4 * P * P
            sc = self._scratch_for(device, R, D, pk.n_pad)
            X, LOW = sc['X'], sc['LOW']
            e = lambda shape, dt=torch.bfloat16: torch.empty(shape, device=device, dtype=dt)
            sv = _Saved()
            sv.dims = (B, N, T, D, P, Ho, Wo, M, R, Rs, Kp, use_cls, causal, causal_diag)
            sv.pk = pk
            sv.PM = e((M, Kp))
            sv.blocks = []
            L, G = self._launch, self._gemm
            self.launches = 0
            drop = self._drop_path_scales(mod, len(pk.blocks), B, N, T, M, R, Rs, use_cls, causal, device)
            if drop is not None and not self.merge_temporal_proj:
                raise NotImplementedError('stochastic depth needs the merged temporal projection (merge_temporal_proj=True)')
            L('patch_gather', ops.patch_gather, frames, query, sv.PM, P, bool(bbm.pretrained), queries_per_video, 0)
            L('embed_init', ops.embed_init, X, pk.patch_b, pk.pos, pk.time, pk.cls, B, N, T, D)
            G('gemm_patch', sv.PM, pk.patch_w, None, X[:M], EPI_F32_ADD)

            def ln_save(rows, params):
                a, xh, rs = e((rows, D)), e((rows, D)), e((rows,), torch.float32)
                L('ln_train', ops.layernorm_train, X[:rows], params[0], params[1], a, xh, rs, nbytes=8.0 * rows * D)
                return a, xh, rs

            for bi, w in enumerate(pk.blocks):
                s = _Saved()
                s.dp = dp = None if drop is None else drop[bi]
                # ---- temporal attention + temporal_fc + residual (vit.py:169-176); cls rows untouched
                s.A_t, s.xh_t, s.rs_t = ln_save(M, w.tn1)
                s.QKV_t, s.O_t = e((M, 3 * D)), e((M, D))
                G('gemm_qkv', s.A_t, w.t_qkv[0], w.t_qkv[1], s.QKV_t, EPI_BF16)
                L('attn_temporal', ops.attn_temporal, s.QKV_t, s.O_t, B * N, T, HEADS, causal_diag)
                if dp is not None:     # drop_path sits between proj and temporal_fc (vit.py:172-174)
                    L('gemm_proj', ops.gemm_add_scaled, s.O_t, w.t_out[0], w.t_out[3], w.t_out[4], dp['rs_t'], dp['rs_t'],
                      X[:M], flops=2.0 * M * D * D)
                elif self.merge_temporal_proj:
                    G('gemm_proj', s.O_t, w.t_out[0], w.t_out[1], X[:M], EPI_F32_ADD)
                else:
                    s.P_t = e((M, D))
                    G('gemm_proj', s.O_t, w.t_proj[0], w.t_proj[1], s.P_t, EPI_BF16)
                    G('gemm_proj', s.P_t, w.t_fc[0], w.t_fc[1], X[:M], EPI_F32_ADD)
                # ---- spatial attention + residual (vit.py:179-215)
                s.A_s, s.xh_s, s.rs_s = ln_save(Rs, w.n1)
                s.QKV_s, s.O_s = e((Rs, 3 * D)), e((Rs, D))
                s.OCLS = e((B, T, D), torch.float32) if use_cls else None
                s.LSE = e((B * T * HEADS, 304), torch.float32)
                G('gemm_qkv', s.A_s, w.s_qkv[0], w.s_qkv[1], s.QKV_s, EPI_BF16)
                L('attn_spatial', ops.attn_spatial_train, s.QKV_s, s.O_s, s.OCLS, s.LSE, B, N, T, HEADS, use_cls, M)
                if use_cls and causal == 0:
                    # mean over frames of the (per-frame dropped) cls outputs (vit.py:186-195)
                    ocls = s.OCLS if dp is None else s.OCLS * dp['ss'].view(B, T, 1)
                    L('cls_merge', ops.cls_merge, ocls, s.O_s, B, T, D, M, 0)
                if dp is not None:
                    L('gemm_proj', ops.gemm_add_scaled, s.O_s, w.s_proj[0], w.s_proj[1], None, dp['rs_s'], dp['bs_s'],
                      X[:Rs], flops=2.0 * Rs * D * D)
                else:
                    G('gemm_proj', s.O_s, w.s_proj[0], w.s_proj[1], X[:Rs], EPI_F32_ADD)
                # ---- MLP on every token incl. cls (vit.py:216)
                s.A_m, s.xh_m, s.rs_m = ln_save(R, w.n2)
                s.Z, s.H = e((R, 4 * D)), e((R, 4 * D))
                L('gemm_fc1', ops.gemm_aux, s.A_m, w.fc1[0], w.fc1[1], s.H, s.Z, EPI_BF16_GELU_AUX,
                  flops=2.0 * R * 4 * D * D)
                if dp is not None:
                    L('gemm_fc2', ops.gemm_add_scaled, s.H, w.fc2[0], w.fc2[1], None, dp['rs_m'], dp['rs_m'], X,
                      flops=2.0 * R * 4 * D * D)
                else:
                    G('gemm_fc2', s.H, w.fc2[0], w.fc2[1], X, EPI_F32_ADD)
                sv.blocks.append(s)
            # ---- head (vision_tf.py:152-153, mask_tracker.py:112-137)
            if mod.norm_embeddings:
                sv.A_f, sv.xh_f, sv.rs_f = ln_save(M, pk.norm)
            else:
                sv.A_f = e((M, D))
                L('ln', ops.layernorm, X[:M], None, None, sv.A_f)
            G('gemm_head', sv.A_f, pk.head[0], pk.head[1], LOW[:M], EPI_F32_STORE)
            C, F = mod.output_channels, mod.flag_channels
            out_mask = torch.empty((B, C, T, Hf, Wf), device=device, dtype=torch.float32)
            out_flags = torch.empty((B, T, F), device=device, dtype=torch.float32) if F > 0 else None
            sv.mode = 1 if (mod.track_map_resize == 'nearest' or pk.stride == 1) else 0
            L('mask_upsample', ops.mask_upsample, LOW[:M], out_mask, B, T, Ho, Wo, C, pk.pp, pk.stride, sv.mode)
            if out_flags is not None:
                L('flag_mean', ops.flag_mean, LOW[:M], out_flags, B, N, T, F, pk.flag_col0)
            sv.C, sv.F, sv.norm_embeddings = C, F, bool(mod.norm_embeddings)
            sv.device = device
        return out_mask, out_flags, sv

    # ------------------------------------------------------------------ backward
    def backward(self, mod, sv, d_mask, d_flags):
        """Returns {reference parameter name (relative to the QueryMaskTracker): fp32 gradient}."""
        B, N, T, D, P, Ho, Wo, M, R, Rs, Kp, use_cls, causal, causal_diag = sv.dims
        pk, device = sv.pk, sv.device
        merged = self.merge_temporal_proj
        with torch.cuda.device(device), torch.no_grad():
            sc = self._scratch_for(device, R, D, pk.n_pad)
            G_, Gb, dA, dO, dQKV, dZ, dLOW, ws = (sc[k] for k in ('G', 'Gb', 'dA', 'dO', 'dQKV', 'dZ', 'dLOW', 'ws'))
            lay = _GradLayout(len(pk.blocks), D, Kp, N + 1, T, pk.n_pad, merged)
            sv.blocks_dp = [None] * len(pk.blocks)
            flat = torch.zeros(lay.total, device=device, dtype=torch.float32)
            gv = lambda name: lay.view(flat, name)
            L, G, WG = self._launch, self._gemm, self._wgrad
            sync = self.grad_sync
            if sync is not None:
                sync.begin(flat)

            def colsum(x, name):
                L('colsum', ops.colsum, x, gv(name), ws, True, nbytes=2.0 * x.numel())

            gs_tag = [None]     # which branch the row-scaled copy in sc['Gs'] currently belongs to (stochastic depth)

            def ln_bwd(rows, xh, rs, gamma, gname, bname, accumulate=True, next_scale=None, tag=None):
                # next_scale: the NEXT branch of the backward runs under DropPath: write its row-scaled dY here too
                if next_scale is None:
                    L('ln_bwd', ops.layernorm_bwd, dA[:rows], xh, rs, gamma, G_[:rows], Gb[:rows], gv(gname), gv(bname), ws,
                      accumulate, nbytes=(14.0 if accumulate else 10.0) * rows * D)
                else:
                    L('ln_bwd', ops.layernorm_bwd, dA[:rows], xh, rs, gamma, G_[:rows], Gb[:rows], gv(gname), gv(bname), ws,
                      accumulate, next_scale, sc['Gs'][:rows], nbytes=(16.0 if accumulate else 12.0) * rows * D)
                    gs_tag[0] = tag

            # ---- mask head (mask_tracker.py:112-137) and the optional final norm (vision_tf.py:152-153)
            d_mask = d_mask.to(torch.float32).contiguous()
            if d_flags is not None:
                d_flags = d_flags.to(torch.float32).contiguous()
            L('mask_head_bwd', ops.mask_head_bwd, d_mask, d_flags, dLOW[:M], B, T, Ho, Wo, sv.C, pk.pp, pk.stride, sv.mode,
              max(sv.F, 0), pk.flag_col0, nbytes=4.0 * d_mask.numel())
            WG('wgrad_head', dLOW[:M], sv.A_f, gv('head_w'), gv('head_b'))
            if sv.norm_embeddings:
                G('dgrad_head', dLOW[:M], pk.head[2], None, dA[:M], EPI_BF16)
                ln_bwd(M, sv.xh_f, sv.rs_f, pk.norm[0], 'norm_g', 'norm_b', accumulate=False)
            else:
                G('dgrad_head', dLOW[:M], pk.head[2], None, G_[:M], EPI_F32_STORE)
                L('ln', ops.layernorm, G_[:M], None, None, Gb[:M])
            G_[M:].zero_()
            Gb[M:].zero_()
            if sync is not None:
                sync.ready(*lay.head_range)

            for bi in reversed(range(len(pk.blocks))):
                w, s = pk.blocks[bi], sv.blocks[bi]
                p = f'b{bi}.'
                dp = s.dp
                Gs = Gb                 # branch gradient = residual gradient, row-scaled under stochastic depth

                def scaled(rows, scale, tag):
                    if gs_tag[0] != tag:      # not produced by the preceding LayerNorm backward: one pass over Gb
                        L('scale_rows', ops.scale_rows, Gb[:rows], scale, sc['Gs'][:rows], nbytes=4.0 * rows * D)
                    gs_tag[0] = None
                    return sc['Gs']
                # ---- MLP (vit.py:216)
                if dp is not None:
                    Gs = scaled(R, dp['rs_m'], ('m', bi))
                WG('wgrad_fc2', Gs, s.H, gv(p + 'fc2_w'), gv(p + 'fc2_b'))
                L('dgrad_fc2', ops.gemm_aux, Gs, w.fc2[2], None, dZ, s.Z, EPI_BF16_DGELU, flops=2.0 * R * 4 * D * D)
                WG('wgrad_fc1', dZ, s.A_m, gv(p + 'fc1_w'), gv(p + 'fc1_b'))
                G('dgrad_fc1', dZ, w.fc1[2], None, dA, EPI_BF16)
                if dp is not None:
                    ln_bwd(R, s.xh_m, s.rs_m, w.n2[0], p + 'n2_g', p + 'n2_b', next_scale=dp['rs_s_pad'], tag=('s', bi))
                else:
                    ln_bwd(R, s.xh_m, s.rs_m, w.n2[0], p + 'n2_g', p + 'n2_b')
                # ---- spatial attention (vit.py:179-215)
                Gs = Gb
                if dp is not None:
                    Gs = scaled(Rs, dp['rs_s'], ('s', bi))
                    colsum(Gs[:M], p + 's_proj_b')
                    if use_cls:      # cls rows: the bias term carries its own scale (mean of the frame scales, causal==0)
                        gv(p + 's_proj_b').add_((dp['bs_s'][M:R, None] * Gb[M:R].float()).sum(0))
                WG('wgrad_proj', Gs[:Rs], s.O_s, gv(p + 's_proj_w'), None if dp is not None else gv(p + 's_proj_b'))
                G('dgrad_proj', Gs[:Rs], w.s_proj[2], None, dO[:Rs], EPI_BF16)
                dOCLS = None
                dCLS = torch.empty((ops.spatial_bwd_scratch_floats(B, T, HEADS),), device=device, dtype=torch.float32)
                if use_cls:
                    dOCLS = torch.empty((B, T, D), device=device, dtype=torch.float32)
                    L('cls_merge_bwd', ops.cls_merge_bwd, dO, dOCLS, B, T, D, M, 0 if causal == 0 else 1)
                    if dp is not None and causal == 0:
                        dOCLS.mul_(dp['ss'].view(B, T, 1))
                S = N + (1 if use_cls else 0)
                L('attn_spatial_bwd', ops.attn_spatial_bwd, s.QKV_s, s.O_s, s.OCLS, dO, dOCLS, s.LSE, dQKV, dCLS, B, N, T,
                  HEADS, use_cls, M, flops=10.0 * B * T * HEADS * S * S * 64, nbytes=16.0 * M * D)
                WG('wgrad_qkv', dQKV[:Rs], s.A_s, gv(p + 's_qkv_w'), gv(p + 's_qkv_b'))
                G('dgrad_qkv', dQKV[:Rs], w.s_qkv[2], None, dA[:Rs], EPI_BF16)
                if dp is not None and merged:
                    ln_bwd(Rs, s.xh_s, s.rs_s, w.n1[0], p + 'n1_g', p + 'n1_b', next_scale=dp['rs_t_pad'], tag=('t', bi))
                else:
                    ln_bwd(Rs, s.xh_s, s.rs_s, w.n1[0], p + 'n1_g', p + 'n1_b')
                # ---- temporal attention + temporal_fc (vit.py:169-176)
                if merged:
                    Gs = Gb
                    if dp is not None:
                        Gs = scaled(M, dp['rs_t'], ('t', bi))
                        colsum(Gb[:M], p + 't_out_b2')          # temporal_fc.bias sits outside DropPath
                    WG('wgrad_proj', Gs[:M], s.O_t, gv(p + 't_out_w'), gv(p + 't_out_b'))
                    G('dgrad_proj', Gs[:M], w.t_out[2], None, dO[:M], EPI_BF16)
                else:
                    WG('wgrad_proj', Gb[:M], s.P_t, gv(p + 't_fc_w'), gv(p + 't_fc_b'))
                    G('dgrad_proj', Gb[:M], w.t_fc[2], None, dA[:M], EPI_BF16)      # d(proj output)
                    WG('wgrad_proj', dA[:M], s.O_t, gv(p + 't_proj_w'), gv(p + 't_proj_b'))
                    G('dgrad_proj', dA[:M], w.t_proj[2], None, dO[:M], EPI_BF16)
                L('attn_temporal_bwd', ops.attn_temporal_bwd, s.QKV_t, s.O_t, dO, dQKV, B * N, T, HEADS, causal_diag,
                  flops=10.0 * B * N * HEADS * T * T * 64, nbytes=16.0 * M * D)
                WG('wgrad_qkv', dQKV[:M], s.A_t, gv(p + 't_qkv_w'), gv(p + 't_qkv_b'))
                G('dgrad_qkv', dQKV[:M], w.t_qkv[2], None, dA[:M], EPI_BF16)
                dp_prev = sv.blocks[bi - 1].dp if bi > 0 else None
                if dp_prev is not None:   # the MLP branch of block bi-1 is next: its scale for the patch rows here, cls rows apart
                    ln_bwd(M, s.xh_t, s.rs_t, w.tn1[0], p + 'tn1_g', p + 'tn1_b', next_scale=dp_prev['rs_m'], tag=('m', bi - 1))
                    sc['Gs'][M:R].copy_((dp_prev['rs_m'][M:R, None] * Gb[M:R].float()).to(torch.bfloat16))
                else:
                    ln_bwd(M, s.xh_t, s.rs_t, w.tn1[0], p + 'tn1_g', p + 'tn1_b')
                sv.blocks_dp[bi] = s
                sv.blocks[bi] = None           # this block's activations are no longer needed
                if sync is not None:
                    sync.ready(*lay.block_ranges[len(pk.blocks) - 1 - bi])
            # ---- embeddings (vision_tf.py:99-138) and the patch projection (vit.py:233-241)
            L('embed_bwd', ops.embed_bwd, G_, gv('pos'), gv('time'), gv('pos')[0], ws, B, N, T, D, True,
              nbytes=8.0 * M * D)
            WG('wgrad_patch', Gb[:M], sv.PM, gv('patch_w'))
            if sync is not None:
                sync.ready(*lay.embed_range)
                sync.finish()
            self.last_flat = flat
            dropped = [b is not None and b.dp is not None for b in sv.blocks_dp]
            return self._unpack(mod, pk, lay, flat, merged, dropped)

    # ------------------------------------------------------------------ packed gradients -> reference parameters
    def _unpack(self, mod, pk, lay, flat, merged, dropped=None):
        gv = lambda name: lay.view(flat, name)
        D = gv('norm_g').shape[0]
        g = {}
        pre = 'tracker_backbone.timesformer.model.'
        g[pre + 'patch_embed.proj.weight'] = gv('patch_w').view(D, 4, mod.patch_size, mod.patch_size)
        g[pre + 'patch_embed.proj.bias'] = gv('time').sum(0)            # every token row carries the conv bias once
        g[pre + 'pos_embed'] = gv('pos')[None]
        g[pre + 'time_embed'] = gv('time')[None]
        g[pre + 'cls_token'] = gv('pos')[0].reshape(1, 1, D).clone()
        if mod.norm_embeddings:
            g[pre + 'norm.weight'], g[pre + 'norm.bias'] = gv('norm_g'), gv('norm_b')
        else:
            # the final norm is not in the reference's graph then (vision_tf.py:152-153): its .grad stays None there, so
            # AdamW's decoupled weight decay (train.py:233) must not see a zero gradient here either
            g[pre + 'norm.weight'] = g[pre + 'norm.bias'] = None
        for i, w in enumerate(pk.blocks):
            p, q = f'b{i}.', pre + f'blocks.{i}.'
            for ours, theirs in (('tn1', 'temporal_norm1'), ('n1', 'norm1'), ('n2', 'norm2')):
                g[q + theirs + '.weight'], g[q + theirs + '.bias'] = gv(p + ours + '_g'), gv(p + ours + '_b')
            for ours, theirs in (('t_qkv', 'temporal_attn.qkv'), ('s_qkv', 'attn.qkv'), ('s_proj', 'attn.proj'),
                                 ('fc1', 'mlp.fc1'), ('fc2', 'mlp.fc2')):
                g[q + theirs + '.weight'], g[q + theirs + '.bias'] = gv(p + ours + '_w'), gv(p + ours + '_b')
            if not merged:
                g[q + 'temporal_fc.weight'], g[q + 'temporal_fc.bias'] = gv(p + 't_fc_w'), gv(p + 't_fc_b')
                g[q + 'temporal_attn.proj.weight'], g[q + 'temporal_attn.proj.bias'] = gv(p + 't_proj_w'), gv(p + 't_proj_b')
        if merged:
            # W_m = Wf Wp, b_m = Wf bp + bf  =>  dWf = dW_m Wp^T + db_1 bp^T, dWp = Wf^T dW_m, dbp = Wf^T db_1, dbf = db_2
            # (under stochastic depth the Wf bp part of the bias is inside DropPath, bf outside: two bias gradients;
            # without it db_2 = db_1).  All 12 blocks in four batched fp32 products.
            nb = len(pk.blocks)
            Wp, bp, Wf, _ = pk.raw_t_all
            dWm = torch.stack([gv(f'b{i}.t_out_w') for i in range(nb)])
            dbm = torch.stack([gv(f'b{i}.t_out_b') for i in range(nb)])
            dbf = torch.stack([gv(f'b{i}.t_out_b2') if (dropped is not None and dropped[i]) else gv(f'b{i}.t_out_b')
                               for i in range(nb)])
            dWf = _mm(dWm, Wp.transpose(1, 2)) + dbm[:, :, None] * bp[:, None, :]
            dWp = _mm(Wf.transpose(1, 2), dWm)
            dbp = _mm(Wf.transpose(1, 2), dbm[:, :, None])[:, :, 0]
            for i in range(nb):
                q = pre + f'blocks.{i}.'
                g[q + 'temporal_fc.weight'], g[q + 'temporal_fc.bias'] = dWf[i], dbf[i]
                g[q + 'temporal_attn.proj.weight'], g[q + 'temporal_attn.proj.bias'] = dWp[i], dbp[i]
        # head: undo the avg-pool fold (each pooled row is the mean of stride^2 rows of tracker_post_linear)
        C, pp, s = mod.output_channels, pk.pp, pk.stride
        n_mask = C * pp * pp
        dWh, dbh = gv('head_w'), gv('head_b')
        dWt = (dWh[:n_mask].reshape(C, pp, 1, pp, 1, D) / (s * s)).expand(C, pp, s, pp, s, D).reshape(C * pp * s * pp * s, D)
        dbt = (dbh[:n_mask].reshape(C, pp, 1, pp, 1) / (s * s)).expand(C, pp, s, pp, s).reshape(-1)
        g['tracker_post_linear.weight'], g['tracker_post_linear.bias'] = dWt.contiguous(), dbt.contiguous()
        if mod.flag_channels > 0:
            F = mod.flag_channels
            g['flag_post_linear.weight'] = dWh[n_mask:n_mask + F]
            g['flag_post_linear.bias'] = dbh[n_mask:n_mask + F]
        return g


class SeekerFunction(torch.autograd.Function):
    """autograd node of one Seeker forward: (frames, query, *parameters) -> (mask logits, flags)."""

    @staticmethod
    def forward(ctx, engine, mod, names, queries_per_video, input_frames, query_mask, *params):
        out_mask, out_flags, sv = engine.forward(mod, input_frames, query_mask, queries_per_video)
        ctx.engine, ctx.mod, ctx.names, ctx.sv = engine, mod, names, sv
        ctx.has_flags = out_flags is not None
        if out_flags is None:
            out_flags = out_mask.new_zeros(())
            ctx.mark_non_differentiable(out_flags)
        return out_mask, out_flags

    @staticmethod
    def backward(ctx, d_mask, d_flags):
        if ctx.sv is None:
            raise RuntimeError('tcow_b200: backward through the same Seeker forward twice is not supported')
        grads = ctx.engine.backward(ctx.mod, ctx.sv, d_mask, d_flags if ctx.has_flags else None)
        ctx.sv = None
        out = []
        for i, name in enumerate(ctx.names):
            need = ctx.needs_input_grad[6 + i]
            out.append(grads.get(name) if need else None)
        return (None, None, None, None, None, None, *out)
