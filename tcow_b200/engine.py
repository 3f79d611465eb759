"""The Seeker forward as a plan of sm_100a kernels (the product path; no eager / CPU fallback).

Executes what the reference runs in model/mask_tracker.py:92-142 -> model/vision_tf.py:68-169 ->
third_party/TimeSformer/timesformer/models/vit.py:155-217 (Block.forward, 12x), through the C ABI in
include/tcow_b200.h.  Data layout in HBM:

  X    fp32 [M+B, D]   residual stream, canonical token rows r=(b*N+n)*T+t, the B cls rows last
  A    bf16 [M+B, D]   LayerNorm output (GEMM A operand)
  QKV  bf16 [M+B, 3D]  qkv projections, columns [q|k|v] head-major
  O    bf16 [M+B, D]   attention output (GEMM A operand)
  H    bf16 [M+B, 4D]  GELU(fc1) (GEMM A operand); also holds the im2col patch matrix [M, 4*P*P]
  LOW  fp32 [M, Npad]  pooled mask-head patch values + flag logits per token

No tensor is ever transposed: the temporal kernels see T consecutive rows per (b,n), the spatial
attention gathers rows with stride T, every GEMM is row-order agnostic.
"""
from __future__ import annotations

import os
import threading

import torch

from . import _lib, ops
from .ops import EPI_BF16, EPI_BF16_GELU, EPI_F32_ADD, EPI_F32_STORE

HEADS = 12


class _Packed:
    """Device-resident bf16 / folded copies of one module's parameters."""
    pass


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class SeekerEngine:
    def __init__(self, tracker, max_chunk=8, merge_temporal_proj=True):
        self.max_chunk = max_chunk
        # temporal_fc o temporal_attn.proj has no nonlinearity in between (vit.py:111 -> :174): at inference
        # the two 768x768 linears are pre-multiplied in fp32 into one.
        self.merge_temporal_proj = merge_temporal_proj
        # LayerNorm stays a stand-alone bandwidth kernel: both fusions were built and measured slower or equal — a read-back
        # tail inside the residual GEMM (round 1, -7 %) and the deferred form with statistics applied in the next GEMM's
        # epilogue (round 2, -0.3 %); numbers in profiles/r01_notes.md and profiles/r02_notes.md.
        # Patch embedding as one kernel (gather + implicit GEMM + embeddings, csrc/patch_embed_fused.cu) whenever the shape
        # allows it (patch 16, T <= 32); TCOW_FUSE_PATCH=0 forces the three-kernel form (gather, embed_init, reduce-add GEMM).
        self.fuse_patch_embed = os.environ.get('TCOW_FUSE_PATCH', '1') != '0'
        self._lock = threading.Lock()
        self._packed = {}      # device index -> (stamp, _Packed)
        self._workspace = {}   # (device index, Bc, shape key) -> dict of tensors
        self.launches = 0      # kernels launched by the last forward (bench.py's gpu_launches)
        self.profile = None    # set to a list to collect (kind, flops, bytes, start_event, end_event) per launch
        # The ~135 launches between the input gather and the output upsample touch only engine-owned buffers, so they
        # are captured once per (device, chunk shape, weights) into a CUDA graph and replayed (no launch gaps).
        self.use_cuda_graph = os.environ.get('TCOW_CUDA_GRAPH', '1') != '0'
        self._graphs = {}      # key -> (torch.cuda.CUDAGraph, launches) ; key -> int warm-up count in _warm
        self._warm = {}
        self.max_cached_shapes = 4   # graphs / workspaces kept per device (LRU); nn.DataParallel threads share the engine

    # ------------------------------------------------------------------ weights
    @staticmethod
    def _stamp(mod):
        # walk the modules' own tensors rather than parameters(): nn.DataParallel replicas (train.py:223) report no
        # parameters() — replicate() leaves them an empty _parameters and the broadcast copies in _former_parameters —
        # and a stale stamp would mean stale packed weights
        def tensors(m):
            yield from m._parameters.values()
            yield from getattr(m, '_former_parameters', {}).values()
        return tuple((p.data_ptr(), p._version) for m in mod.modules() for p in tensors(m) if p is not None)

    def _pack(self, mod, device):
        bb = mod.tracker_backbone.timesformer.model
        D = bb.embed_dim
        P = mod.patch_size
        bf = lambda t: _f32(t, device).to(torch.bfloat16).contiguous()
        pk = _Packed()
        pk.patch_w = bf(bb.patch_embed.proj.weight.reshape(D, -1))           # [D, 4*P*P], K = (c, r, w)
        pk.patch_b = _f32(bb.patch_embed.proj.bias, device)
        pk.pos = _f32(bb.pos_embed, device).reshape(-1, D)
        pk.time = _f32(bb.time_embed, device).reshape(-1, D)
        pk.cls = _f32(bb.cls_token, device).reshape(D)
        pk.blocks = []
        for blk in bb.blocks:
            w = _Packed()
            w.tn1 = (_f32(blk.temporal_norm1.weight, device), _f32(blk.temporal_norm1.bias, device))
            w.n1 = (_f32(blk.norm1.weight, device), _f32(blk.norm1.bias, device))
            w.n2 = (_f32(blk.norm2.weight, device), _f32(blk.norm2.bias, device))
            w.t_qkv = (bf(blk.temporal_attn.qkv.weight), _f32(blk.temporal_attn.qkv.bias, device))
            Wp, bp = _f32(blk.temporal_attn.proj.weight, device), _f32(blk.temporal_attn.proj.bias, device)
            Wf, bfc = _f32(blk.temporal_fc.weight, device), _f32(blk.temporal_fc.bias, device)
            if self.merge_temporal_proj:
                # fc(proj(o)) = o (Wf Wp)^T + (Wf bp + bf); exact fp32 products (TF32 off), then one bf16 rounding
                prev = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                try:
                    w.t_out = ((Wf @ Wp).to(torch.bfloat16).contiguous(), (Wf @ bp + bfc).contiguous())
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = prev
                w.t_proj = w.t_fc = None
            else:
                w.t_out = None
                w.t_proj = (Wp.to(torch.bfloat16).contiguous(), bp)
                w.t_fc = (Wf.to(torch.bfloat16).contiguous(), bfc)
            w.s_qkv = (bf(blk.attn.qkv.weight), _f32(blk.attn.qkv.bias, device))
            w.s_proj = (bf(blk.attn.proj.weight), _f32(blk.attn.proj.bias, device))
            w.fc1 = (bf(blk.mlp.fc1.weight), _f32(blk.mlp.fc1.bias, device))
            w.fc2 = (bf(blk.mlp.fc2.weight), _f32(blk.mlp.fc2.bias, device))
            pk.blocks.append(w)
        pk.norm = (_f32(bb.norm.weight, device), _f32(bb.norm.bias, device))
        # Head: fold avg_pool2d(stride) into tracker_post_linear (mask_tracker.py:113-122; pooling is linear and
        # acts inside one patch because stride | P), append the flag rows, pad N to a multiple of 64.
        C, s = mod.output_channels, max(int(mod.track_map_stride), 1)
        if P % s != 0:
            raise NotImplementedError(f'track_map_stride={s} must divide patch_size={P}')
        pp = P // s
        Wt, bt = _f32(mod.tracker_post_linear.weight, device), _f32(mod.tracker_post_linear.bias, device)
        Wt = Wt.reshape(C, pp, s, pp, s, D).mean((2, 4)).reshape(C * pp * pp, D)
        bt = bt.reshape(C, pp, s, pp, s).mean((2, 4)).reshape(-1)
        rows, biases = [Wt], [bt]
        F = mod.flag_channels
        if F > 0:
            rows.append(_f32(mod.flag_post_linear.weight, device))
            biases.append(_f32(mod.flag_post_linear.bias, device))
        n_used = C * pp * pp + max(F, 0)
        n_pad = (n_used + 63) // 64 * 64
        Wh = torch.zeros(n_pad, D, device=device, dtype=torch.float32)
        bh = torch.zeros(n_pad, device=device, dtype=torch.float32)
        Wh[:n_used] = torch.cat(rows, 0)
        bh[:n_used] = torch.cat(biases, 0)
        pk.head_w, pk.head_b = Wh.to(torch.bfloat16).contiguous(), bh
        pk.pp, pk.stride, pk.flag_col0, pk.n_pad = pp, s, C * pp * pp, n_pad
        return pk

    def packed(self, mod, device):
        stamp = (self._stamp(mod), self.merge_temporal_proj)
        with self._lock:
            hit = self._packed.get(device.index)
            if hit is not None and hit[0] == stamp:
                return hit[1]
        with torch.no_grad():
            pk = self._pack(mod, device)
        with self._lock:
            self._packed[device.index] = (stamp, pk)
            self._graphs = {k: v for k, v in self._graphs.items() if k[0] != device.index}   # they hold the old weights
        return pk

    def invalidate(self):
        """Forget the packed weights and the captured graphs.  The cache key is (data_ptr, _version) of every parameter,
        which in-place writes through `.data` (p.data.copy_(), an EMA swap) do not change — call this after such a write."""
        with self._lock:
            self._packed.clear()
            self._graphs.clear()
            self._warm.clear()

    def release(self):
        """invalidate() and drop the activation workspaces as well (returns the memory to the caching allocator)."""
        self.invalidate()
        with self._lock:
            self._workspace.clear()

    def _ws(self, device, Bc, N, T, D, K_patch, n_pad):
        key = (device.index, Bc, N, T, D, K_patch, n_pad)
        with self._lock:
            ws = self._workspace.pop(key, None)
            if ws is not None:
                self._workspace[key] = ws                      # most recently used last
        if ws is None:
            M = Bc * N * T
            R = M + Bc
            e = lambda shape, dt: torch.empty(shape, device=device, dtype=dt)
            ws = dict(X=e((R, D), torch.float32), A=e((R, D), torch.bfloat16), QKV=e((R, 3 * D), torch.bfloat16),
                      O=e((R, D), torch.bfloat16), H=e((R * max(4 * D, K_patch),), torch.bfloat16),
                      OCLS=e((Bc, T, D), torch.float32), LOW=e((M, n_pad), torch.float32))
            with self._lock:
                self._workspace[key] = ws
                self._evict(self._workspace, device.index, self.max_cached_shapes)
        return ws

    @staticmethod
    def _evict(cache, device_index, limit):
        """Least-recently-used eviction per device (dict order = use order); caller holds the lock."""
        mine = [k for k in cache if k[0] == device_index]
        for k in mine[:max(0, len(mine) - limit)]:
            del cache[k]

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

    # ------------------------------------------------------------------ forward
    def forward(self, mod, input_frames, query_mask, queries_per_video=1, frame_scale=1.0):
        """input_frames (V,3,T,Hf,Wf); query_mask (V*queries_per_video,1,T,Hf,Wf), sample s belongs to video
        s // queries_per_video.  queries_per_video=1 is the reference's Seeker.forward contract.
        fp32 and uint8 inputs are read as they are (no staging copy); other dtypes are cast to fp32 first, as
        mask_tracker.py:103-104 does.  frame_scale multiplies the RGB values in the gather kernel: 1/255 turns decoder-style
        uint8 frames into the [0,1] floats data/data_plugin.py:174 builds on the host."""
        if not input_frames.is_cuda:
            raise RuntimeError('tcow_b200 runs on a CUDA sm_100 device only; move the module and inputs to '
                               'the GPU (there is no CPU fallback)')
        if torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters()):
            raise RuntimeError('SeekerEngine is the inference plan; gradient mode goes through '
                               'tcow_b200.train_engine (QueryMaskTracker.forward dispatches on torch.is_grad_enabled())')
        device = input_frames.device
        bbm = mod.tracker_backbone
        V, Cin, T, Hf, Wf = input_frames.shape
        B = V * queries_per_video
        if Cin != 3:
            raise RuntimeError(f'expected 3 RGB channels (+1 query channel = in_chans 4), got {Cin}')
        if tuple(query_mask.shape) != (B, 1, T, Hf, Wf):
            raise RuntimeError(f'query_mask shape {tuple(query_mask.shape)} does not match frames {tuple(input_frames.shape)}')
        assert T == bbm.T                                   # vision_tf.py:96
        assert Hf == bbm.Hf and Wf == bbm.Wf                # pos_embed is built for this grid (vision_tf.py:97)
        P, D = mod.patch_size, bbm.output_feature_dim
        Ho, Wo = Hf // P, Wf // P
        N = Ho * Wo
        causal = int(mod.causal_attention)
        if causal in (0, 1):
            use_cls = True
        elif causal >= 2 or causal == -1:
            use_cls = False
        else:
            raise ValueError(f'unsupported causal_attention={causal}')  # vit.py:179-208 has no branch either
        causal_diag = -1 if causal <= 0 else (0 if causal <= 2 else causal - 2)   # vit.py:93-99
        if mod.track_map_resize not in ('bilinear', 'nearest'):
            # mask_tracker.py:124-130 silently skips the resize for other values; reject instead of guessing
            raise ValueError(f'unsupported track_map_resize={mod.track_map_resize!r}')

        with torch.cuda.device(device):
            _lib.call('tcow_check_device')
            pk = self.packed(mod, device)
            as_is = (torch.float32, torch.uint8)
            frames = (input_frames if input_frames.dtype in as_is else input_frames.to(torch.float32)).contiguous()
            query = (query_mask if query_mask.dtype in as_is else query_mask.to(torch.float32)).contiguous()
            C, F = mod.output_channels, mod.flag_channels
            out_mask = torch.empty((B, C, T, Hf, Wf), device=device, dtype=torch.float32)
            out_flags = torch.empty((B, T, F), device=device, dtype=torch.float32) if F > 0 else None
            self.launches = 0
            for b0 in range(0, B, self.max_chunk):
                b1 = min(B, b0 + self.max_chunk)
                self._run_chunk(mod, pk, frames, query[b0:b1], out_mask[b0:b1],
                                None if out_flags is None else out_flags[b0:b1],
                                N, T, D, P, Ho, Wo, use_cls, causal, causal_diag, queries_per_video, b0, frame_scale)
        return out_mask, out_flags

    def _run_chunk(self, mod, pk, frames, query, out_mask, out_flags, N, T, D, P, Ho, Wo, use_cls, causal,
                   causal_diag, qpv=1, sample0=0, frame_scale=1.0):
        Bc = query.shape[0]
        M = Bc * N * T
        R = M + Bc
        Kp = 4 * P * P
        ws = self._ws(query.device, Bc, N, T, D, Kp, pk.n_pad)
        X, A, QKV, O, OCLS, LOW = ws['X'], ws['A'], ws['QKV'], ws['O'], ws['OCLS'], ws['LOW']
        H = ws['H'][:R * 4 * D].view(R, 4 * D)
        PM = ws['H'][:M * Kp].view(M, Kp)
        L, G = self._launch, self._gemm
        # ---- patch embedding + embeddings (mask_tracker.py:107-108, vit.py:235-241, vision_tf.py:99-138)
        fused_embed = self.fuse_patch_embed and P == 16 and T <= ops.PATCH_EMBED_FUSED_MAX_T and D % 256 == 0
        in_bytes = (3.0 * frames.element_size() / qpv + query.element_size()) * M * Kp / 4
        if fused_embed:
            L('patch_embed', ops.patch_embed_fused, frames, query, pk.patch_w, pk.patch_b, pk.pos, pk.time, pk.cls, X, P,
              bool(mod.tracker_backbone.pretrained), qpv, sample0, frame_scale, flops=2.0 * M * D * Kp,
              nbytes=in_bytes + 4.0 * R * D)
        else:
            L('patch_gather', ops.patch_gather, frames, query, PM, P, bool(mod.tracker_backbone.pretrained), qpv, sample0,
              frame_scale, nbytes=in_bytes + 2.0 * M * Kp)
        # ---- everything from the embeddings to the head GEMM: engine-owned buffers only -> CUDA-graph replay
        key = (query.device.index, Bc, N, T, use_cls, causal, causal_diag, bool(mod.norm_embeddings), id(pk), fused_embed)
        core = lambda: self._core(mod, pk, ws, Bc, M, R, N, T, D, Kp, use_cls, causal, causal_diag, fused_embed)
        if self.use_cuda_graph and self.profile is None:
            with self._lock:
                hit = self._graphs.pop(key, None)
                if hit is not None:
                    self._graphs[key] = hit                  # most recently used last
                warm = self._warm.get(key, 0)
            if hit is not None:
                hit[0].replay()
                self.launches += hit[1]
            elif warm < 1:
                core()                                   # first pass eager: kernel attributes get configured
                with self._lock:
                    self._warm[key] = 1
            else:
                n0 = self.launches
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode='thread_local'):
                    core()
                with self._lock:
                    self._graphs[key] = (g, self.launches - n0, ws)   # the graph's kernels point into ws: keep it alive
                    self._evict(self._graphs, key[0], self.max_cached_shapes)
                g.replay()
        else:
            core()
        mode = 1 if (mod.track_map_resize == 'nearest' or pk.stride == 1) else 0
        L('mask_upsample', ops.mask_upsample, LOW, out_mask, Bc, T, Ho, Wo, mod.output_channels, pk.pp, pk.stride,
          mode, nbytes=4.0 * out_mask.numel())
        if out_flags is not None:
            L('flag_mean', ops.flag_mean, LOW, out_flags, Bc, N, T, mod.flag_channels, pk.flag_col0)

    def _core(self, mod, pk, ws, Bc, M, R, N, T, D, Kp, use_cls, causal, causal_diag, fused_embed=False):
        X, A, QKV, O, OCLS, LOW = ws['X'], ws['A'], ws['QKV'], ws['O'], ws['OCLS'], ws['LOW']
        H = ws['H'][:R * 4 * D].view(R, 4 * D)
        PM = ws['H'][:M * Kp].view(M, Kp)
        L, G = self._launch, self._gemm
        if not fused_embed:
            L('embed_init', ops.embed_init, X, pk.patch_b, pk.pos, pk.time, pk.cls, Bc, N, T, D, nbytes=4.0 * R * D)
        Rs = R if use_cls else M
        ln_bytes = lambda rows: 6.0 * rows * D
        final_ln = pk.norm if mod.norm_embeddings else (None, None)

        def residual(kind, a, wb, rows, ln_params, ln_rows):
            """X[:rows] += a @ W^T + b (TMA reduce-add epilogue), then A[:ln_rows] = LN(X[:ln_rows])."""
            G(kind, a, wb[0], wb[1], X[:rows], EPI_F32_ADD)
            L('ln', ops.layernorm, X[:ln_rows], ln_params[0], ln_params[1], A[:ln_rows], nbytes=ln_bytes(ln_rows))

        nblk = len(pk.blocks)
        if fused_embed:     # X already holds the embedded tokens: only temporal_norm1 of block 0 remains
            L('ln', ops.layernorm, X[:M], pk.blocks[0].tn1[0], pk.blocks[0].tn1[1], A[:M], nbytes=ln_bytes(M))
        else:
            residual('gemm_patch', PM, (pk.patch_w, None), M, pk.blocks[0].tn1, M)   # + temporal_norm1 of block 0
        for bi, w in enumerate(pk.blocks):
            # temporal attention + temporal_fc + residual (vit.py:169-176); cls rows untouched.  A = LN_tn1(X[:M]).
            G('gemm_qkv', A[:M], w.t_qkv[0], w.t_qkv[1], QKV[:M], EPI_BF16)
            L('attn_temporal', ops.attn_temporal, QKV, O, Bc * N, T, HEADS, causal_diag,
              flops=4.0 * Bc * N * HEADS * T * T * 64, nbytes=8.0 * M * D)
            # norm1 for the spatial branch; the cls rows enter the spatial attention through norm1 too (vit.py:180-186):
            # they follow the patch rows in X, so the stand-alone LayerNorm covers them in the same launch
            if w.t_out is not None:
                residual('gemm_proj', O[:M], w.t_out, M, w.n1, Rs)
            else:
                tmp = H[:M, :D]                                                    # H is idle here
                G('gemm_proj', O[:M], w.t_proj[0], w.t_proj[1], tmp, EPI_BF16)
                residual('gemm_proj', tmp, w.t_fc, M, w.n1, Rs)
            # spatial attention + residual (vit.py:179-215); cls is key/query 0 of every frame
            G('gemm_qkv', A[:Rs], w.s_qkv[0], w.s_qkv[1], QKV[:Rs], EPI_BF16)
            S = N + (1 if use_cls else 0)
            L('attn_spatial', ops.attn_spatial, QKV, O, OCLS if use_cls else None, Bc, N, T, HEADS, use_cls, M,
              flops=4.0 * Bc * T * HEADS * S * S * 64, nbytes=8.0 * M * D)
            if use_cls and causal == 0:   # mean over frames (vit.py:195); causal==1 takes frame 0, written in-kernel
                L('cls_merge', ops.cls_merge, OCLS, O, Bc, T, D, M, 0)
            residual('gemm_proj', O[:Rs], w.s_proj, Rs, w.n2, Rs)                  # + norm2 for the MLP
            if not use_cls:   # the cls rows skip the spatial branch but still go through the MLP (vit.py:215-216)
                L('ln', ops.layernorm, X[M:R], w.n2[0], w.n2[1], A[M:R], nbytes=ln_bytes(Bc))
            # MLP on every token incl. cls (vit.py:216); its residual GEMM carries the next block's temporal_norm1
            # (or, after the last block, the optional final norm / plain cast that feeds the head)
            G('gemm_fc1', A, w.fc1[0], w.fc1[1], H, EPI_BF16_GELU)
            nxt = pk.blocks[bi + 1].tn1 if bi + 1 < nblk else final_ln
            residual('gemm_fc2', H, w.fc2, R, nxt, M)
        # ---- head (mask_tracker.py:112-137) on A = final norm (vision_tf.py:152-153) or plain bf16 cast of X[:M]
        G('gemm_head', A[:M], pk.head_w, pk.head_b, LOW, EPI_F32_STORE)
