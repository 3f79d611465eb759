#!/usr/bin/env python
"""Benchmark of the one hot path this repo accelerates: TCOW Seeker forward, clips/s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores
  python bench.py --impl eager --steps K --warmup W        # the reference algorithm under stock eager PyTorch on the GPU

Metric (BASELINE.json): Seeker fwd clips/s at T=30, 240x320 — one "clip" = one (video clip, query) sample
through Seeker.forward (model/seeker.py:24).  Workload: configs[1] — batch 8 synthetic clips per GPU, bf16
tensor-core GEMMs with fp32 accumulate / fp32 residual stream, random-init ViT-B/16 weights (every tensor
non-trivial, tcow_b200/synth.py).  A step = one forward over one batch.  N > 1: independent replicas, each
rank its own clips, no data-path collective (weak scaling); time = max over ranks.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T, HF, WF, BATCH = 30, 240, 320, 8
# Algorithmic FLOPs of one reference forward (torch FlopCounterMode on the reference == closed form).
FLOP_PER_CLIP = 2302.61e9
SEEKER_KW = dict(num_total_frames=T, num_visible_frames=T, frame_height=HF, frame_width=WF, tracker_pretrained=False,
                 attention_type='divided_space_time', patch_size=16, causal_attention=1, norm_embeddings=False,
                 drop_path_rate=0.1, network_depth=12, track_map_stride=4, track_map_resize='bilinear',
                 query_channels=1, output_channels=3, flag_channels=3)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d.get('bf16_tflops', 1590.0), sustained=d.get('bf16_tflops_sustained', 1400.0),
                    hbm=d.get('hbm_gbs', 6650.0), source='measured')
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def dist_env():
    """(world, rank, local, device) of this process; NCCL process group when launched by torchrun."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, local, dev


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mask_iou(a, b):
    a, b = a > 0, b > 0                                     # eval/metrics.py:18
    union = (a | b).sum().item()
    return 1.0 if union == 0 else (a & b).sum().item() / union


def e2e_measure(net, rgb_h, q_h, frame_scale, steps, warm, world, dev):
    """The metric through the module's own forward with HOST buffers: every step the inputs are copied from pinned host
    memory (copy stream, double-buffered: step i+1 uploads while step i computes) and the step's full result — the
    (B,3,T,Hf,Wf) fp32 logits Seeker.forward returns plus the flags, what eval/inference.py:91 moves to the host — is
    copied back to pinned host memory on a second copy stream (double-buffered: step i's logits travel while step i+1
    computes; the host waits for result i-1 before it enqueues step i+1).  Returns (ms per step, h2d bytes, d2h bytes)."""
    import torch
    cur = torch.cuda.current_stream()
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bufs = [(torch.empty(rgb_h.shape, dtype=rgb_h.dtype, device=dev), torch.empty(q_h.shape, dtype=q_h.dtype, device=dev))
            for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    out_h = []

    def upload(i):
        k = i & 1
        with torch.cuda.stream(up):
            up.wait_event(ev_free[k])
            bufs[k][0].copy_(rgb_h, non_blocking=True)
            bufs[k][1].copy_(q_h, non_blocking=True)
            ev_ready[k].record(up)

    def run(n):
        for k in range(2):
            ev_free[k].record(cur)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            k = i & 1
            cur.wait_event(ev_ready[k])
            mask, flags = net(bufs[k][0], bufs[k][1], frame_scale=frame_scale)
            ev_free[k].record(cur)
            done = torch.cuda.Event()
            done.record(cur)
            if not out_h:
                for _ in range(2):
                    out_h.append((torch.empty(mask.shape, dtype=mask.dtype).pin_memory(),
                                  torch.empty(flags.shape, dtype=flags.dtype).pin_memory()))
            with torch.cuda.stream(down):
                down.wait_event(done)
                out_h[k][0].copy_(mask, non_blocking=True)
                out_h[k][1].copy_(flags, non_blocking=True)
                mask.record_stream(down)
                flags.record_stream(down)
                ev_out[k].record(down)
            if i >= 1:
                ev_out[k ^ 1].synchronize()          # result i-1 has arrived in host memory
        ev_out[(n - 1) & 1].synchronize()

    with torch.no_grad():
        run(max(min(warm, 3), 2))
        barrier(world)
        t0 = time.perf_counter()
        run(steps)
        barrier(world)
        ms = max_over_ranks(1e3 * (time.perf_counter() - t0), world, dev)
    h2d = int(rgb_h.numel() * rgb_h.element_size() + q_h.numel() * q_h.element_size())
    d2h = int(sum(t.numel() * t.element_size() for t in out_h[0]))
    return ms / steps, h2d, d2h, out_h[(steps - 1) & 1]


def load_traffic():
    """DRAM bytes per launch per kernel class, written by tools/ncu_traffic.py from an ncu capture of this bench."""
    p = os.path.join(ROOT, 'profiles', 'kernel_traffic.json')
    if not os.path.exists(p):
        return None
    return json.load(open(p))


def cpu_forward_timing(n_forwards, threads):
    """The reference algorithm on the host cores: the oracle restatement (fp32, torch CPU), B=1 per forward."""
    import torch
    from oracle import seeker_oracle
    from tcow_b200 import synth
    torch.set_num_threads(threads)
    sd = synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF)
    times = []
    for i in range(n_forwards):
        rgb, q = synth.make_batch([i], num_frames=T, frame_height=HF, frame_width=WF)
        t0 = time.perf_counter()
        with torch.no_grad():
            seeker_oracle.seeker_forward(sd, rgb, q, causal_attention=1)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself is
    Python/PyTorch and lives only in the build container), all host threads, one clip per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    times = cpu_forward_timing(args.warmup + args.steps, cores)[args.warmup:]
    tot = sum(times)
    val = len(times) / tot
    line = {'impl': 'reference', 'metric': 'seeker_fwd_clips_per_s', 'value': val, 'unit': 'clips/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot / len(times), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'TCOW Seeker forward, random-init ViT-B/16 divided space-time, T=30 240x320, causal '
                                   'temporal attn, batch 1 per step, fp32 on host CPU (oracle port of the reference path)'},
            'cpu_baseline': {'value': val, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{len(times)} forwards of 1 clip (T=30, 240x320), torch {torch.__version__} CPU fp32'},
            'e2e': {'value': val, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def eager_gpu_baseline(dev, rgb_d, q_d, ours_mask, steps, warm):
    """SURVEY §8(d) last row: the reference algorithm (the oracle restatement of model/seeker.py:24 — the reference tree
    itself does not exist on the GPU box) under stock eager PyTorch on the same B200, same weights and clips: fp32 with
    TF32 off (oracle grade) and torch.autocast('cuda', bfloat16) (cuBLAS/cuDNN/ATen bf16 — "the code to beat").  A
    reported baseline; nothing of it is on the product path."""
    import torch
    from oracle import seeker_oracle
    from tcow_b200 import synth
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF).items()}
    B = rgb_d.shape[0]
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out = {'kind': 'port', 'what': 'oracle/seeker_oracle.py (op-by-op restatement of the reference forward) on cuda, stock '
                                   f'eager PyTorch {torch.__version__}, batch {B}, same weights and clips'}

    def timed(fn, n_warm, n):
        for _ in range(n_warm):
            r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return r, e0.elapsed_time(e1) / n

    try:
        with torch.no_grad():
            (m32, f32), ms32 = timed(lambda: seeker_oracle.seeker_forward(sd, rgb_d, q_d, causal_attention=1), 1, max(2, steps // 4))

            def bf():
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    return seeker_oracle.seeker_forward(sd, rgb_d, q_d, causal_attention=1)
            (m16, f16), ms16 = timed(bf, max(2, warm // 2), max(3, steps // 2))
        m16 = m16.float()
        out['fp32_tf32_off'] = {'clips_per_s': B / (ms32 * 1e-3), 'ms_per_step': ms32}
        out['autocast_bf16'] = {'clips_per_s': B / (ms16 * 1e-3), 'ms_per_step': ms16,
                                'max_dlogit_vs_fp32': (m16 - m32).abs().max().item(),
                                'min_iou_vs_fp32': min(mask_iou(m16[b], m32[b]) for b in range(B))}
        out['ours_vs_fp32_eager'] = {'max_dlogit': (ours_mask - m32).abs().max().item(),
                                     'min_iou': min(mask_iou(ours_mask[b], m32[b]) for b in range(B))}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    del sd
    torch.cuda.empty_cache()
    return out


def golden_parity(mask, flags):
    """Clips 0 and 7 of the batch this bench times (rank 0) against the UNMODIFIED reference's outputs on the same
    clips (tests/golden/full_bench_b8.npz, made by oracle/make_golden.py).  Tolerance: BASELINE.json north_star."""
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'full_bench_b8.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    ly, lx = meta['lattice']
    g = torch.from_numpy(z['mask'])
    m = mask[meta['samples']].cpu()[:, :, :, ::ly, ::lx]
    f = flags[meta['samples']].cpu()
    res = {'source': 'tests/golden/full_bench_b8.npz (unmodified reference, fp32 CPU)', 'clips': meta['samples'],
           'max_dlogit': (m - g).abs().max().item(), 'min_iou': min(mask_iou(m[i], g[i]) for i in range(g.shape[0])),
           'flags_max_err': (f - torch.from_numpy(z['flags'])).abs().max().item(), 'tol_dlogit': 1e-2, 'tol_iou': 0.99,
           # pixels of the reference whose logit is within 0.01 of the threshold: these may legitimately flip
           'near_threshold_pct': round(100.0 * (g.abs() < 0.01).float().mean().item(), 3)}
    res['ok'] = bool(res['max_dlogit'] <= 1e-2 and res['min_iou'] >= 0.99 and res['flags_max_err'] <= 2e-2)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import synth

    world, rank, local, dev = dist_env()
    peaks = load_peaks()

    net = tcow_b200.Seeker(logging.getLogger('bench'), **SEEKER_KW)
    net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF))
    net = net.to(dev).eval()
    eng = net.seeker.engine()
    if args.chunk > 0:
        eng.max_chunk = args.chunk
    B = args.batch
    # 3 queries per video (README.md:42): clips come in triples sharing the RGB and differing in the query.
    rgb_f, q_f = synth.bench_clips(rank, B, T, HF, WF)
    rgb_h, q_h = rgb_f.pin_memory(), q_f.pin_memory()
    rgb_d, q_d = rgb_h.to(dev), q_h.to(dev)

    with torch.no_grad():
        # ---------------------------------------------------------- device-resident throughput (`value`)
        for _ in range(args.warmup):
            net(rgb_d, q_d)
        barrier(world)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(args.steps):
            mask, flags = net(rgb_d, q_d)           # the product path: CUDA-graph replay of the engine-owned launches
            launches += eng.launches
        e1.record()
        barrier(world)
        ms_total = max_over_ranks(e0.elapsed_time(e1), world, dev)
        # the same K steps again with a CUDA-event pair around every kernel launch (no graph replay): per-kernel durations
        # for `roofline` and `breakdown`; `ms_per_step_profiled` says what the instrumentation costs
        eng.profile = []
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            net(rgb_d, q_d)
        p1.record()
        barrier(world)
        ms_profiled = p0.elapsed_time(p1)
        clocks = sampler.stop() if rank == 0 else None
        prof, eng.profile = eng.profile, None
        # the batch that was just timed, checked against the reference's own outputs (rank 0 holds the fixture's clips)
        parity = golden_parity(mask, flags) if (rank == 0 and B == 8) else None
        ours_mask = mask if rank == 0 else None
        del mask, flags

    # -------------------------------------------------------------- end to end through Seeker.forward with host buffers
    # (1) decoder-style inputs: uint8 frames + uint8 query masks in pinned host memory (what a video decoder / the loader
    #     hold before data/data_plugin.py:174 expands them), scaled by 1/255 in the gather kernel;
    # (2) the reference loader's fp32 tensors (data_plugin.py:199-200), 4x the upload.  Both read back the full logits.
    rgb_u8 = (rgb_f * 255.0).round().to(torch.uint8).pin_memory()
    q_u8 = q_f.to(torch.uint8).pin_memory()
    e2e_ms, e2e_h2d, e2e_d2h, out_last = e2e_measure(net, rgb_u8, q_u8, 1.0 / 255.0, args.steps, args.warmup, world, dev)
    e2e_par = None
    if rank == 0:
        # the uint8 path against the fp32 path on the same (quantised) frames: differs only by the rounding of x*(1/255) vs x/255 before the bf16 cast
        with torch.no_grad():
            m_ref, _ = net((rgb_u8.float() / 255.0).to(dev), q_d)
        e2e_par = {'max_dlogit_u8_vs_f32_inputs': (out_last[0].to(dev) - m_ref).abs().max().item()}
        del m_ref
    f32_ms, f32_h2d, f32_d2h, _ = e2e_measure(net, rgb_h, q_h, 1.0, args.steps, args.warmup, world, dev)

    # -------------------------------------------------------------- training step under the same clock (BASELINE configs[3])
    train = None
    if not args.no_train:
        eng.release()
        torch.cuda.empty_cache()
        train = train_measure(args, world, rank, local, dev, steps=args.train_steps, warmup=3, clocks=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    # ---- per-kernel-class breakdown from the CUDA events recorded inside the timed steps
    agg = {}
    for kind, flops, nbytes, a, b in prof:
        d = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += flops; d[2] += nbytes; d[3] += 1
    gemm_ms = sum(v[0] for k, v in agg.items() if k.startswith('gemm'))
    gemm_flops = sum(v[1] for k, v in agg.items() if k.startswith('gemm'))
    gemm_n = sum(v[3] for k, v in agg.items() if k.startswith('gemm'))
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    breakdown = {k: {'ms_per_step': round(v[0] / args.steps, 4), 'launches_per_step': v[3] // args.steps,
                     'tflops': round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[1] and v[0] else None,
                     'gbs': round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[2] and v[0] else None}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    step_tflops = value / world * FLOP_PER_CLIP / 1e12
    traffic = load_traffic()
    gemm_traffic = (traffic or {}).get('classes', {}).get('gemm_bf16_tn_kernel') if B == 8 else None

    line = {
        'metric': 'seeker_fwd_clips_per_s', 'value': value, 'unit': 'clips/s', 'videos_per_s': value / 3.0,
        'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'ms_per_step_profiled': ms_profiled / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'TCOW Seeker forward, random-init ViT-B/16 divided space-time, T=30 240x320, causal temporal '
                               f'attn, batch {B} clips per GPU (3 queries per video), bf16 GEMM operands / fp32 accumulate '
                               f'and residual stream (BASELINE configs[1])',
                   'batch_per_gpu': B, 'parallelism': f'replicas x{world}, no collective',
                   'l2': 'activation working set ~1.2 GB per step >> 126 MB L2; no flush needed'},
        'roofline': {'bound': 'tensor',
                     # whole step first: algorithmic FLOPs of the reference forward over the step time
                     'step_tflops_algorithmic': round(step_tflops, 1),
                     'step_frac_of_burst_peak': round(step_tflops / peaks['burst'], 4),
                     'step_frac_of_sustained_peak': round(step_tflops / peaks['sustained'], 4),
                     # dominant kernel: 2MNK of every GEMM launch of the step over their CUDA-event durations
                     'achieved': round(achieved, 1), 'peak': peaks['sustained'], 'unit': 'TFLOP/s',
                     'frac': round(achieved / peaks['sustained'], 4),
                     'frac_of_burst_peak': round(achieved / peaks['burst'], 4),
                     'traffic': None if gemm_traffic is None else gemm_traffic['dram_bytes_per_launch'],
                     'traffic_unit': 'bytes/launch',
                     'traffic_source': None if traffic is None else traffic.get('source'),
                     'kernel': 'gemm_bf16_tn_kernel (tcgen05), all launches of the step',
                     'launches_per_step': gemm_n // args.steps, 'gemm_ms_per_step': round(gemm_ms / args.steps, 3),
                     'peak_source': peaks['source'] + ' bf16_tflops_sustained (kernel timed inside a long step)'},
        'parity': parity,
        'breakdown': breakdown,
        'clocks': clocks,
        'e2e': {'value': world * B / (e2e_ms * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': e2e_h2d,
                'd2h_bytes_per_step': e2e_d2h, 'ms_per_step': e2e_ms,
                'inputs': 'uint8 frames + uint8 query masks in pinned host memory (decoder format), x1/255 in the gather kernel',
                'outputs': 'full (B,3,T,Hf,Wf) fp32 logits + flags copied to pinned host memory every step', **(e2e_par or {})},
        'e2e_f32_inputs': {'value': world * B / (f32_ms * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': f32_h2d,
                           'd2h_bytes_per_step': f32_d2h, 'ms_per_step': f32_ms,
                           'inputs': "fp32 frames + fp32 query masks, the reference loader's tensors (data_plugin.py:199-200)"},
        'gpu_launches': launches,
    }
    if train is not None:
        line['train'] = train
    if world == 1 and not args.no_eager_baseline:
        line['gpu_eager_baseline'] = eager_gpu_baseline(dev, rgb_d, q_d, ours_mask, args.steps, args.warmup)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times = cpu_forward_timing(3, cores)[1:]
        line['cpu_baseline'] = {'value': len(times) / sum(times), 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                                'sample': '2 timed forwards of 1 clip (after 1 warm-up) of the same T=30 240x320 workload, '
                                          'fp32 oracle port on the host cores'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity['ok']:
        raise SystemExit(f'bench: the timed batch disagrees with the reference golden vectors: {parity}')


def run_eager(args):
    """--impl eager: the reference algorithm under stock eager PyTorch on the GPU (autocast bf16), same line format."""
    import torch

    from tcow_b200 import synth
    if int(os.environ.get('RANK', '0')) != 0:
        return
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    B = args.batch
    rgb, q = synth.bench_clips(0, B, T, HF, WF)
    rgb, q = rgb.to(dev), q.to(dev)
    r = eager_gpu_baseline(dev, rgb, q, torch.zeros((B, 3, T, HF, WF), device=dev), args.steps, args.warmup)
    v = r['autocast_bf16']['clips_per_s']
    line = {'impl': 'eager', 'metric': 'seeker_fwd_clips_per_s', 'value': v, 'unit': 'clips/s', 'n_gpus': 1,
            'steps': max(3, args.steps // 2), 'warmup': max(2, args.warmup // 2), 'ms_per_step': r['autocast_bf16']['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'TCOW Seeker forward, T=30 240x320, causal, batch {B}: the reference algorithm under stock '
                                   f'eager PyTorch torch.autocast(cuda, bf16) on one B200 (cuBLAS/cuDNN/ATen kernels)'},
            'gpu_eager_baseline': {k: v for k, v in r.items() if k != 'ours_vs_fp32_eager'}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def train_measure(args, world, rank, local, dev, steps, warmup, clocks=True):
    """BASELINE configs[3]: fwd + hand-written bwd + fused AdamW, 2 videos x 3 queries per GPU, data-parallel with the
    bucketed gradient all-reduce of tcow_b200/ddp.py overlapped with the backward (NCCL over NVLink; train.py:86-102,
    222-223).  Returns the JSON record on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import ddp, synth

    peaks = load_peaks()
    kw = dict(SEEKER_KW)
    kw['drop_path_rate'] = args.drop_path
    net = tcow_b200.Seeker(logging.getLogger('bench'), **kw)
    net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF))
    net = net.to(dev).train()
    ddp.broadcast_parameters(net)
    grad_group = ddp.make_gradient_group(args.nccl_ctas) if (world > 1 and args.nccl_ctas > 0) else None
    sync = ddp.attach(net, process_group=grad_group) if world > 1 else None
    if sync is not None:
        sync.timing = True
    eng = net.seeker.train_engine()
    V, Q = args.videos, 3
    B = V * Q
    g = torch.Generator().manual_seed(77 + rank)
    rgb_h = torch.rand(V, 3, T, HF, WF, generator=g).pin_memory()
    q_h = torch.stack([torch.stack([synth.make_clip(1000 + rank * B + v * Q + j, T, HF, WF)[1] for j in range(Q)])
                       for v in range(V)]).pin_memory()
    tm, tf = synth.make_targets(list(range(rank * B, rank * B + B)), T, HF, WF)
    tm, tf = tm.to(dev), tf.to(dev)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)       # args.py:108,179 defaults

    # Inputs start in pinned host memory every step; a copy stream uploads step i+1 while step i computes (what a
    # DataLoader with pin_memory + non_blocking does, train.py's loaders included).
    copy_stream = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream()
    bufs = [(torch.empty((V, 3, T, HF, WF), device=dev), torch.empty((V, Q, 1, T, HF, WF), device=dev)) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for k in range(2):
        ev_free[k].record(cur)
    state = {'i': 0}

    def upload(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])
            bufs[k][0].copy_(rgb_h, non_blocking=True)
            bufs[k][1].copy_(q_h, non_blocking=True)
            ev_ready[k].record(copy_stream)

    upload(0)

    def step():
        i = state['i']
        state['i'] = i + 1
        upload(i + 1)
        k = i & 1
        cur.wait_event(ev_ready[k])
        rgb, q = bufs[k]
        opt.zero_grad(set_to_none=True)
        mask, flags = net.forward_queries(rgb, q)
        loss = synth.training_loss(mask.flatten(0, 1), flags.flatten(0, 1), tm, tf)
        loss.backward()
        ev_free[k].record(cur)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    barrier(world)
    sampler = ClockSampler(local) if (clocks and rank == 0) else None
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    exposed = []
    launches = 0
    for _ in range(steps):
        loss = step()
        launches += eng.launches
        if sync is not None:
            exposed.append(sync.exposed_time_ms())
    e1.record()
    loss_val = float(loss.detach())       # the step's result read back to the host
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clk = sampler.stop() if sampler is not None else None
    # per-kernel-class breakdown of one more (untimed) step
    eng.profile = []
    step()
    torch.cuda.synchronize()
    prof, eng.profile = eng.profile, None
    line = None
    if rank == 0:
        agg = {}
        for kind, flops, nbytes, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += flops; d[2] += nbytes; d[3] += 1
        breakdown = {k: {'ms_per_step': round(v[0], 3), 'launches_per_step': v[3],
                         'tflops': round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[1] and v[0] else None,
                         'gbs': round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[2] and v[0] else None}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        gemm = [v for k, v in agg.items() if k.startswith(('gemm', 'dgrad', 'wgrad'))]
        gemm_ms, gemm_fl = sum(v[0] for v in gemm), sum(v[1] for v in gemm)
        ms_step = ms / steps
        value = world * B * steps / (ms * 1e-3)
        line = {'metric': 'seeker_train_samples_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                'steps': steps, 'warmup': warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': f'TCOW Seeker training step (fwd + hand-written bwd + fused AdamW), T=30 240x320, causal, '
                                       f'{V} videos x {Q} queries per GPU, drop_path_rate={args.drop_path}, data-parallel x{world} '
                                       f'with bucketed NCCL gradient all-reduce overlapped with the backward (BASELINE configs[3])',
                           'batch_per_gpu': B, 'parallelism': f'dp{world}',
                           'l2': 'activation working set >> 126 MB L2; no flush needed'},
                'roofline': {'bound': 'tensor', 'achieved': round(gemm_fl / (gemm_ms * 1e-3) / 1e12, 1) if gemm_ms else None,
                             'peak': peaks['sustained'], 'unit': 'TFLOP/s',
                             'frac': round(gemm_fl / (gemm_ms * 1e-3) / 1e12 / peaks['sustained'], 4) if gemm_ms else None,
                             'traffic': None, 'kernel': 'gemm_bf16_tn_kernel + gemm_bf16_wgrad_kernel (tcgen05), all launches of a step',
                             'step_tflops_algorithmic': round(3 * FLOP_PER_CLIP * B / (ms_step * 1e-3) / 1e12, 1),
                             'step_frac_of_burst_peak': round(3 * FLOP_PER_CLIP * B / (ms_step * 1e-3) / 1e12 / peaks['burst'], 4)},
                'allreduce': {'bytes_per_step': int(eng.last_flat.numel() * 4) if eng.last_flat is not None else None,
                              'exposed_ms_per_step': round(statistics.mean(x for x in exposed if x is not None), 3) if exposed and exposed[0] is not None else 0.0},
                'breakdown': breakdown, 'clocks': clk, 'loss': loss_val,
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': int(rgb_h.numel() * 4 + q_h.numel() * 4),
                        'd2h_bytes_per_step': 4, 'note': 'inputs are uploaded from pinned host memory every step on a copy stream (double-buffered)'},
                'gpu_launches': launches}
    del net, opt, eng, bufs
    torch.cuda.empty_cache()
    return line


def run_train(args):
    """--workload train: the training record alone (it is also the `train` sub-record of the default line)."""
    import torch.distributed as dist
    world, rank, local, dev = dist_env()
    line = train_measure(args, world, rank, local, dev, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sweep_bench(args):
    """--workload sweep (BASELINE configs[2]): eval/test.py-style inference sweep — V synthetic 200-frame videos x 4 queries
    x the temporal strides that fit (6) — sharded by clip over the ranks, no collective on the data path (strong
    scaling: the sweep is fixed, time = max over ranks, host clip assembly + upload + metrics read-back included)."""
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import sweep, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    net = tcow_b200.Seeker(logging.getLogger('bench'), **SEEKER_KW)
    net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF))
    net = net.to(dev).eval()
    V, Q, F = args.videos_total, 4, 200
    items = sweep.plan_sweep(V, Q, F, T, query_idx=15)
    mine = sweep.shard_clips(items, rank, world)
    my_videos = sorted({it.video for it in mine})
    g = torch.Generator().manual_seed(5)
    # decoded uint8 frames and uint8 target masks in pinned host memory, as a video loader would leave them
    base = torch.randint(0, 256, (3, F, HF, WF), generator=g, dtype=torch.uint8)   # one synthetic video, rolled per id
    videos = {v: torch.roll(base, shifts=7 * v, dims=3).pin_memory() for v in my_videos}
    tgt = (torch.rand(3, F, HF // 8, WF // 8, generator=g) > 0.6).to(torch.uint8).repeat_interleave(8, 2) \
        .repeat_interleave(8, 3).pin_memory()

    def get_query(v, q):
        m = torch.zeros(HF, WF)
        m[20 + 30 * q:60 + 30 * q, 40 + 50 * q:100 + 50 * q] = 1
        return m

    fn = lambda its: sweep.run_sweep(net, its, videos.__getitem__, get_query, lambda v, q: tgt, T, dev, clips_per_pass=args.clips_per_pass)
    fn(mine[:8])                                                   # warm-up (weights packed, graphs captured)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = fn(mine)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    merged = sweep.gather_to_rank0({k: v for k, v in res.items()})
    if rank == 0:
        assert len(merged) == len(items)
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        sweep.write_itemized_csv(os.path.join(ROOT, 'gpurun_out', f'itemized_results_{world}gpu.csv'), merged)
        line = {'metric': 'seeker_sweep_samples_per_s', 'value': len(items) / dt, 'unit': 'samples/s', 'n_gpus': world,
                'steps': 1, 'warmup': 1, 'ms_per_step': 1e3 * dt, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': f'eval/test.py-style sweep: {V} videos (200 frames, 240x320) x {Q} queries x 6 strides = '
                                       f'{len(items)} samples, sharded by clip over {world} GPU(s), no collective (BASELINE configs[2]); '
                                       f'host clip assembly, upload and IoU read-back inside the timed region',
                           'parallelism': f'clip-sharded x{world}'},
                'e2e': {'value': len(items) / dt, 'unit': 'samples/s',
                        # every video (and its target masks) is uploaded once; clips are gathered on the device
                        'h2d_bytes_per_step': int(V * (1 + Q) * 3 * F * HF * WF + len(items) * HF * WF * 4),
                        'd2h_bytes_per_step': int(len(items) * (3 * T * 3 + 3) * 4)},
                'gpu_launches': net.seeker.engine().launches * (len(mine) // (2 * Q))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_hires(args):
    """--workload hires (BASELINE configs[4]): T=60, 480x640 (1200 patches per frame), non-causal temporal attention (cls
    mean path), bf16, one clip per pass and GPU; N > 1: independent replicas, no collective."""
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import synth

    world, rank, local, dev = dist_env()
    peaks = load_peaks()
    Th, Hh, Wh = 60, 480, 640
    kw = dict(SEEKER_KW, num_total_frames=Th, num_visible_frames=Th, frame_height=Hh, frame_width=Wh, causal_attention=0)
    net = tcow_b200.Seeker(logging.getLogger('bench'), **kw)
    net.load_state_dict(synth.make_state_dict(901, num_frames=Th, frame_height=Hh, frame_width=Wh))
    net = net.to(dev).eval()
    rgb_f, q_f = synth.make_batch([20 + rank], num_frames=Th, frame_height=Hh, frame_width=Wh)
    flop = 20878.31e9          # equals FlopCounterMode on the reference
    eng = net.seeker.engine()

    with torch.no_grad():
        rgb, q = rgb_f.to(dev), q_f.to(dev)
        for _ in range(args.warmup):
            net(rgb, q)
        barrier(world)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(args.steps):
            mask, flags = net(rgb, q)
            launches += eng.launches
        e1.record()
        barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
        clocks = sampler.stop() if rank == 0 else None
        parity = None
        if rank == 0:
            # the clip that was just timed is the one the unmodified reference ran for tests/golden/hires_causal0.npz
            import numpy as np
            z = np.load(os.path.join(ROOT, 'tests', 'golden', 'hires_causal0.npz'))
            meta = json.loads(bytes(z['meta']).decode())
            ly, lx = meta['lattice']
            g = torch.from_numpy(z['mask'])
            m = mask.cpu()[:, :, :, ::ly, ::lx]
            parity = {'source': 'tests/golden/hires_causal0.npz (unmodified reference, fp32 CPU)',
                      'max_dlogit': (m - g).abs().max().item(), 'min_iou': mask_iou(m, g),
                      'flags_max_err': (flags.cpu() - torch.from_numpy(z['flags'])).abs().max().item()}
        del mask, flags
    # end to end: uint8 clip in pinned host memory every step, full logits read back (see e2e_measure)
    rgb_u8 = (rgb_f * 255.0).round().to(torch.uint8).pin_memory()
    q_u8 = q_f.to(torch.uint8).pin_memory()
    e2e_ms, h2d, d2h, _ = e2e_measure(net, rgb_u8, q_u8, 1.0 / 255.0, args.steps, args.warmup, world, dev)
    if rank == 0:
        value = world * args.steps / (ms * 1e-3)
        tfl = value / world * flop / 1e12
        print(json.dumps({
            'metric': 'seeker_fwd_clips_per_s_hires', 'value': value, 'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'TCOW Seeker forward, T=60 480x640 (S=1201 tokens per frame, 72 000 per clip), causal_attention=0, '
                                   'batch 1 clip per GPU (BASELINE configs[4])', 'batch_per_gpu': 1,
                       'parallelism': f'replicas x{world}, no collective', 'l2': 'activation working set ~1.2 GB >> 126 MB L2'},
            'roofline': {'bound': 'tensor', 'achieved': round(tfl, 1), 'peak': peaks['burst'], 'unit': 'TFLOP/s',
                         'frac': round(tfl / peaks['burst'], 4), 'traffic': None,
                         'kernel': 'whole step, algorithmic FLOPs of the reference forward (20 878 GFLOP per clip)'},
            'parity': parity, 'clocks': clocks,
            'e2e': {'value': world / (e2e_ms * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms},
            'gpu_launches': launches}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'eager'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true', help='skip the stock-PyTorch-on-GPU baseline leg')
    ap.add_argument('--no-train', action='store_true', help='skip the training sub-record (BASELINE configs[3]) of the default line')
    ap.add_argument('--train-steps', type=int, default=10, help='timed training steps of the `train` sub-record')
    ap.add_argument('--chunk', type=int, default=0, help='clips per engine pass (0 = engine default)')
    ap.add_argument('--workload', default='infer', choices=['infer', 'train', 'sweep', 'hires'],
                    help='infer = BASELINE configs[1] (the headline metric, default); train = configs[3] (fwd+bwd+AdamW, DDP); '
                         'sweep = configs[2] (clip-sharded evaluation sweep); hires = configs[4] (T=60, 480x640)')
    ap.add_argument('--clips-per-pass', type=int, default=2, help='sweep: clips (x4 queries) per forward_queries call')
    ap.add_argument('--videos-total', type=int, default=8, help='sweep: videos in the whole sweep (fixed as GPUs grow)')
    ap.add_argument('--videos', type=int, default=2, help='train: videos per GPU (x3 queries each)')
    ap.add_argument('--drop-path', type=float, default=0.1, help='train: stochastic-depth rate (args.py default 0.1)')
    ap.add_argument('--nccl-ctas', type=int, default=0,
                    help='train: cap on the SMs the gradient all-reduce may occupy (0 = NCCL default, which measured faster)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 1)
    if args.impl == 'reference':
        run_reference(args)
    elif args.impl == 'eager':
        run_eager(args)
    elif args.workload == 'train':
        run_train(args)
    elif args.workload == 'sweep':
        run_sweep_bench(args)
    elif args.workload == 'hires':
        run_hires(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
