#!/usr/bin/env python
"""Benchmark of the one hot path this repo accelerates: TCOW Seeker forward, clips/s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

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


def run_ours(args):
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()

    net = tcow_b200.Seeker(logging.getLogger('bench'), **SEEKER_KW)
    net.load_state_dict(synth.make_state_dict(901, num_frames=T, frame_height=HF, frame_width=WF))
    net = net.to(dev).eval()
    eng = net.seeker.engine()
    if args.chunk > 0:
        eng.max_chunk = args.chunk
    B = args.batch
    # 3 queries per video (README.md:42): clips come in triples sharing the RGB and differing in the query.
    rgb_l, q_l = [], []
    for i in range(B):
        vid = (rank * B + i) // 3
        r, _ = synth.make_clip(vid, T, HF, WF)
        _, q = synth.make_clip(1000 + rank * B + i, T, HF, WF)
        rgb_l.append(r); q_l.append(q)
    rgb_h = torch.stack(rgb_l).pin_memory()
    q_h = torch.stack(q_l).pin_memory()
    rgb_d, q_d = rgb_h.to(dev), q_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ---------------------------------------------------------- device-resident throughput (`value`)
        for _ in range(args.warmup):
            net(rgb_d, q_d)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        eng.profile = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(args.steps):
            net(rgb_d, q_d)
            launches += eng.launches
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        prof, eng.profile = eng.profile, None

        # ---------------------------------------------------------- end to end through Seeker.forward with host buffers
        res_h = torch.empty((B, T, 3), dtype=torch.float32).pin_memory()
        area_h = torch.empty((B, 3, T), dtype=torch.float32).pin_memory()
        # Inputs start in pinned host memory every step; a copy stream uploads step i+1 while step i computes
        # (what a DataLoader with pin_memory + non_blocking does); the step's result is read back to the host.
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        bufs = [(torch.empty_like(rgb_d), torch.empty_like(q_d)) for _ in range(2)]
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            k = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[k])
                bufs[k][0].copy_(rgb_h, non_blocking=True)
                bufs[k][1].copy_(q_h, non_blocking=True)
                ev_ready[k].record(copy_stream)

        def e2e_run(n):
            for k in range(2):
                ev_free[k].record(cur)
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                k = i & 1
                cur.wait_event(ev_ready[k])
                mask, flags = net(bufs[k][0], bufs[k][1])
                ev_free[k].record(cur)
                res_h.copy_(flags, non_blocking=True)
                area_h.copy_((mask > 0).float().mean(dim=(3, 4)), non_blocking=True)   # per-frame mask area read back
                cur.synchronize()

        e2e_run(min(args.warmup, 3))
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))

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

    line = {
        'metric': 'seeker_fwd_clips_per_s', 'value': value, 'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'TCOW Seeker forward, random-init ViT-B/16 divided space-time, T=30 240x320, causal temporal '
                               f'attn, batch {B} clips per GPU (3 queries per video), bf16 GEMM operands / fp32 accumulate '
                               f'and residual stream (BASELINE configs[1])',
                   'batch_per_gpu': B, 'parallelism': f'replicas x{world}, no collective',
                   'l2': 'activation working set ~1.2 GB per step >> 126 MB L2; no flush needed'},
        'roofline': {'bound': 'tensor', 'achieved': round(achieved, 1), 'peak': peaks['sustained'], 'unit': 'TFLOP/s',
                     'frac': round(achieved / peaks['sustained'], 4),
                     # dram__bytes_read+write per launch, launch-weighted mean over the step's GEMM classes, from the
                     # `ncu --set full` capture summarised in profiles/r01e_ncu_full.txt (B=8; bytes: qkv 392e6 x24,
                     # fc1 502e6 x12, fc2 860e6 x12, proj 497e6 x24); algorithmic mean 583e6 — no operand re-reads
                     'traffic': 523.3e6 if B == 8 else None, 'traffic_unit': 'bytes/launch',
                     'traffic_source': 'profiles/r01e_ncu_full.txt',
                     'kernel': 'gemm_bf16_tn_kernel (tcgen05), all launches of the step',
                     'launches_per_step': gemm_n // args.steps, 'gemm_ms_per_step': round(gemm_ms / args.steps, 3),
                     'peak_source': peaks['source'] + ' bf16_tflops_sustained (kernel timed inside a long step)',
                     'step_tflops_algorithmic': round(step_tflops, 1),
                     'step_frac_of_burst_peak': round(step_tflops / peaks['burst'], 4),
                     'step_frac_of_sustained_peak': round(step_tflops / peaks['sustained'], 4)},
        'breakdown': breakdown,
        'clocks': clocks,
        'e2e': {'value': world * B * args.steps / (e2e_ms * 1e-3), 'unit': 'clips/s',
                'h2d_bytes_per_step': int(rgb_h.numel() * 4 + q_h.numel() * 4),
                'd2h_bytes_per_step': int(res_h.numel() * 4 + area_h.numel() * 4), 'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': launches,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times = cpu_forward_timing(3, cores)[1:]
        line['cpu_baseline'] = {'value': len(times) / sum(times), 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                                'sample': '2 timed forwards of 1 clip (after 1 warm-up) of the same T=30 240x320 workload, '
                                          'fp32 oracle port on the host cores'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """--workload train (BASELINE configs[3]): fwd + bwd + AdamW step, 2 videos x 3 queries per GPU, data-parallel with
    the bucketed gradient all-reduce of tcow_b200/ddp.py overlapped with the backward (NCCL over NVLink)."""
    import torch
    import torch.distributed as dist

    import tcow_b200
    from tcow_b200 import ddp, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    exposed = []
    launches = 0
    for _ in range(args.steps):
        loss = step()
        launches += eng.launches
        if sync is not None:
            exposed.append(sync.exposed_time_ms())
    e1.record()
    loss_val = float(loss.detach())       # the step's result read back to the host
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel-class breakdown of one more (untimed) step
    eng.profile = []
    step()
    torch.cuda.synchronize()
    prof, eng.profile = eng.profile, None
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
        ms_step = ms / args.steps
        value = world * B * args.steps / (ms * 1e-3)
        line = {'metric': 'seeker_train_samples_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
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
                'breakdown': breakdown, 'clocks': clocks, 'loss': loss_val,
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': int(rgb_h.numel() * 4 + q_h.numel() * 4),
                        'd2h_bytes_per_step': 4, 'note': 'inputs are uploaded from pinned host memory every step on a copy stream (double-buffered)'},
                'gpu_launches': launches}
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

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()
    Th, Hh, Wh = 60, 480, 640
    kw = dict(SEEKER_KW, num_total_frames=Th, num_visible_frames=Th, frame_height=Hh, frame_width=Wh, causal_attention=0)
    net = tcow_b200.Seeker(logging.getLogger('bench'), **kw)
    net.load_state_dict(synth.make_state_dict(901, num_frames=Th, frame_height=Hh, frame_width=Wh))
    net = net.to(dev).eval()
    rgb_h, q_h = synth.make_batch([rank], num_frames=Th, frame_height=Hh, frame_width=Wh)
    rgb_h, q_h = rgb_h.pin_memory(), q_h.pin_memory()
    flop = 20878.31e9          # equals FlopCounterMode on the reference
    eng = net.seeker.engine()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        rgb, q = rgb_h.to(dev), q_h.to(dev)
        for _ in range(args.warmup):
            net(rgb, q)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = 0
        for _ in range(args.steps):
            net(rgb, q)
            launches += eng.launches
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        # end to end: the clip starts in pinned host memory every step (copy stream, double-buffered: clip i+1 uploads
        # while clip i computes); flags + per-frame mask areas are read back every step
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        bufs = [(torch.empty_like(rgb), torch.empty_like(q)) for _ in range(2)]
        ev_ready = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        res_f = torch.empty((1, Th, 3), dtype=torch.float32).pin_memory()
        res_a = torch.empty((1, 3, Th), dtype=torch.float32).pin_memory()

        def upload(i):
            k = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[k])
                bufs[k][0].copy_(rgb_h, non_blocking=True)
                bufs[k][1].copy_(q_h, non_blocking=True)
                ev_ready[k].record(copy_stream)

        for k in range(2):
            ev_free[k].record(cur)
        upload(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if i + 1 < args.steps:
                upload(i + 1)
            k = i & 1
            cur.wait_event(ev_ready[k])
            mask, flags = net(bufs[k][0], bufs[k][1])
            ev_free[k].record(cur)
            res_f.copy_(flags, non_blocking=True)
            res_a.copy_((mask > 0).float().mean(dim=(3, 4)), non_blocking=True)
            cur.synchronize()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
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
            'clocks': clocks,
            'e2e': {'value': world * args.steps / (e2e_ms * 1e-3), 'unit': 'clips/s',
                    'h2d_bytes_per_step': int(rgb_h.numel() * 4 + q_h.numel() * 4),
                    'd2h_bytes_per_step': int((3 * Th + Th * 3) * 4)},
            'gpu_launches': launches}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
