"""Data-parallel training of the Seeker (BASELINE configs[3]; the reference wraps its networks in nn.DataParallel,
train.py:223, and lets autograd + a single-process gather do the reduction).

One process per GPU; every rank runs the same hand-written backward (train_engine.py).  Its gradients live in one flat
fp32 buffer whose ranges complete in a known order (head, block 11, ..., block 0, embeddings), so the exchange is a
bucketed all-reduce overlapped with the rest of the backward: as soon as a range is final, a side stream waits on an
event of the compute stream and all-reduces it (NCCL over NVLink/NVSwitch on B200; gloo on CPU in the tests) while the
compute stream differentiates the next block.  The packed->reference gradient mapping is linear, so it runs after the
reduction, once.  No collective sits on the forward or inference path."""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None, average=True, bucket_bytes=25 << 20):
        self.pg = process_group
        self.average = average
        self.bucket_floats = max(1, bucket_bytes // 4)
        self.flat = None
        self.handles = []
        self.stream = None
        self.ranges = []          # (lo, hi) all-reduced so far — the tests check coverage and order
        self.exposed_ms = None    # time the compute stream waited in finish() (measured when timing=True)
        self.timing = False
        self._pending = None

    def world_size(self):
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def begin(self, flat):
        self.flat = flat
        self.handles, self.ranges, self._pending = [], [], None
        if flat.is_cuda and self.stream is None:
            self.stream = torch.cuda.Stream(device=flat.device)

    def _reduce(self, lo, hi):
        chunk = self.flat[lo:hi]
        ws = self.world_size()
        if ws == 1:
            return
        backend = dist.get_backend(self.pg)
        op = dist.ReduceOp.AVG if (self.average and backend == 'nccl') else dist.ReduceOp.SUM
        if self.flat.is_cuda:
            ev = torch.cuda.Event()
            ev.record()                                   # everything that produced [lo, hi) is before this point
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                h = dist.all_reduce(chunk, op=op, group=self.pg, async_op=True)
                if self.average and op == dist.ReduceOp.SUM:
                    h.wait()
                    chunk.div_(ws)
        else:
            h = dist.all_reduce(chunk, op=op, group=self.pg, async_op=True)
            if self.average and op == dist.ReduceOp.SUM:
                h.wait()
                chunk.div_(ws)
        self.handles.append(h)

    def ready(self, lo, hi):
        """[lo, hi) of the flat buffer is final.  Small neighbouring ranges are merged up to the bucket size."""
        if self._pending is not None and self._pending[1] == lo:
            lo = self._pending[0]
        elif self._pending is not None:
            self._flush()
        self._pending = (lo, hi)
        if hi - lo >= self.bucket_floats:
            self._flush()

    def _flush(self):
        if self._pending is None:
            return
        lo, hi = self._pending
        self._pending = None
        self.ranges.append((lo, hi))
        self._reduce(lo, hi)

    def finish(self):
        self._flush()
        if self.flat is not None and self.flat.is_cuda:
            cur = torch.cuda.current_stream(self.flat.device)
            if self.timing:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
            for h in self.handles:
                h.wait()
            cur.wait_stream(self.stream)
            if self.timing:
                e1.record(cur)
                self._timing_events = (e0, e1)
        else:
            for h in self.handles:
                h.wait()
        self.handles = []

    def exposed_time_ms(self):
        if getattr(self, '_timing_events', None) is None:
            return None
        e0, e1 = self._timing_events
        e1.synchronize()
        return e0.elapsed_time(e1)


def make_gradient_group(max_ctas=4):
    """A dedicated NCCL communicator for the gradient exchange that occupies at most `max_ctas` SMs.  The step needs about
    8 GB/s of all-reduce bandwidth (458 MB per ~60 ms), a small fraction of NVLink; NCCL's default channel count would take
    dozens of SMs away from the persistent one-CTA-per-SM GEMM kernels the exchange overlaps with (their tiles on the
    occupied SMs then finish a full tile late).  Falls back to the default group when the option is unavailable.
    MEASURED (8 x B200, bench.py --workload train): max_ctas=4 is SLOWER than NCCL's default (676 vs 728 samples/s: the
    throttled all-reduce no longer hides behind the backward, 1.7 ms exposed) — kept as an option, not the default."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != 'nccl':
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(max_ctas)
        opts.config.min_ctas = 1
        return dist.new_group(backend='nccl', pg_options=opts)
    except (AttributeError, TypeError, RuntimeError):
        return None


def attach(seeker_module, process_group=None, average=True, bucket_bytes=25 << 20):
    """Make `seeker_module` (tcow_b200.Seeker or QueryMaskTracker) average its gradients over the process group
    inside its own backward.  Parameters must already be identical on every rank (broadcast_parameters)."""
    tracker = getattr(seeker_module, 'seeker', seeker_module)
    sync = GradSync(process_group, average, bucket_bytes)
    tracker.train_engine().grad_sync = sync
    return sync


def broadcast_parameters(module, src=0, process_group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return
    # p.detach() shares the parameter's version counter (p.data does not): the in-place write of the broadcast bumps
    # p._version, which is what SeekerEngine's packed-weight / CUDA-graph cache is stamped with (engine.py:_stamp)
    with torch.no_grad():
        for p in module.parameters():
            dist.broadcast(p.detach(), src=src, group=process_group)
        for b in module.buffers():
            dist.broadcast(b.detach(), src=src, group=process_group)
    tracker = getattr(module, 'seeker', module)
    if getattr(tracker, '_engine', None) is not None:
        tracker._engine.invalidate()
