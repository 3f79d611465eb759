"""Sharding of the inference sweep (config 3) incl. a world_size-2 gloo run on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import pytest

from tcow_b200 import sweep


def test_clip_strides_match_reference_rule():
    # demo/teaduck2.mp4: 200 frames, query at frame 15, T=30, query_time=0 -> strides 1..6
    assert sweep.clip_strides(200, 30, 15) == [(15, s) for s in range(1, 7)]
    assert sweep.clip_strides(30, 30, 0) == [(0, 1)]
    assert sweep.clip_strides(29, 30, 0) == []
    # seeker_query_time > 0 shifts the clip start before the query frame (data_utils.py:325)
    assert sweep.clip_strides(200, 30, 20, query_time=2)[:2] == [(18, 1), (16, 2)]


def test_shard_partitions_everything_once():
    items = sweep.plan_sweep(num_videos=3, num_queries=4, num_video_frames=200, num_frames=30, query_idx=15)
    assert len(items) == 3 * 6 * 4
    for world in (1, 2, 4, 8):
        parts = [sweep.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, []), key=items.index) == items
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert [len(b) for b in sweep.batches(items[:10], 4)] == [4, 4, 2]


def test_shard_clips_keeps_queries_together():
    items = sweep.plan_sweep(num_videos=3, num_queries=4, num_video_frames=200, num_frames=30, query_idx=15)
    for world in (1, 2, 4, 8):
        parts = [sweep.shard_clips(items, r, world) for r in range(world)]
        assert sorted(sum(parts, []), key=items.index) == items
        for part in parts:                       # every clip of a part carries all 4 of its queries
            for _, its in sweep.group_by_clip(part):
                assert [i.query for i in its] == [0, 1, 2, 3]


def test_itemized_csv_round_trip(tmp_path):
    rows = {(0, 1, 15, 2): dict(video=0, query=1, frame_start=15, frame_stride=2, mean_snitch_iou=0.5, count_snitch_iou=30),
            (0, 0, 15, 1): dict(video=0, query=0, frame_start=15, frame_stride=1, mean_snitch_iou=0.25, count_snitch_iou=7)}
    p = tmp_path / 'itemized_results.csv'
    sweep.write_itemized_csv(str(p), rows)
    import pandas as pd
    df = pd.read_csv(p)
    assert list(df.columns) == sweep.CSV_COLUMNS and len(df) == 2
    assert df.iloc[0]['query'] == 0 and abs(df.iloc[1]['mean_snitch_iou'] - 0.5) < 1e-9


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    items = sweep.plan_sweep(2, 4, 200, 30, 15)
    mine = sweep.shard(items, rank, world)
    local = {(it.video, it.query, it.frame_stride): float(it.video * 100 + it.query * 10 + it.frame_stride)
             for it in mine}
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    merged = sweep.gather_to_rank0(local)
    if rank == 0:
        q.put((len(merged), float(t), sorted(merged) == sorted((i.video, i.query, i.frame_stride) for i in items)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sweep():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    n, tmax, complete = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert n == 48 and tmax == 2.0 and complete


def test_run_sweep_host_logic_with_a_stub_network(monkeypatch):
    """Clip grouping, ragged query groups (padded by repeating the last query), uint8 expansion and the per-item rows of
    run_sweep, with the network and the device kernel replaced by torch stubs (the real ones are covered on the GPU)."""
    import torch

    from tcow_b200 import ops
    T, Hf, Wf, F = 3, 8, 12, 10

    class StubNet:
        calls = []

        def forward_queries(self, rgb, q):
            StubNet.calls.append((tuple(rgb.shape), tuple(q.shape)))
            B, Qs = q.shape[:2]
            # logits > 0 exactly where the query rectangle of frame 0 is, in every frame and channel
            m = (q[:, :, :, 0:1] > 0).float().expand(B, Qs, 3, T, Hf, Wf) * 2 - 1
            flags = rgb.mean(dim=(1, 3, 4))[:, None, :, None].expand(B, Qs, T, 3).contiguous()
            return m.contiguous(), flags

    def areas(logits, target):
        p, g = logits > 0, target > 0.5
        return torch.stack([g.sum((-1, -2)), (p & g).sum((-1, -2)), (p | g).sum((-1, -2))], -1).float()

    monkeypatch.setattr(ops, 'mask_iou_areas', areas)
    items = sweep.plan_sweep(num_videos=2, num_queries=3, num_video_frames=F, num_frames=T, query_idx=1)
    items = [it for it in items if not (it.video == 1 and it.frame_stride == 2 and it.query == 2)]   # one ragged group
    vids = {v: torch.full((3, F, Hf, Wf), 51 * (v + 1), dtype=torch.uint8) for v in range(2)}      # uint8: /255 on "device"

    def get_query(v, q):
        m = torch.zeros(Hf, Wf)
        m[q:q + 2, 0:4] = 1
        return m

    tgt = torch.zeros(3, F, Hf, Wf)
    tgt[0, :, 0:2, 0:4] = 1           # channel 0: identical to query 0's rectangle -> IoU 1 for q = 0
    res = sweep.run_sweep(StubNet(), items, vids.__getitem__, get_query, lambda v, q: tgt, T, torch.device('cpu'),
                          clips_per_pass=2)
    assert sorted(res) == sorted((i.video, i.query, i.frame_start, i.frame_stride) for i in items)
    n_groups = len(sweep.group_by_clip(items))
    assert len(StubNet.calls) == (n_groups + 1) // 2 and all(c[1][1] == 3 for c in StubNet.calls)   # padded to 3 queries
    r0 = res[(0, 0, 1, 1)]
    assert r0['mean_snitch_iou'] == pytest.approx(1.0) and r0['count_snitch_iou'] == T
    assert r0['count_occl_mask_iou'] == 0 and r0['mean_occl_mask_iou'] == -1.0                       # empty target channel
    assert res[(0, 2, 1, 1)]['mean_snitch_iou'] == pytest.approx(0.0)                                 # disjoint rectangle
    assert res[(1, 0, 1, 3)]['flag_occl_mean'] == pytest.approx(102 / 255, abs=1e-6)                 # uint8 frames expanded
