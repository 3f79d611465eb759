"""Sharding of the inference sweep over GPUs — independent (video, query, stride) samples, one process per
GPU, NO data-path collective.  The sweep shape follows the reference's evaluation driver:
eval/test.py:23-60 loops over the usage modes of a video (data/data_utils.py:301-342, strides 1..10 that
fit the video) and pipeline.py:134-158 loops over the queries of a clip.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


@dataclass(frozen=True)
class SweepItem:
    video: int
    query: int
    frame_start: int
    frame_stride: int


def clip_strides(num_video_frames: int, num_frames: int, query_idx: int, query_time: int = 0,
                 max_stride: int = 10) -> List[tuple]:
    """(frame_start, frame_stride) pairs of a video that fit a fixed-length clip; mirrors the frame-range test
    of data/data_utils.py:325-331 (annotation-coverage filtering is dataset-specific and not modelled)."""
    modes = []
    for s in range(1, max_stride + 1):
        first = query_idx - query_time * s
        last = first + (num_frames - 1) * s
        if first < 0 or last > num_video_frames - 1:
            continue
        modes.append((first, s))
    return modes


def plan_sweep(num_videos: int, num_queries: int, num_video_frames: int, num_frames: int, query_idx: int = 0,
               query_time: int = 0) -> List[SweepItem]:
    items = []
    for v in range(num_videos):
        for (start, stride) in clip_strides(num_video_frames, num_frames, query_idx, query_time):
            for q in range(num_queries):
                items.append(SweepItem(v, q, start, stride))
    return items


def shard(items: Sequence, rank: int, world: int) -> list:
    """Static round-robin shard; every item lands on exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of size {world}')
    return list(items[rank::world])


def batches(items: Sequence, batch: int):
    for i in range(0, len(items), batch):
        yield list(items[i:i + batch])


def gather_to_rank0(local_results: dict, group=None):
    """Control-plane only (after the timed region): collect per-rank result dicts on rank 0."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local_results)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(dict(local_results), out, dst=0, group=group)
    if out is None:
        return None
    merged = {}
    for part in out:
        dup = set(merged) & set(part)
        if dup:
            raise RuntimeError(f'sweep items computed twice: {sorted(dup)[:3]}')
        merged.update(part)
    return merged
