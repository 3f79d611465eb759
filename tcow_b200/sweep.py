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
    query_time: int = 0          # clip frame that carries the query mask (--seeker_query_time, data/data_utils.py:431)


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
                items.append(SweepItem(v, q, start, stride, query_time))
    return items


def shard(items: Sequence, rank: int, world: int) -> list:
    """Static round-robin shard; every item lands on exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of size {world}')
    return list(items[rank::world])


def shard_clips(items: Sequence[SweepItem], rank: int, world: int) -> list:
    """Round-robin over CLIPS (video, start, stride) rather than items, so that all queries of a clip stay on one
    rank and share its frames (one forward_queries call)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of size {world}')
    out = []
    for _, its in group_by_clip(items)[rank::world]:
        out.extend(its)
    return out


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


# ------------------------------------------------------------------------------------------------ sweep runner
CSV_COLUMNS = ['video', 'query', 'frame_start', 'frame_stride', 'mean_snitch_iou', 'mean_occl_mask_iou',
               'mean_cont_mask_iou', 'count_snitch_iou', 'count_occl_mask_iou', 'count_cont_mask_iou', 'flag_occl_mean',
               'flag_cont_mean']


def group_by_clip(items: Sequence[SweepItem]):
    """Items of one (video, start, stride) clip share the RGB frames: they become ONE forward_queries call
    (pipeline.py:134-182 loops over them one forward at a time)."""
    groups = {}
    for it in items:
        groups.setdefault((it.video, it.frame_start, it.frame_stride), []).append(it)
    return [(k, sorted(v, key=lambda i: i.query)) for k, v in sorted(groups.items())]


def run_sweep(net, items: Sequence[SweepItem], get_video, get_query, get_target, num_frames, device, clips_per_pass=2):
    """Run this rank's share of the evaluation sweep (eval/test.py:23-60 + pipeline.py:134-182) on `device`.

    get_video(v) -> (3, F, H, W) host tensor; get_query(v, q) -> (H, W) binary mask at the query frame;
    get_target(v, q) -> (3, F, H, W) target masks (negative = frame not annotated, data/data_plugin.py:187: such frames
    are skipped like eval/metrics.py:21 marks them) or None.  Returns {(video, query, start, stride): row dict} with the
    per-item IoU means / counts of eval/metrics.py:43-82 (areas computed on the device by tcow_mask_iou_areas) and
    the mean flag logits."""
    import torch

    from . import ops
    results = {}
    groups = group_by_clip(items)
    pending = []
    vcache, tcache, qcache = {}, {}, {}   # the current video (its targets, its query masks) live on the device; clips are gathered there

    def video_on_device(v):
        if v not in vcache:
            vcache.clear()
            tcache.clear()
            qcache.clear()
            vid = get_video(v).to(device, non_blocking=True)
            # decoded frames arrive as uint8 (data/data_plugin.py:174 divides by 255 on the host): expand on the device
            vcache[v] = vid.float().div_(255.0) if vid.dtype == torch.uint8 else vid.float()
        return vcache[v]

    def target_on_device(v, q):
        if (v, q) not in tcache:
            t = get_target(v, q)
            tcache[(v, q)] = None if t is None else t.to(device, non_blocking=True).float()
        return tcache[(v, q)]

    def query_on_device(v, q):
        # once per (video, query), not once per clip: a copy from pageable host memory waits for the stream to drain, and one
        # per sample would serialise the host's preparation of a pass with the device's execution of the previous one
        if (v, q) not in qcache:
            qcache[(v, q)] = get_query(v, q).to(device, non_blocking=True)
        return qcache[(v, q)]

    with torch.no_grad():
        for g0 in range(0, len(groups), clips_per_pass):
            part = groups[g0:g0 + clips_per_pass]
            nq = max(len(its) for _, its in part)
            rgb, qm, tg = [], [], []
            have_target = False
            for (v, start, stride), its in part:
                vid = video_on_device(v)
                idx = torch.arange(num_frames, device=device) * stride + start
                rgb.append(vid.index_select(1, idx))
                qs, ts = [], []
                for j in range(nq):
                    it = its[min(j, len(its) - 1)]            # pad ragged groups by repeating the last query
                    q = torch.zeros(1, num_frames, *vid.shape[-2:], device=device)
                    q[0, it.query_time] = query_on_device(v, it.query)
                    qs.append(q)
                    t = target_on_device(v, it.query)
                    # a (clip, query) without ground truth gets an all-negative target: every frame is then ignored
                    ts.append(torch.full((3, num_frames, *vid.shape[-2:]), -1.0, device=device) if t is None
                              else t.index_select(1, idx))
                    have_target = have_target or t is not None
                qm.append(torch.stack(qs))
                tg.append(torch.stack(ts))
            rgb_d = torch.stack(rgb)
            qm_d = torch.stack(qm)
            mask, flags = net.forward_queries(rgb_d, qm_d)                       # (Bv, Q, 3, T, H, W), (Bv, Q, T, 3)
            areas = None
            if have_target:
                tgt = torch.stack(tg).contiguous()
                areas = ops.mask_iou_areas(mask.contiguous(), tgt)
                # frames without annotation (any negative target pixel, eval/metrics.py:21) count as "target absent"
                ignore = (tgt < 0).flatten(-2).any(-1)
                areas[..., 0] = torch.where(ignore, torch.zeros_like(areas[..., 0]), areas[..., 0])
            pending.append((part, areas, None if flags is None else flags.mean(2)))
        # one device->host transfer of all per-item numbers at the end: no per-pass synchronisation
        for part, areas, flags_c in pending:
            areas = None if areas is None else areas.cpu()
            flags_c = None if flags_c is None else flags_c.cpu()
            for bi, ((v, start, stride), its) in enumerate(part):
                for j, it in enumerate(its):
                    row = dict(video=v, query=it.query, frame_start=start, frame_stride=stride)
                    if areas is not None:
                        a = areas[bi, j]                                         # (3, T, 3): gt, inter, union
                        for c, name in enumerate(['snitch_iou', 'occl_mask_iou', 'cont_mask_iou']):
                            valid = a[c, :, 0] > 0                               # frames where the target is present
                            iou = a[c, :, 1] / (a[c, :, 2] + 1e-7)
                            row['count_' + name] = int(valid.sum())
                            row['mean_' + name] = float(iou[valid].mean()) if valid.any() else -1.0
                    if flags_c is not None and flags_c.shape[-1] >= 2:
                        row['flag_occl_mean'], row['flag_cont_mean'] = float(flags_c[bi, j, 0]), float(flags_c[bi, j, 1])
                    results[(v, it.query, start, stride)] = row
    return results


def write_itemized_csv(path, merged):
    """rank 0: the per-item table the reference's evaluation leaves behind as itemized_results.csv."""
    import csv
    with open(path, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=CSV_COLUMNS, extrasaction='ignore')
        w.writeheader()
        for key in sorted(merged):
            w.writerow(merged[key])
