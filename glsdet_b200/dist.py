"""Image-sharded multi-GPU execution: one process per GPU, no data-path collective until the end.

Replaces mmdet's collect_results_gpu (yolox-ufp/mmdet/apis/test.py:161-191: pickle -> uint8 tensor ->
all_gather(sizes) -> all_gather(padded payload)) with a fixed-layout gather: counts [B] int32 and detection rows
[B, max_rows, 7] fp32 - two collectives, no pickling, no host round trip.  NCCL over NVLink on GPUs, gloo on CPU
tensors (tests).  Rank r owns images r, r + world, r + 2*world, ... (DistributedSampler's round-robin order,
test.py:186-190 re-interleaves the same way).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_indices(num_images: int, rank: int, world: int) -> List[int]:
    return list(range(rank, num_images, world))


def gather_detections(det: torch.Tensor, count: torch.Tensor, max_rows: Optional[int] = None, dst: Optional[int] = None,
                      group=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """det [B, R, 7] fp32, count [B] int32 (same B and R on every rank).  Returns (det_all [world*B, max_rows, 7],
    count_all [world*B]) in rank-major order on every rank (dst=None) or only on rank `dst` (others get None)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        r = det.shape[1] if max_rows is None else min(max_rows, det.shape[1])
        return det[:, :r], count.clamp(max=r)
    world = dist.get_world_size(group)
    r = det.shape[1] if max_rows is None else min(max_rows, det.shape[1])
    send_det = det[:, :r].contiguous()
    send_cnt = count.clamp(max=r).contiguous()
    b = det.shape[0]
    det_all = torch.empty((world * b, r, 7), dtype=det.dtype, device=det.device)
    cnt_all = torch.empty((world * b,), dtype=count.dtype, device=count.device)
    dist.all_gather_into_tensor(cnt_all, send_cnt, group=group)
    dist.all_gather_into_tensor(det_all, send_det, group=group)
    if dst is not None and dist.get_rank(group) != dst:
        return None, None
    return det_all, cnt_all


def interleave_round_robin(det_all: torch.Tensor, cnt_all: torch.Tensor, world: int, num_images: int):
    """Undo the round-robin sharding: rank-major [world, B] order -> original image order, truncated to
    `num_images` (mmdet: zip(*part_list) then [:size], test.py:186-190)."""
    b = det_all.shape[0] // world
    order = torch.arange(world * b, device=det_all.device).view(world, b).t().reshape(-1)[:num_images]
    return det_all[order], cnt_all[order]
