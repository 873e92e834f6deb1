"""Image-sharded multi-GPU execution: one process per GPU, no data-path collective until the end.

Replaces the multi-GPU test path of the reference (yolox-ufp/mmdet/apis/test.py):

  * `collect_results_gpu` (:161-191: pickle -> uint8 tensor -> all_gather(sizes) -> all_gather(padded payload) -> unpickle)
    becomes a fixed-layout gather: detection rows [B, R, 7] fp32 with the per-image count packed into a header row, ONE
    `all_gather_into_tensor` per step, no pickling, no host round trip (`gather_detections`, `DetectionGather`).
    `DetectionGather` issues it asynchronously (NCCL's own stream) into double-buffered outputs, so the gather of step i
    overlaps the compute of step i + 1;
  * `multi_gpu_test` (:70-115) / `single_gpu_test` (:16-67) keep their signatures and loop structure (rank 0 progress,
    results in dataset order, truncated to len(dataset)); detectors of this package return fixed-layout device results,
    anything else (mmdet-style lists of per-class arrays) falls back to an object gather;
  * `collect_results_cpu` (:118-158) is kept for callers that pass tmpdir;
  * `measure_inference_speed` is the FPS loop of tools/analysis_tools/benchmark.py:100-130 (5 warm-up iterations,
    synchronise around every call, same log lines).

NCCL over NVLink on GPUs, gloo on CPU tensors (tests).  Rank r owns images r, r + world, r + 2*world, ...
(DistributedSampler's round-robin order; test.py:186-190 re-interleaves the same way).
"""
from __future__ import annotations

import os
import os.path as osp
import pickle
import shutil
import sys
import tempfile
import time
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def get_dist_info(group=None) -> Tuple[int, int]:
    """mmcv.runner.get_dist_info: (rank, world_size), (0, 1) without a process group."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(num_images: int, rank: int, world: int) -> List[int]:
    return list(range(rank, num_images, world))


def _pack(det: torch.Tensor, count: torch.Tensor, r: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, 1 + r, 7]: row 0 carries the count (exact in fp32 up to 2^24 rows), rows 1.. the detections."""
    b = det.shape[0]
    payload = out if out is not None else torch.empty((b, 1 + r, 7), dtype=det.dtype, device=det.device)
    payload[:, 0, 0] = count.clamp(max=r).to(det.dtype)
    payload[:, 1:] = det[:, :r]
    return payload


def _unpack(payload_all: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return payload_all[:, 1:], payload_all[:, 0, 0].round().to(torch.int32)


def gather_detections(det: torch.Tensor, count: torch.Tensor, max_rows: Optional[int] = None, dst: Optional[int] = None,
                      group=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """det [B, R, 7] fp32, count [B] int32 (same B and R on every rank).  Returns (det_all [world*B, max_rows, 7],
    count_all [world*B]) in rank-major order on every rank (dst=None) or only on rank `dst` (others get None).
    One collective: the counts travel in a header row of the payload."""
    r = det.shape[1] if max_rows is None else min(max_rows, det.shape[1])
    rank, world = get_dist_info(group)
    if world == 1:
        return det[:, :r], count.clamp(max=r)
    payload = _pack(det, count, r)
    b = det.shape[0]
    all_ = torch.empty((world * b, 1 + r, 7), dtype=det.dtype, device=det.device)
    dist.all_gather_into_tensor(all_, payload, group=group)
    if dst is not None and rank != dst:
        return None, None
    return _unpack(all_)


class DetectionGather:
    """Asynchronous, double-buffered detection gather for a pipelined test loop.

        g = DetectionGather(batch, rows, device)                       # padded layout, every rank receives everything
        g = DetectionGather(batch, rows, device, total_rows=T, dst=0)  # packed layout, only rank 0 receives
        t = g.submit(det, cnt)          # after step i: packs on the current stream, the collective runs on NCCL's stream
        ... launch step i + 1 ...
        det_all, cnt_all = g.result(t)  # the current stream waits for gather i only now

    Padded layout: [B, 1 + rows, 7] per rank, the count of an image in its header row.  Packed layout (`total_rows`): the
    rows that exist are stored back to back - [ceil(B / 7) header rows with the counts | sum(count) rows | padding up to
    total_rows] - which is 2 - 4 x smaller than the padded block when survivor counts vary between images (an uncapped
    step of the bench: 2.5 MB instead of 9 MB per rank); rows beyond total_rows are dropped (size it from a probe step).
    `dst`: gather to that rank only (mmdet's collect_results_gpu only uses the result on rank 0, apis/test.py:178-191):
    the other ranks just send.  Two send / receive buffer pairs alternate, so at most two gathers may be outstanding."""

    def __init__(self, batch: int, max_rows: int, device, dtype=torch.float32, group=None, total_rows: Optional[int] = None,
                 dst: Optional[int] = None):
        self.group = group
        self.rank, self.world = get_dist_info(group)
        self.batch, self.rows, self.dst = batch, max_rows, dst
        self.total_rows = total_rows
        self.hdr = (batch + 6) // 7
        shape = (batch, 1 + max_rows, 7) if total_rows is None else (self.hdr + total_rows + 1, 7)   # + 1: dump row
        self.send = [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)]
        recv_here = dst is None or self.rank == dst
        self.recv = [torch.empty((self.world,) + shape, dtype=dtype, device=device) if recv_here else None for _ in range(2)]
        self.work = [None, None]
        self.n = 0
        if total_rows is not None:
            self._ar = torch.arange(max_rows, device=device)[None, :]

    def _pack_rows(self, det: torch.Tensor, count: torch.Tensor, out: torch.Tensor) -> None:
        """Back-to-back rows without a host round trip: row r of image b goes to hdr + offset[b] + r; rows that do not
        exist (or overflow total_rows) go to the dump row at the end."""
        r = self.rows
        cnt = count.clamp(max=r).to(torch.int64)
        off = torch.cumsum(cnt, 0) - cnt
        idx = self.hdr + off[:, None] + self._ar
        dump = self.hdr + self.total_rows
        idx = torch.where((self._ar < cnt[:, None]) & (idx < dump), idx, torch.full_like(idx, dump))
        out[:self.hdr].view(-1)[:self.batch] = cnt.to(out.dtype)
        out.index_copy_(0, idx.reshape(-1), det[:, :r].reshape(-1, 7))

    def pack(self, det: torch.Tensor, count: torch.Tensor) -> int:
        """Build the payload of a step on the current stream; the collective is issued by `launch(slot)`."""
        slot = self.n & 1
        self.n += 1
        if self.work[slot] is not None:   # the buffer pair is reused: its previous gather must have been consumed
            self.work[slot].wait()
            self.work[slot] = None
        if (det.is_cuda and det.dtype == torch.float32 and det.is_contiguous() and count.dtype == torch.int32 and
                det.shape[1] >= self.rows):
            from . import _native as N          # one native launch instead of ~10 framework kernels on the compute stream
            N.check(N.load().glsdet_pack_detections(det.data_ptr(), count.contiguous().data_ptr(), self.batch, det.shape[1],
                                                    self.rows, self.total_rows or 0,
                                                    self.send[slot].data_ptr(), N.stream_ptr()), "glsdet_pack_detections")
        elif self.total_rows is None:
            _pack(det, count, self.rows, out=self.send[slot])
        else:
            self._pack_rows(det, count, self.send[slot])
        return slot

    def launch(self, slot: int) -> None:
        """Issue the collective of a packed slot.  It is ordered after everything enqueued on the current stream so far, so
        WHERE this is called decides what it overlaps: the persistent conv kernels hold every SM (one CTA each), and an NCCL
        kernel that takes an SM away makes one of them wait a full CTA lifetime - call it before the (small-grid)
        post-processing kernels of the next step rather than before its conv segment."""
        if self.world == 1:
            self.recv[slot][0].copy_(self.send[slot])
        elif self.dst is None:
            flat = self.recv[slot].view((-1,) + tuple(self.send[slot].shape[1:]))     # concatenation layout along dim 0
            self.work[slot] = dist.all_gather_into_tensor(flat, self.send[slot], group=self.group, async_op=True)
        else:
            gl = list(self.recv[slot].unbind(0)) if self.rank == self.dst else None
            self.work[slot] = dist.gather(self.send[slot], gl, dst=self.dst, group=self.group, async_op=True)

    def submit(self, det: torch.Tensor, count: torch.Tensor) -> int:
        slot = self.pack(det, count)
        self.launch(slot)
        return slot

    def result(self, slot: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Padded layout: (det_all [world*B, rows, 7], count_all [world*B]).  Packed layout: (packed [world, hdr + T + 1, 7],
        count_all [world*B]) - see `unpack`.  (None, None) on ranks that do not receive."""
        if self.work[slot] is not None:
            self.work[slot].wait()        # stream-level wait on GPU process groups (no host synchronisation)
            self.work[slot] = None
        if self.recv[slot] is None:
            return None, None
        if self.total_rows is None:
            return _unpack(self.recv[slot].view(-1, 1 + self.rows, 7))
        packed = self.recv[slot]
        cnt_all = packed[:, :self.hdr].reshape(self.world, -1)[:, :self.batch].round().to(torch.int32).reshape(-1)
        return packed, cnt_all

    def unpack(self, packed: torch.Tensor, cnt_all: torch.Tensor) -> List[torch.Tensor]:
        """Packed result -> per image a [count, 7] view (host-side slicing: one read of the counts)."""
        counts = cnt_all.view(self.world, self.batch).cpu().tolist()
        out = []
        for r in range(self.world):
            o = self.hdr
            for c in counts[r]:
                n = max(0, min(c, self.hdr + self.total_rows - o))
                out.append(packed[r, o:o + n])
                o += c
        return out


def interleave_round_robin(det_all: torch.Tensor, cnt_all: torch.Tensor, world: int, num_images: int):
    """Undo the round-robin sharding: rank-major [world, B] order -> original image order, truncated to
    `num_images` (mmdet: zip(*part_list) then [:size], test.py:186-190)."""
    b = det_all.shape[0] // world
    order = torch.arange(world * b, device=det_all.device).view(world, b).t().reshape(-1)[:num_images]
    return det_all[order], cnt_all[order]


# ----------------------------------------------------------------------------------------------- mmdet.apis.test
class ProgressBar:
    """Minimal stand-in for mmcv.ProgressBar (rank 0 only prints)."""

    def __init__(self, task_num: int, file=sys.stderr):
        self.task_num, self.done, self.file, self.t0 = task_num, 0, file, time.time()

    def update(self, n: int = 1):
        self.done = min(self.done + n, self.task_num)
        if self.done == self.task_num or self.done % max(1, self.task_num // 20) == 0:
            el = time.time() - self.t0
            self.file.write(f"\r[{self.done}/{self.task_num}] {self.done / max(el, 1e-9):.1f} task/s, elapsed {el:.0f}s")
            self.file.flush()


def _is_fixed_layout(result) -> bool:
    return (isinstance(result, tuple) and len(result) == 2 and isinstance(result[0], torch.Tensor) and
            isinstance(result[1], torch.Tensor) and result[0].dim() == 3 and result[0].shape[-1] == 7)


def _split_fixed(det: torch.Tensor, cnt: torch.Tensor) -> list:
    """(det [B, R, 7], count [B]) -> per image an ndarray [K, 7] (one host copy of the rows that exist)."""
    counts = cnt.cpu().numpy()
    kmax = int(counts.max()) if len(counts) else 0
    rows = det[:, :kmax].cpu().numpy()
    return [rows[i, :counts[i]].copy() for i in range(len(counts))]


def single_gpu_test(model: Callable, data_loader: Iterable, show: bool = False, out_dir=None, show_score_thr: float = 0.3):
    """test.py:16-67 without the visualisation branch (show / out_dir need mmcv image I/O and are refused)."""
    if show or out_dir:
        raise NotImplementedError("single_gpu_test: result visualisation is outside the native path")
    if hasattr(model, "eval"):
        model.eval()
    results = []
    prog_bar = ProgressBar(len(data_loader.dataset))
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **data)
        if _is_fixed_layout(result):
            result = _split_fixed(*result)
        results.extend(result)
        prog_bar.update(len(result))
    return results


def multi_gpu_test(model: Callable, data_loader: Iterable, tmpdir: Optional[str] = None, gpu_collect: bool = False,
                   max_rows: Optional[int] = None):
    """test.py:70-115.  `model(return_loss=False, rescale=True, **data)` returns either the mmdet result list of the batch
    or, for the detectors of this package, the fixed-layout pair (det [B, R, 7], count [B]) on the device; in the latter
    case every batch's gather is issued asynchronously and consumed one batch later (`DetectionGather`), and rank 0 gets
    the per-image arrays in dataset order.  Other ranks return None, like the reference."""
    if hasattr(model, "eval"):
        model.eval()
    results = []
    dataset = data_loader.dataset
    rank, world_size = get_dist_info()
    prog_bar = ProgressBar(len(dataset)) if rank == 0 else None
    gather, pending = None, None
    fixed_parts: List[Tuple[torch.Tensor, torch.Tensor]] = []

    def consume(ticket):
        det_all, cnt_all = gather.result(ticket)     # (None, None) on the ranks that only send
        if rank == 0:   # rank-major [world, B] -> this step's images in dataset order
            b = det_all.shape[0] // world_size
            order = torch.arange(world_size * b, device=det_all.device).view(world_size, b).t().reshape(-1)
            fixed_parts.append((det_all[order].clone(), cnt_all[order].clone()))

    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **data)
        if _is_fixed_layout(result) and (gpu_collect or tmpdir is None):
            det, cnt = result
            if gather is None:
                gather = DetectionGather(det.shape[0], det.shape[1] if max_rows is None else min(max_rows, det.shape[1]), det.device,
                                         dst=0)
            ticket = gather.submit(det, cnt)
            if pending is not None:
                consume(pending)
            pending = ticket
            n = det.shape[0]
        else:
            if _is_fixed_layout(result):
                result = _split_fixed(*result)
            results.extend(result)
            n = len(result)
        if rank == 0:
            prog_bar.update(n * world_size)
    if gather is not None:
        consume(pending)
        if rank != 0:
            return None
        out = []
        for det_all, cnt_all in fixed_parts:
            out.extend(_split_fixed(det_all, cnt_all))
        return out[:len(dataset)]
    if gpu_collect:
        return collect_results_gpu(results, len(dataset))
    return collect_results_cpu(results, len(dataset), tmpdir)


def collect_results_cpu(result_part, size, tmpdir=None):
    """test.py:118-158: every rank pickles its part into a shared directory, rank 0 re-interleaves."""
    rank, world_size = get_dist_info()
    if world_size == 1:
        return result_part[:size]
    backend_cuda = dist.get_backend() == "nccl"
    if tmpdir is None:
        max_len = 512
        dir_tensor = torch.full((max_len,), 32, dtype=torch.uint8, device="cuda" if backend_cuda else "cpu")
        if rank == 0:
            os.makedirs(".dist_test", exist_ok=True)
            name = tempfile.mkdtemp(dir=".dist_test").encode()
            dir_tensor[:len(name)] = torch.tensor(bytearray(name), dtype=torch.uint8, device=dir_tensor.device)
        dist.broadcast(dir_tensor, 0)
        tmpdir = dir_tensor.cpu().numpy().tobytes().decode().rstrip()
    else:
        os.makedirs(tmpdir, exist_ok=True)
    with open(osp.join(tmpdir, f"part_{rank}.pkl"), "wb") as f:
        pickle.dump(result_part, f)
    dist.barrier()
    if rank != 0:
        return None
    part_list = []
    for i in range(world_size):
        with open(osp.join(tmpdir, f"part_{i}.pkl"), "rb") as f:
            part_list.append(pickle.load(f))
    ordered = []
    for res in zip(*part_list):
        ordered.extend(list(res))
    shutil.rmtree(tmpdir)
    return ordered[:size]


def collect_results_gpu(result_part, size):
    """test.py:161-191 for arbitrary picklable results (the fixed-layout path above never gets here)."""
    rank, world_size = get_dist_info()
    if world_size == 1:
        return result_part[:size]
    parts = [None] * world_size
    dist.all_gather_object(parts, result_part)
    if rank != 0:
        return None
    ordered = []
    for res in zip(*parts):
        ordered.extend(list(res))
    return ordered[:size]


# ----------------------------------------------------------------------------------------------- FPS harness
def measure_inference_speed(model: Callable, data_loader: Iterable, max_iter: int = 2000, log_interval: int = 50,
                            num_warmup: int = 5, log=print) -> float:
    """tools/analysis_tools/benchmark.py:100-130: synchronise, time one model call per batch with perf_counter, skip the
    first `num_warmup` iterations, log every `log_interval`, stop at `max_iter`.  fps counts iterations like the reference
    (its loader has one image per batch)."""
    pure_inf_time = 0.0
    fps = 0.0
    i = -1
    sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
    for i, data in enumerate(data_loader):
        sync()
        start_time = time.perf_counter()
        with torch.no_grad():
            model(return_loss=False, rescale=True, **data)
        sync()
        elapsed = time.perf_counter() - start_time
        if i >= num_warmup:
            pure_inf_time += elapsed
            if (i + 1) % log_interval == 0:
                fps = (i + 1 - num_warmup) / pure_inf_time
                log(f"Done image [{i + 1:<3}/ {max_iter}], fps: {fps:.1f} img / s, times per image: {1000 / fps:.1f} ms / img")
        if (i + 1) == max_iter:
            fps = (i + 1 - num_warmup) / pure_inf_time
            log(f"Overall fps: {fps:.1f} img / s, times per image: {1000 / fps:.1f} ms / img")
            break
    else:
        if i >= num_warmup and pure_inf_time > 0:
            fps = (i + 1 - num_warmup) / pure_inf_time
            log(f"Overall fps: {fps:.1f} img / s, times per image: {1000 / fps:.1f} ms / img")
    return fps


def repeat_measure_inference_speed(model: Callable, data_loader: Iterable, max_iter: int = 2000, log_interval: int = 50,
                                   repeat_num: int = 1, log=print):
    """benchmark.py:133-160: repeat the measurement, report the mean and the standard deviation."""
    fps_list = [measure_inference_speed(model, data_loader, max_iter, log_interval, log=log) for _ in range(repeat_num)]
    if repeat_num > 1:
        mean = sum(fps_list) / len(fps_list)
        std = (sum((f - mean) ** 2 for f in fps_list) / len(fps_list)) ** 0.5
        log(f"Overall fps: {[round(f, 1) for f in fps_list]}[{mean:.1f} +- {std:.1f}] img / s")
        return fps_list
    return fps_list[0]
