"""world_size-2 gloo test of the detection gather (the only exchange step of the path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_images, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glsdet_b200.dist import gather_detections, interleave_round_robin, shard_indices

    mine = shard_indices(num_images, rank, world)
    b = (num_images + world - 1) // world
    det = torch.zeros(b, 6, 7)
    cnt = torch.zeros(b, dtype=torch.int32)
    for j, img in enumerate(mine):
        cnt[j] = img % 5 + 1
        det[j, :cnt[j], 0] = img
        det[j, :cnt[j], 4] = torch.arange(int(cnt[j])).float()
    det_all, cnt_all = gather_detections(det, cnt, max_rows=4)
    d, c = interleave_round_robin(det_all, cnt_all, world, num_images)
    ok = d.shape == (num_images, 4, 7)
    for img in range(num_images):
        n = min(img % 5 + 1, 4)
        ok &= int(c[img]) == n and bool((d[img, :n, 0] == img).all())
    only0 = gather_detections(det, cnt, max_rows=4, dst=0)
    ok &= (only0[0] is not None) == (rank == 0)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    world, num_images = 2, 7
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_images, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_gather_without_process_group_is_identity():
    from glsdet_b200.dist import gather_detections

    det, cnt = torch.rand(3, 5, 7), torch.tensor([5, 2, 0], dtype=torch.int32)
    d, c = gather_detections(det, cnt, max_rows=3)
    assert d.shape == (3, 3, 7) and c.tolist() == [3, 2, 0]


class _Loader:
    """Stand-in for a DistributedSampler-backed loader: rank r sees images r, r + world, ... in batches of `bs`."""

    def __init__(self, num_images, rank, world, bs):
        self.dataset = list(range(num_images))
        per = (num_images + world - 1) // world
        mine = [(rank + k * world) % num_images for k in range(per)]      # the sampler pads by wrapping around
        self.batches = [mine[i:i + bs] for i in range(0, len(mine), bs)]

    def __iter__(self):
        for b in self.batches:
            yield {"img_ids": b}

    def __len__(self):
        return len(self.batches)


def _fixed_model(rows):
    def model(return_loss=False, rescale=True, img_ids=None):
        det = torch.zeros(len(img_ids), rows, 7)
        cnt = torch.zeros(len(img_ids), dtype=torch.int32)
        for j, img in enumerate(img_ids):
            cnt[j] = img % rows
            det[j, :cnt[j], 0] = img
            det[j, :cnt[j], 4] = torch.arange(int(cnt[j])).float()
        return det, cnt
    return model


def _mmdet_model(return_loss=False, rescale=True, img_ids=None):
    import numpy as np
    return [[np.full((img % 3, 5), img, dtype=np.float32), np.zeros((0, 5), np.float32)] for img in img_ids]


def _worker_test_loop(rank, world, port, num_images, tmp, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glsdet_b200.dist import DetectionGather, multi_gpu_test

    ok = True
    # fixed-layout detectors: asynchronous double-buffered gather, dataset order on rank 0, None elsewhere
    loader = _Loader(num_images, rank, world, bs=2)
    res = multi_gpu_test(_fixed_model(5), loader, gpu_collect=True)
    if rank == 0:
        ok &= len(res) == num_images
        for img, r in enumerate(res):
            ok &= r.shape == (img % 5, 7) and bool((r[:, 0] == img).all())
    else:
        ok &= res is None
    # mmdet-style results: object gather (gpu_collect) and the tmpdir path
    for kw in (dict(gpu_collect=True), dict(tmpdir=os.path.join(tmp, "parts"))):
        res = multi_gpu_test(_mmdet_model, _Loader(num_images, rank, world, bs=1), **kw)
        if rank == 0:
            ok &= len(res) == num_images and all(r[0].shape == (img % 3, 5) and (r[0] == img).all() for img, r in enumerate(res))
        else:
            ok &= res is None
    # two gathers in flight, consumed late and out of submission order
    g = DetectionGather(2, 3, "cpu")
    t0 = g.submit(torch.full((2, 3, 7), float(rank)), torch.tensor([1, 2], dtype=torch.int32))
    t1 = g.submit(torch.full((2, 3, 7), float(rank + 10)), torch.tensor([3, 9], dtype=torch.int32))
    d1, c1 = g.result(t1)
    d0, c0 = g.result(t0)
    ok &= c0.tolist() == [1, 2] * world and c1.tolist() == [3, 3] * world      # counts clamp to the gathered rows
    ok &= all(bool((d0[2 * r:2 * r + 2] == r).all()) and bool((d1[2 * r:2 * r + 2] == r + 10).all()) for r in range(world))
    # packed layout gathered to rank 0 only: rows back to back, counts in the header, overflow dropped at total_rows
    gp = DetectionGather(3, 4, "cpu", total_rows=6, dst=0)
    det = torch.arange(3 * 4 * 7, dtype=torch.float32).view(3, 4, 7) + 1000 * rank
    tk = gp.submit(det, torch.tensor([2, 0, 3], dtype=torch.int32))
    packed, call = gp.result(tk)
    if rank == 0:
        rows = gp.unpack(packed, call)
        ok &= call.tolist() == [2, 0, 3] * world and [len(r) for r in rows] == [2, 0, 3] * world
        for r in range(world):
            ok &= bool(torch.equal(rows[3 * r], det[0, :2] - 1000 * rank + 1000 * r)) and bool(torch.equal(rows[3 * r + 2], det[2, :3] - 1000 * rank + 1000 * r))
    else:
        ok &= packed is None and call is None
    gp2 = DetectionGather(2, 4, "cpu", total_rows=5, dst=0)      # 4 + 3 rows into 5: the last image is cut
    tk = gp2.submit(torch.ones(2, 4, 7), torch.tensor([4, 3], dtype=torch.int32))
    packed, call = gp2.result(tk)
    if rank == 0:
        ok &= [len(r) for r in gp2.unpack(packed, call)] == [4, 1] * world
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_multi_gpu_test_loop_world2_gloo(tmp_path):
    world, num_images = 2, 9
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_test_loop, args=(world, _free_port(), num_images, str(tmp_path), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_single_process_test_loop_and_fps_harness():
    from glsdet_b200.dist import measure_inference_speed, multi_gpu_test, single_gpu_test

    loader = _Loader(7, 0, 1, bs=3)
    res = single_gpu_test(_fixed_model(4), loader)
    assert [r.shape[0] for r in res] == [i % 4 for i in range(7)]
    res = multi_gpu_test(_fixed_model(4), loader, gpu_collect=True)
    assert [r.shape[0] for r in res] == [i % 4 for i in range(7)]
    lines = []
    fps = measure_inference_speed(_mmdet_model, _Loader(40, 0, 1, bs=1), max_iter=30, log_interval=10, log=lines.append)
    assert fps > 0 and lines[-1].startswith("Overall fps:") and sum(l.startswith("Done image") for l in lines) == 3
