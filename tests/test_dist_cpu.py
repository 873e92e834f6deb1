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
