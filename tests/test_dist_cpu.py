"""world_size-2 gloo test of the multi-GPU host logic: batch sharding + the single all-gather of pose records."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centerpose_b200 import _lib as L
from centerpose_b200.dist import PoseBuffer, all_gather_poses, pack, shard_range, unpack


def test_shard_range_covers_batch():
    for n in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    p = torch.randn(3, 100, L.CP_POSE_RECORD)
    n = torch.tensor([0, 7, 100], dtype=torch.int32)
    p2, n2 = unpack(pack(p, n), 3, 100)
    assert torch.equal(p, p2) and torch.equal(n, n2)


def test_pose_buffer_views_alias_one_flat_tensor():
    buf = PoseBuffer(4, 100, "cpu", world=1)
    buf.poses.copy_(torch.arange(buf.n_pose, dtype=torch.float32).view(4, 100, L.CP_POSE_RECORD))
    buf.n_valid.copy_(torch.tensor([3, 0, 100, 7], dtype=torch.int32))
    assert buf.poses.data_ptr() == buf.flat.data_ptr() and buf.poses.is_contiguous() and buf.n_valid.is_contiguous()
    buf.all_gather()
    buf.to_host()
    p, n = buf.host_views()
    assert p.shape == (4, 100, L.CP_POSE_RECORD) and float(p[3, 99, -1]) == buf.n_pose - 1
    assert n.tolist() == [3, 0, 100, 7]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, K = 6, 100
        lo, hi = shard_range(B, rank, world)
        g = torch.Generator().manual_seed(0)
        full = torch.randn(B, K, L.CP_POSE_RECORD, generator=g)
        nv = torch.arange(B, dtype=torch.int32)
        poses, n_valid = all_gather_poses(full[lo:hi].clone(), nv[lo:hi].clone())
        ok = torch.equal(poses, full) and torch.equal(n_valid, nv)
        # the persistent-buffer form the bench / serving path uses
        buf = PoseBuffer(hi - lo, K, "cpu", world=world, pin=False)
        buf.poses.copy_(full[lo:hi])
        buf.n_valid.copy_(nv[lo:hi])
        buf.all_gather()
        buf.to_host()
        hp, hn = buf.host_views()
        ok = ok and np.array_equal(hp, full.numpy()) and hn.tolist() == nv.tolist()
        dp, dn = buf.views(buf.gathered)
        ok = ok and torch.equal(dp, full) and torch.equal(dn, nv)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_all_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
