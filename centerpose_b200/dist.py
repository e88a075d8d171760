"""Multi-GPU plumbing: frames are independent, so a batch is
sharded over ranks with no data-path collective except ONE all-gather of the
fixed-shape pose tensor (+ n_valid packed into the same buffer).  Replaces the
reference's training-only single-process DataParallel
(models/data_parallel.py:10-84) on the inference path.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n, rank, world):
    """Contiguous [start, stop) of frame indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack(poses, n_valid):
    """[b,K,R] fp32 + [b] int32 -> [b, K*R + 1] fp32 (n_valid rides in the last column; exact below 2^24)."""
    b = poses.shape[0]
    return torch.cat([poses.reshape(b, -1), n_valid.to(torch.float32).reshape(b, 1)], dim=1).contiguous()


def unpack(buf, K, R=_lib.CP_POSE_RECORD):
    n = buf.shape[0]
    return buf[:, :K * R].reshape(n, K, R), buf[:, K * R].round().to(torch.int32)


def all_gather_poses(poses, n_valid, group=None):
    """One collective: every rank receives the pose records of the whole batch, in rank order.
    Requires equal per-rank batch (the weak-scaling configuration)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return poses, n_valid
    world = dist.get_world_size(group)
    local = pack(poses, n_valid)
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return unpack(out, poses.shape[1], poses.shape[2])
