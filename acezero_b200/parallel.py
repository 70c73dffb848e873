"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

  training      rank r owns rows [r*b/G, (r+1)*b/G) of every batch (identical permutation on every rank); the loss
                divisor is the global batch; one all-reduce(SUM) of the flat fp32 gradient + one of 5 scalars
                (loss sum, inlier count, valid count, non-finite-loss flag, GradScaler inf flag — the skip decision
                must be global)
  registration  image i belongs to rank i % G; per-image RNG keys make results independent of G; rank 0 gathers
                (pose, inliers) rows and writes the pose file
"""
import torch
import torch.distributed as dist


def shard_bounds(rank, world_size, b_global):
    if b_global % world_size != 0:
        raise ValueError(f"batch size {b_global} is not divisible by {world_size} ranks")
    b = b_global // world_size
    return rank * b, (rank + 1) * b


def image_owner(index, world_size):
    return int(index) % world_size


def allreduce_training_state(grads, stats, found_inf, group=None):
    """In place: grads <- sum over ranks; stats[0:3] <- sums, stats[3] / found_inf <- logical OR."""
    dist.all_reduce(grads, group=group)
    small = torch.cat([stats.float(), found_inf.float()])
    dist.all_reduce(small, group=group)
    stats[:3].copy_(small[:3])
    stats[3:4].copy_((small[3:4] > 0).to(stats.dtype))
    found_inf.copy_((small[4:] > 0).to(found_inf.dtype))


def gather_registration(results, world_size, group=None):
    """results: list of dicts with 'index', 'pose' (4x4 numpy), 'inliers', 'file', 'focal' on each rank.
    Returns the merged list (sorted by dataset index) on rank 0, None elsewhere."""
    if world_size == 1:
        return sorted(results, key=lambda r: r["index"])
    gathered = [None] * world_size if dist.get_rank(group) == 0 else None
    dist.gather_object(results, gathered, dst=0, group=group)
    if gathered is None:
        return None
    merged = [r for part in gathered for r in part]
    return sorted(merged, key=lambda r: r["index"])
