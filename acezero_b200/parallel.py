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


def rows_capacity_per_rank(total_rows, rows_per_image, world_size):
    """Upper bound of the rows one rank fills when images (<= rows_per_image rows each) are dealt round-robin."""
    images = (total_rows + rows_per_image - 1) // rows_per_image
    return ((images + world_size - 1) // world_size + 1) * rows_per_image


def build_row_map(records, world_size, local_stride, total_rows):
    """records: (owner rank, first global row, rows, first local row) per image. Returns int64 numpy array `src` with
    src[g] = owner * local_stride + local_row: where global row g sits in the rank-major all-gathered staging."""
    import numpy as np
    src = np.full(total_rows, -1, dtype=np.int64)
    for owner, g0, n, l0 in records:
        n = min(n, total_rows - g0)
        if n <= 0:
            continue
        src[g0:g0 + n] = owner * local_stride + l0 + np.arange(n, dtype=np.int64)
    if (src < 0).any():
        raise ValueError("row map does not cover the buffer")
    return src


def allgather_buffer_rows(local, records, local_rows, total_rows, world_size, permute_rows, group=None):
    """Replicate the patch buffer that `world_size` ranks filled cooperatively.

    local: dict of this rank's staging arrays (its own images' rows, packed in loader order); local_rows[r]: rows rank r
    filled. Array by array: all-gather the staging (rank-major) and permute the rows into the order of the single-GPU
    buffer (`permute_rows(src_bytes_2d, index_i64, out_bytes_2d)` — the library's row-gather kernel on the GPU). The
    result is bit-identical on every rank to the buffer one GPU would have built, so the epoch permutation of
    ace_trainer.py:466-477 addresses the same rows."""
    stride = max(local_rows)
    some = next(iter(local.values()))
    index = torch.from_numpy(build_row_map(records, world_size, stride, total_rows)).to(some.device)
    out = {}
    for k, t in local.items():
        row_bytes = t.element_size() * t[0].numel()
        mine = t[:stride].contiguous().view(torch.uint8).view(stride, row_bytes)
        gathered = torch.empty((world_size * stride, row_bytes), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        final = torch.empty((total_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        permute_rows(gathered, index, final.view(torch.uint8).view(total_rows, row_bytes))
        out[k] = final
        del gathered
    return out


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
