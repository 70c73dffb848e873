"""world_size-2 `gloo` tests of the multi-GPU host logic (no GPU): batch sharding, gradient / flag all-reduce
semantics, registration ownership + gather, and that every rank derives the same permutation / schedule."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acezero_b200 import parallel
    from acezero_b200.trainer import Schedule
    from types import SimpleNamespace
    res = {}
    # identical permutation on every rank (training_generator seed, ace_trainer.py:79-80) and disjoint shards
    g = torch.Generator(); g.manual_seed(2089 + 8191)
    perm = torch.randperm(20480, generator=g)
    lo, hi = parallel.shard_bounds(rank, world, 5120)
    res["shard"] = perm[:5120][lo:hi].clone()
    # gradient sum; stats sums; flags OR
    grads = torch.full((1000,), float(rank + 1))
    stats = torch.tensor([1.5 * (rank + 1), 10.0 * (rank + 1), 100.0, float(rank == 1)])
    found = torch.tensor([1 if rank == 0 else 0], dtype=torch.int32)
    parallel.allreduce_training_state(grads, stats, found)
    res["grads"], res["stats"], res["found"] = grads, stats, found
    # registration: ownership + gather on rank 0
    mine = [{"index": i, "pose": np.eye(4) * i, "inliers": 100 + i, "file": f"f{i}", "focal": 525.0}
            for i in range(7) if parallel.image_owner(i, world) == rank]
    merged = parallel.gather_registration(mine, world)
    res["merged"] = None if merged is None else [m["index"] for m in merged]
    # the schedule is a pure function of the iteration count
    o = SimpleNamespace(learning_rate_schedule="circle", iterations=100, learning_rate_min=5e-4, learning_rate_max=5e-3)
    sch = Schedule(o)
    lrs = []
    for _ in range(5):
        lrs.append(sch.lr()); sch.step(0.0)
    res["lrs"] = lrs
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    g = torch.Generator(); g.manual_seed(2089 + 8191)
    perm = torch.randperm(20480, generator=g)[:5120]
    assert torch.equal(torch.cat([a["shard"], b["shard"]]), perm)            # shards tile the global batch, in order
    for r in (a, b):
        assert torch.equal(r["grads"], torch.full((1000,), 3.0))
        assert torch.allclose(r["stats"], torch.tensor([4.5, 30.0, 200.0, 1.0]))
        assert int(r["found"]) == 1
    assert a["merged"] == list(range(7)) and b["merged"] is None
    assert a["lrs"] == b["lrs"]


def test_shard_bounds_validation():
    from acezero_b200 import parallel
    assert parallel.shard_bounds(3, 8, 5120) == (1920, 2560)
    with pytest.raises(ValueError):
        parallel.shard_bounds(0, 3, 5120)


def test_schedule_matches_torch_schedulers():
    """Schedule (pure functions) == the torch scheduler objects the reference instantiates (ace_schedule.py:22-69)."""
    from types import SimpleNamespace
    from acezero_b200.trainer import Schedule
    from torch import optim
    p = [torch.zeros(1, requires_grad=True)]
    # circle
    o = SimpleNamespace(learning_rate_schedule="circle", iterations=500, learning_rate_min=5e-4, learning_rate_max=5e-3)
    opt = optim.AdamW(p, lr=o.learning_rate_min)
    ref = optim.lr_scheduler.OneCycleLR(opt, max_lr=o.learning_rate_max, total_steps=500, cycle_momentum=False)
    sch = Schedule(o)
    for i in range(499):
        assert abs(sch.lr() - opt.param_groups[0]["lr"]) < 1e-12, i
        opt.step(); ref.step(); sch.step(0.0)
    # 1cyclepoly: warm-up, dynamic cool-down trigger, shortened max_iterations
    o = SimpleNamespace(learning_rate_schedule="1cyclepoly", iterations=400, learning_rate_min=5e-4, learning_rate_max=3e-3,
                        learning_rate_warmup_iterations=50, learning_rate_warmup_learning_rate=5e-4,
                        learning_rate_cooldown_iterations=100, learning_rate_cooldown_trigger_percent_threshold=0.7)
    opt = optim.AdamW(p, lr=o.learning_rate_max)
    warm = optim.lr_scheduler.LinearLR(opt, start_factor=o.learning_rate_warmup_learning_rate / o.learning_rate_max,
                                       total_iters=50)
    cool = optim.lr_scheduler.LinearLR(opt, start_factor=1, end_factor=o.learning_rate_min / o.learning_rate_max,
                                       total_iters=100)
    sch = Schedule(o)
    cur, buf, max_it, in_cool = warm, [], 400, False
    it = 0
    while it < max_it:
        # reference check_and_set_cooldown (ace_schedule.py:72-101)
        if not in_cool and it >= 50 and (it >= max_it - 100 or min(buf) > 0.7):
            cur, max_it, in_cool = cool, it + 100, True
        sch.check_and_set_cooldown(it)
        assert sch.max_iterations == max_it and sch.in_cooldown_phase == in_cool
        assert abs(sch.lr() - opt.param_groups[0]["lr"]) < 1e-9, (it, sch.lr(), opt.param_groups[0]["lr"])
        inl = 0.5 if it < 120 else 0.9
        opt.step(); cur.step(); sch.step(inl)
        buf.append(inl); buf = buf[-100:]
        it += 1
    assert max_it < 400


# ------------------------------------------------------------------------------------------------------------------
# sharded buffer creation: row map + all-gather of the ranks' staging rows
# ------------------------------------------------------------------------------------------------------------------
def _records(world, n_images, rows_per_image, total_rows, ragged_at=None):
    """The bookkeeping TrainerACE.create_training_buffer keeps on every rank: images dealt round-robin in loader order."""
    records, local_rows, g = [], [0] * world, 0
    for k in range(n_images):
        n = rows_per_image if k != ragged_at else rows_per_image // 3       # e.g. a partly masked-out image
        n = min(n, total_rows - g)
        if n <= 0:
            break
        owner = k % world
        records.append((owner, g, n, local_rows[owner]))
        local_rows[owner] += n
        g += n
    return records, local_rows, g


def _permute_rows_cpu(src, index, out):
    out.copy_(src.index_select(0, index))


def _buffer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acezero_b200 import parallel
    records, local_rows, total = _records(world, 11, 8, 80, ragged_at=4)
    # the single-GPU buffer: row g holds (g, 2g) / int16 g
    full = {"features": torch.arange(total, dtype=torch.float16).view(-1, 1).repeat(1, 4),
            "target_px": torch.stack([torch.arange(total, dtype=torch.float32), 2 * torch.arange(total, dtype=torch.float32)], 1),
            "pose_idx": torch.arange(total, dtype=torch.int16).view(-1, 1)}
    cap = parallel.rows_capacity_per_rank(80, 8, world)
    local = {k: torch.zeros((cap,) + tuple(v.shape[1:]), dtype=v.dtype) for k, v in full.items()}
    for owner, g0, n, l0 in records:
        if owner == rank:
            for k in full:
                local[k][l0:l0 + n] = full[k][g0:g0 + n]
    merged = parallel.allgather_buffer_rows(local, records, local_rows, total, world, permute_rows=_permute_rows_cpu)
    out[rank] = all(torch.equal(merged[k], full[k]) for k in full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_buffer_rows_allgather_to_the_single_gpu_buffer():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_buffer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] and out[1]


def test_row_map_covers_every_row_once():
    from acezero_b200 import parallel
    records, local_rows, total = _records(4, 37, 1024, 36000, ragged_at=9)
    stride = max(local_rows)
    src = parallel.build_row_map(records, 4, stride, total)
    assert len(np.unique(src)) == total                       # a permutation: every staged row is used exactly once
    owner = src // stride
    assert (owner[:1024] == 0).all() and (owner[1024:2048] == 1).all()
    with pytest.raises(ValueError):
        parallel.build_row_map(records[:-1], 4, stride, total)
    assert parallel.rows_capacity_per_rank(36000, 1024, 4) >= stride


# ------------------------------------------------------------------------------------------------------------------
# stage-executable placement (acezero_b200/launch.py)
# ------------------------------------------------------------------------------------------------------------------
def test_launch_policy(tmp_path, monkeypatch):
    from acezero_b200 import launch
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("ACEZ_GPUS", raising=False)
    assert launch.requested_gpus(0) == 1 and launch.requested_gpus(4) == 4
    monkeypatch.setenv("ACEZ_GPUS", "8")
    assert launch.requested_gpus(0) == 8 and launch.requested_gpus(2) == 2
    assert launch.world() == (0, 1, 0) and not launch.in_worker()
    cmd = launch.torchrun_command("train_ace.py", ["a", "b.pt", "--iterations", 5], 8, port=29999)
    assert cmd[1:4] == ["-m", "torch.distributed.run", "--nnodes=1"] and "--nproc-per-node=8" in cmd
    assert cmd[cmd.index("--master-addr") + 1] == "127.0.0.1" and cmd[-5:] == ["train_ace.py", "a", "b.pt", "--iterations", "5"]
    # single GPU, a worker of a group, or a small job (seed trial): the caller does the work itself (no exec)
    launch.maybe_self_launch("train_ace.py", [], 1)
    launch.maybe_self_launch("train_ace.py", [], 8, small_job=True)
    monkeypatch.setenv("RANK", "3"); monkeypatch.setenv("WORLD_SIZE", "8"); monkeypatch.setenv("LOCAL_RANK", "3")
    launch.maybe_self_launch("train_ace.py", [], 8)
    assert launch.world() == (3, 8, 3)


def _lease_worker(lock_dir, q, hold):
    from acezero_b200 import launch
    q.put(launch.lease_gpu(4, lock_dir=lock_dir, blocking_fallback=False))
    hold.wait(20)


def test_parallel_seed_workers_lease_distinct_gpus(tmp_path):
    """Four concurrent single-GPU stage processes (ace_zero.py --seed_parallel_workers 4) on a 4-GPU box: one GPU each."""
    ctx = mp.get_context("spawn")
    q, hold = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_lease_worker, args=(str(tmp_path), q, hold)) for _ in range(4)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=60) for _ in procs)
    hold.set()
    for p in procs:
        p.join(30)
    assert got == [0, 1, 2, 3]
