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
