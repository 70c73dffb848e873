"""The device-side schedule (csrc/schedule.cu: lr, loss weight, cool-down trigger, max_iterations evaluated by the first kernel
of every iteration) against the host restatement of ScheduleACE (acezero_b200.trainer.Schedule, itself checked against the
torch scheduler objects the reference instantiates in tests/test_parallel_cpu.py) and ace_loss.py:53-69."""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _opts(kind, iterations, **kw):
    o = SimpleNamespace(learning_rate_schedule=kind, iterations=iterations, learning_rate_min=5e-4, learning_rate_max=3e-3,
                        learning_rate_warmup_iterations=50, learning_rate_warmup_learning_rate=5e-4,
                        learning_rate_cooldown_iterations=100, learning_rate_cooldown_trigger_percent_threshold=0.7,
                        repro_loss_type="dyntanh", repro_loss_schedule="circle", repro_loss_soft_clamp=50,
                        repro_loss_soft_clamp_min=1, batch_size=5120)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _params(o):
    from acezero_b200 import _lib
    sp = _lib.ScheduleParams()
    sp.kind = _lib.SCHED_KINDS[o.learning_rate_schedule]
    sp.iterations = o.iterations
    sp.lr_min, sp.lr_max = o.learning_rate_min, o.learning_rate_max
    sp.warmup_iterations, sp.warmup_lr = o.learning_rate_warmup_iterations, o.learning_rate_warmup_learning_rate
    sp.cooldown_iterations = o.learning_rate_cooldown_iterations
    sp.cooldown_trigger = o.learning_rate_cooldown_trigger_percent_threshold
    sp.batch_global = o.batch_size
    sp.loss_dyntanh = int(o.repro_loss_type == "dyntanh")
    sp.loss_schedule_circle = int(o.repro_loss_schedule == "circle")
    sp.soft_clamp, sp.soft_clamp_min = o.repro_loss_soft_clamp, o.repro_loss_soft_clamp_min
    return sp


@pytest.mark.parametrize("kind,inl_fn", [("circle", lambda i: 0.5), ("constant", lambda i: 0.5),
                                         ("1cyclepoly", lambda i: 0.5 if i < 120 else 0.9),      # dynamic trigger
                                         ("1cyclepoly", lambda i: 0.3)])                         # trigger by duration
def test_device_schedule_follows_schedule_ace(kind, inl_fn):
    from acezero_b200 import _lib
    from acezero_b200.trainer import Schedule, loss_weight
    lib = _lib.load()
    o = _opts(kind, 400)
    sp = _params(o)
    dev = torch.device("cuda")
    state = torch.zeros(_lib.SCHED_STATE_FLOATS, device=dev)
    hyper = torch.zeros(8, device=dev)
    inl = torch.zeros(1, device=dev)
    _lib.check(lib.acez_schedule_init(C.byref(sp), _lib.ptr(state), _lib.stream_ptr()))
    sch = Schedule(o)
    it = 0
    while True:
        # host restatement of one TrainLoop iteration: check_and_set_cooldown, stop test, lr of this iteration, step(inliers)
        sch.check_and_set_cooldown(it)
        done = it >= sch.max_iterations
        _lib.check(lib.acez_schedule_step(C.byref(sp), _lib.ptr(state), _lib.ptr(inl), _lib.ptr(hyper), _lib.stream_ptr()))
        h = hyper.cpu().numpy()
        st = state.cpu().numpy()
        assert int(st[4]) == sch.max_iterations, (it, st[4], sch.max_iterations)
        assert bool(st[2]) == sch.in_cooldown_phase
        if done:
            assert st[7] == 1 and h[0] == 0.0 and int(st[0]) == it       # frozen: lr 0, the counter stops
            break
        assert abs(h[0] - np.float32(sch.lr())) <= 2e-7 * sch.lr() + 1e-12, (it, h[0], sch.lr())
        assert abs(h[5] - np.float32(loss_weight(o, it))) < 1e-4
        assert int(st[0]) == it + 1
        frac = inl_fn(it)
        inl.fill_(frac * o.batch_size)        # what the tail kernel leaves in stats[1] after this iteration
        sch.step(frac)
        it += 1
    assert it == sch.max_iterations
    if kind == "1cyclepoly":
        assert sch.in_cooldown_phase and (sch.max_iterations < 400) == (inl_fn(399) > 0.7)
    # iterations enqueued past the end stay frozen
    _lib.check(lib.acez_schedule_step(C.byref(sp), _lib.ptr(state), _lib.ptr(inl), _lib.ptr(hyper), _lib.stream_ptr()))
    assert float(hyper[0]) == 0.0 and int(state[0]) == it


def test_training_stops_at_the_dynamic_max_iterations_without_per_iteration_sync():
    """TrainLoop with 1cyclepoly: the cool-down is triggered on the device; the host learns the shortened max_iterations from
    lagged snapshots and ends with exactly that many iterations; the weights equal those of a run that syncs every step."""
    import bench
    from acezero_b200.head import HeadEngine
    from acezero_b200.trainer import TrainLoop
    from oracle import ace_ref
    dev = torch.device("cuda")
    b = 1024
    buf = bench.synth_buffer(8 * b, dev, 5)
    finals = []
    for want_every in (False, True):
        o = bench.options(b, 400)
        o.learning_rate_schedule = "1cyclepoly"
        o.learning_rate_max = 3e-3
        o.learning_rate_warmup_iterations = 20
        o.learning_rate_cooldown_iterations = 60
        o.learning_rate_cooldown_trigger_percent_threshold = -1.0     # every inlier fraction passes: trigger at warm-up end
        head = HeadEngine(1, True, (0.0, 0.0, 0.0), max_rows=b, training=True, device=dev)
        head.load_state(ace_ref.make_head_state(200, 1, True))
        loop = TrainLoop(head, o, buf, use_graph=True)
        perm = torch.randperm(8 * b, generator=loop.training_generator)
        n = 0
        while loop.train_iteration(perm[(n % 8) * b:(n % 8 + 1) * b], want_stats=want_every):
            n += 1
            assert n < 400
        loop.finish()
        assert loop.schedule.max_iterations == 20 + 60 and loop.iteration == 80
        finals.append(head.params.clone())
    assert torch.equal(finals[0], finals[1])
