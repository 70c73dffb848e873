"""Warm per-segment timing of one training iteration (each segment captured in its own CUDA graph, replayed 50x)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
import bench
from acezero_b200 import _lib
from acezero_b200.head import HeadEngine
from acezero_b200.trainer import TrainLoop
from oracle import ace_ref

dev = torch.device("cuda")
B = 5120
head = HeadEngine(1, True, (0, 0, 0), max_rows=B, training=True)
head.load_state(ace_ref.make_head_state(200, 1, True))
buf = bench.synth_buffer(262144, dev, 1)
loop = TrainLoop(head, bench.options(B), buf, use_graph=False)
perm = torch.randperm(262144, generator=loop.training_generator)
for i in range(3):
    loop.train_iteration(perm[i * B:(i + 1) * B])
torch.cuda.synchronize()
lib = head.lib
o = loop.o
lp = head.loss_params(o.repro_loss_type, 30.0, B)
bt = loop.batch

def seg_gather(): loop._gather()
def seg_fwd(): _lib.check(lib.acez_head_forward(head.plan, None, B, None, _lib.stream_ptr()))
def seg_fwd_tail():
    sc = seg_fwd_tail.sc
    _lib.check(lib.acez_head_forward(head.plan, None, B, _lib.ptr(sc), _lib.stream_ptr()))
seg_fwd_tail.sc = torch.empty((B, 3), device=dev)
def seg_fwd_bwd():
    head.train_fwd_bwd(B, lp, bt["target_px"], bt["intrinsics"], bt["intrinsics_inv"], aug_inv=bt["aug_poses_inv"],
                       pose_inv=bt["poses_inv"], use_device_scale=True)
def seg_adamw(): head.adamw_step(use_scaler=True)
def seg_all():
    seg_gather(); seg_fwd_bwd(); seg_adamw()

def timeit(name, fn, reps=50):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps * 1000:8.1f} us", flush=True)

timeit("gather", seg_gather)
timeit("fwd GEMM chain (8)", seg_fwd)
timeit("fwd chain + fwd-only tail", seg_fwd_tail)
timeit("fwd + tail + bwd (full)", seg_fwd_bwd)
timeit("adamw (+scaler)", seg_adamw)
timeit("whole iteration", seg_all)
