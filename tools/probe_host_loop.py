"""Host cost of one training iteration: enqueue time of TrainLoop.train_iteration with an EMPTY device queue ahead of it
(the loop is timed in bursts of 20 iterations right after a synchronize, so the host never blocks on the launch queue),
next to the device time per iteration. If host >= device the step is launch-bound, not kernel-bound."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
from acezero_b200.head import HeadEngine
from acezero_b200.trainer import TrainLoop
from oracle import ace_ref

dev = torch.device("cuda")
B = 5120
head = HeadEngine(1, True, (0, 0, 0), max_rows=B, training=True)
head.load_state(ace_ref.make_head_state(200, 1, True))
buf = bench.synth_buffer(262144, dev, 1)
loop = TrainLoop(head, bench.options(B), buf, use_graph=True)
perm = torch.randperm(262144, generator=loop.training_generator)
nb = 262144 // B
for i in range(6):
    loop.train_iteration(perm[(i % nb) * B:(i % nb + 1) * B])
torch.cuda.synchronize()
host = []
for burst in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        loop.train_iteration(perm[(i % nb) * B:(i % nb + 1) * B])
    host.append((time.perf_counter() - t0) / 20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for i in range(400):
    loop.train_iteration(perm[(i % nb) * B:(i % nb + 1) * B])
e1.record()
torch.cuda.synchronize()
print(f"host enqueue per iteration: median {sorted(host)[len(host) // 2] * 1e6:.1f} us (min {min(host) * 1e6:.1f}); "
      f"device per iteration (400 back to back): {e0.elapsed_time(e1) / 400 * 1000:.1f} us", flush=True)
