"""Where does the end-to-end (host batch in, loss out) step time go? Run on a GPU box."""
import time
import torch
import bench
from acezero_b200.head import HeadEngine
from acezero_b200.trainer import TrainLoop, BUFFER_KEYS
from oracle import ace_ref

dev = torch.device("cuda", 0)
B = bench.B
o = bench.options(B, 5000)
head = HeadEngine(1, True, (0.0, 0.0, 0.0), max_rows=B, training=True, device=dev)
head.load_state(ace_ref.make_head_state(200, 1, True))
buf = bench.synth_buffer(200000, dev, 1)
loop = TrainLoop(head, o, buf, use_graph=True)
hbs = []
for i in range(4):
    hb = loop.new_host_batch()
    idx = torch.randint(0, 200000, (B,), device=dev)
    for k in BUFFER_KEYS:
        hb[k].copy_(buf[k][idx])
    hbs.append(hb)
torch.cuda.synchronize()
st = torch.empty_like(hbs[0]["_packed"], device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    st.copy_(hbs[0]["_packed"], non_blocking=True)
torch.cuda.synchronize()
e0.record()
for i in range(20):
    st.copy_(hbs[i % 4]["_packed"], non_blocking=True)
e1.record()
torch.cuda.synchronize()
print("packed H2D us:", e0.elapsed_time(e1) * 50, "GB/s:", st.numel() / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)

def run(n, read, lag=0):
    loop.prefetch_host_batch(hbs[0])
    t0 = time.perf_counter()
    for i in range(n):
        loop.prefetch_host_batch(hbs[(i + 1) % 4])
        loop.train_step_prefetched(read_loss=read, lag=lag)
    loop.drain_prefetched()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e6
    # drain the extra prefetched slot
    loop.train_step_prefetched()
    return dt

run(10, True)
print("pipelined, loss read every step: us/step", run(200, True))
print("pipelined, loss read lag 1:       us/step", run(200, True, 1))
print("pipelined, no loss read:          us/step", run(200, False))
t0 = time.perf_counter()
for i in range(200):
    loop.train_step_from_host(hbs[i % 4])
print("synchronous from_host:            us/step", (time.perf_counter() - t0) / 200 * 1e6)
# host-side issue cost only
t0 = time.perf_counter()
for i in range(200):
    loop.prefetch_host_batch(hbs[(i + 1) % 4])
    loop._stage_r ^= 1; 
t1 = time.perf_counter()
torch.cuda.synchronize()
print("prefetch issue cost us:", (t1 - t0) / 200 * 1e6)
