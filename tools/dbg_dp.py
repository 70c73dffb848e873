import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import bench
from acezero_b200.head import HeadEngine
from acezero_b200.trainer import TrainLoop
from acezero_b200 import parallel
from oracle import ace_ref
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
o = bench.options(1024, iterations=1000)
head = HeadEngine(1, True, (0, 0, 0), max_rows=512, training=True, device=dev)
head.load_state(ace_ref.make_head_state(200, 1, True))
buf = bench.synth_buffer(65536, dev, 7)
loop = TrainLoop(head, o, buf, rank=rank, world_size=world, use_graph=(len(sys.argv) > 1))
perm = torch.randperm(65536, generator=loop.training_generator)
for i in range(8):
    loop.train_iteration(perm[i * 1024:(i + 1) * 1024], want_stats=True)
    torch.cuda.synchronize()
    print(f"[r{rank}] it{i} found {int(head.found_inf)} scale {float(head.scaler_state[0])} step {float(head.scaler_state[2])} cnt {head.scaler_state[3].view(torch.int32).item()} loss {float(loop.last_stats[0]):.4f} max|g| {float(head.grads.abs().max()):.1f}", flush=True)
dist.destroy_process_group()
