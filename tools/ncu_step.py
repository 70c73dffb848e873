"""Eight eager (no CUDA graph) training iterations at b = 5120 for `ncu`: every kernel of the step is a plain launch.

    ncu --set full --clock-control none --import-source on -k regex:'head_chain4|gemm2cta|head_tail|adamw_kernel|fc3_reduce|gather_rows' \
        --launch-skip 35 --launch-count 7 -o gpurun_out/step -f python tools/ncu_step.py
(7 matching launches per iteration: gather, forward chain, tail, fc3 reduce, dgrad chain, weight-gradient GEMM, AdamW)
"""
import sys
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
loop = TrainLoop(head, bench.options(B), buf, use_graph=False)
perm = torch.randperm(262144, generator=loop.training_generator)
for i in range(8):
    loop.train_iteration(perm[i * B:(i + 1) * B])
torch.cuda.synchronize()
print("loss", float(head.stats[0]))
