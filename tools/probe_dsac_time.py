"""Poses/s of the batched DSAC* solver (64 hypotheses, 1024 maps of 60x80 cells per call), as bench.py measures it."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from acezero_b200 import dsac

n = 1024
maps = torch.from_numpy(bench.synth_scene_maps(n, 1000)).cuda()
kw = dict(hyps=64, inlier_threshold=10.0, inlier_alpha=100.0, max_reproj=100.0, subsample=8, seed=2089, max_tries=16)
for _ in range(3):
    dsac.forward_rgb_batch(maps, 525.0, 320.0, 240.0, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    p, k = dsac.forward_rgb_batch(maps, 525.0, 320.0, 240.0, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{n} poses in {ms:.3f} ms = {n / ms * 1000:.0f} poses/s; mean inliers {k.float().mean().item():.0f}")
