"""Where the host time of TrainerACE.create_training_buffer goes (cProfile of one warm pass over pre-rendered frames)."""
import cProfile
import pstats
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, ".")
import train_ace  # noqa: E402
from ace_trainer import TrainerACE  # noqa: E402
from acezero_b200.synthetic import CachedDataset, SyntheticDataset  # noqa: E402
from acezero_b200.weights import random_encoder_state  # noqa: E402

n, passes = 64, 4
ds = CachedDataset(SyntheticDataset(n, H=480, W=640, focal=525.0, device="cuda"))
with tempfile.TemporaryDirectory() as tmp:
    o = train_ace.build_parser().parse_args(["synthetic", str(Path(tmp) / "map.pt")])
    o.encoder_state_dict = random_encoder_state(77)
    o.num_data_workers = 0
    o.max_dataset_passes = passes
    tr = TrainerACE(o, dataset=ds)
    tr.create_training_buffer()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    tr.create_training_buffer()
    pr.disable()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
print(f"host {t_host * 1e3:.1f} ms, with device drain {t_all * 1e3:.1f} ms for {n * passes} images")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
