"""Buffer-fill rate of ACE mapping (reference ace_trainer.py:293-452): images/s through
`TrainerACE.create_training_buffer` = encoder (batch 1) + mask resize + torch.multinomial (bit-exact indices) + fused
scatter kernel, on pre-rendered 480x640 frames (the renderer / dataset I/O is outside the hot path and is not timed).

    python tools/bench_buffer_fill.py [n_images] [passes]
"""
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ds = CachedDataset(SyntheticDataset(n, H=480, W=640, focal=525.0, device="cuda"))
    with tempfile.TemporaryDirectory() as tmp:
        o = train_ace.build_parser().parse_args(["synthetic", str(Path(tmp) / "map.pt")])
        o.encoder_state_dict = random_encoder_state(77)
        o.num_data_workers = 0
        o.samples_per_image = 1024
        o.max_dataset_passes = passes
        tr = TrainerACE(o, dataset=ds)
        tr.create_training_buffer()            # warm-up (plans, allocations)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.create_training_buffer()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    rows = tr.training_buffer_size
    print(f"{n} images x {passes} passes: {rows} rows ({rows * 1230 / 1e6:.0f} MB) in {dt * 1e3:.1f} ms = "
          f"{n * passes / dt:.0f} images/s, {rows * 1230 / dt / 1e9:.2f} GB/s of buffer rows")


if __name__ == "__main__":
    main()
