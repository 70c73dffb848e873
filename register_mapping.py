#!/usr/bin/env python3
"""ACE registration stage — drop-in for the reference's `register_mapping.py` (reference register_mapping.py:39-298):
same positional arguments / flags / defaults, same output `poses_<session>.txt` next to the network file (world-to-cam
quaternion lines, confidence = inlier count). Encoder, head and DSAC* run on the GPU; scene coordinates never leave it.

Extensions: `--synthetic N`, `--encoder_seed S` as in train_ace.py.
"""
import argparse
import logging
import random
import time
from pathlib import Path

import numpy as np

_logger = logging.getLogger(__name__)


def _strtobool(x):
    return str(x).lower() in ("1", "true", "yes", "y", "t", "on")


def build_parser():
    p = argparse.ArgumentParser(description="Estimate camera poses with a trained ACE network (sm_100a).",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("rgb_files", type=str)
    p.add_argument("network", type=Path, help="path to a network trained for the scene (just the head weights)")
    p.add_argument("--encoder_path", type=Path, default=Path(__file__).parent / "ace_encoder_pretrained.pt")
    p.add_argument("--session", "-sid", default="")
    p.add_argument("--image_resolution", type=int, default=480)
    p.add_argument("--num_data_workers", type=int, default=12)
    p.add_argument("--hypotheses", "-hyps", type=int, default=64)
    p.add_argument("--hypotheses_max_tries", type=int, default=1000000)
    p.add_argument("--threshold", "-t", type=float, default=10)
    p.add_argument("--inlieralpha", "-ia", type=float, default=100)
    p.add_argument("--maxpixelerror", "-maxerrr", type=float, default=100)
    p.add_argument("--render_visualization", type=_strtobool, default=False)
    p.add_argument("--render_target_path", type=Path, default=Path("renderings"))
    p.add_argument("--render_flipped_portrait", type=_strtobool, default=False)
    p.add_argument("--render_pose_conf_threshold", type=int, default=5000)
    p.add_argument("--render_map_depth_filter", type=int, default=10)
    p.add_argument("--render_camera_z_offset", type=int, default=4)
    p.add_argument("--base_seed", type=int, default=1305)
    p.add_argument("--confidence_threshold", type=float, default=1000)
    p.add_argument("--max_estimates", type=int, default=-1)
    p.add_argument("--use_external_focal_length", type=float, default=-1)
    p.add_argument("--render_marker_size", type=float, default=0.03)
    p.add_argument("--synthetic", type=int, default=0)
    p.add_argument("--synthetic_seed", type=int, default=2089)
    p.add_argument("--synthetic_offset", type=int, default=0, help="first trajectory index of the synthetic frames")
    p.add_argument("--encoder_seed", type=int, default=None)
    p.add_argument("--gpus", type=int, default=0,
                   help="extension: GPUs (ranks) for this stage; 0 = $ACEZ_GPUS or 1 (acezero_b200/launch.py)")
    p.add_argument("--micro_batch", type=int, default=16, help="extension: images per DSAC* launch pair")
    p.add_argument("--loader_batch", type=int, default=8,
                   help="extension: images per DataLoader step (collated on the workers, pinned); 1 = the reference's loader")
    return p


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    opt = build_parser().parse_args(argv)
    if opt.render_visualization:
        raise NotImplementedError("the visualiser is out of scope (SURVEY §2.1 row 13)")
    from acezero_b200 import launch
    # the fast registration check of a seed trial (ace_zero_util.py:242-259: --max_estimates 1000) stays on one leased GPU
    small_job = opt.max_estimates > 0
    if launch.requested_gpus(opt.gpus) == 1:
        from acezero_b200 import worker   # opt-in persistent stage worker (ACEZ_WORKER), see train_ace.py
        worker.try_forward("register_mapping", argv)
    launch.maybe_self_launch(Path(__file__).resolve(), argv, launch.requested_gpus(opt.gpus), small_job=small_job)
    import torch
    rank, world = launch.select_device(small_job=small_job)
    if rank != 0:
        logging.getLogger().setLevel(logging.WARNING)
    from torch.utils.data import DataLoader
    from ace_network import Regressor
    from acezero_b200 import posefile
    from acezero_b200.registration import register

    torch.manual_seed(opt.base_seed)
    np.random.seed(opt.base_seed)
    random.seed(opt.base_seed)
    device = torch.device("cuda", torch.cuda.current_device())

    if opt.synthetic > 0:
        from acezero_b200.synthetic import SyntheticDataset
        testset = SyntheticDataset(opt.synthetic, seed=opt.synthetic_seed,
                                   focal=opt.use_external_focal_length if opt.use_external_focal_length > 0 else 525.0,
                                   device=str(device), indices=range(opt.synthetic_offset, opt.synthetic_offset + opt.synthetic))
        workers = 0
    else:
        try:
            from dataset import CamLocDataset
        except ImportError as e:
            raise RuntimeError(f"the reference's dataset.CamLocDataset cannot be imported ({e}); dataset I/O is outside "
                               "the hot path (SURVEY §2.1 row 9) — run inside an ACE0 checkout or use --synthetic") from e
        testset = CamLocDataset(rgb_files=opt.rgb_files, image_short_size=opt.image_resolution,
                                use_heuristic_focal_length=opt.use_external_focal_length < 0)
        if opt.use_external_focal_length > 0:
            testset.set_external_focal_length(opt.use_external_focal_length)
        workers = opt.num_data_workers
    # reference :147 (shuffle: the order only affects the line order of the pose file). Batches are collated by the workers
    # and pinned, so that the main process only enqueues copies and kernels
    from acezero_b200.registration import collate_same_size
    loader = DataLoader(testset, shuffle=True, num_workers=workers, batch_size=max(1, opt.loader_batch),
                        collate_fn=collate_same_size, pin_memory=True)

    if opt.encoder_seed is not None:
        from acezero_b200.weights import random_encoder_state
        encoder_state_dict = random_encoder_state(opt.encoder_seed)
    else:
        encoder_state_dict = torch.load(opt.encoder_path, map_location="cpu")
    head_state_dict = torch.load(opt.network, map_location="cpu")
    network = Regressor.create_from_split_state_dict(encoder_state_dict, head_state_dict).to(device)
    network.eval()

    pose_log_file = Path(opt.network).parent / f"poses_{opt.session}.txt"
    _logger.info(f"Saving per-frame poses and errors to: {pose_log_file}")
    results, stats = register(network, loader, opt.hypotheses, opt.threshold, opt.inlieralpha, opt.maxpixelerror,
                              opt.base_seed, opt.hypotheses_max_tries, opt.max_estimates, micro_batch=opt.micro_batch,
                              rank=rank, world_size=world, device=device)
    if world > 1:
        # image i was solved by rank i % world (per-image RNG key: the poses do not depend on the number of ranks);
        # rank 0 collects the (pose, inlier count) rows and writes the file
        from acezero_b200.parallel import gather_registration
        results = gather_registration(results, world)
        launch.shutdown_distributed()
        if rank != 0:
            return
    launch.release_gpu()
    with open(pose_log_file, "w", 1) as pose_log:
        for r in results:
            _logger.info(f"Frame: {r['file']}, Confidence: {r['inliers']}")
            posefile.write_pose_to_pose_file(pose_log, rgb_file=r["file"], pose=np.linalg.inv(r["pose"]),
                                             confidence=r["inliers"], focal_length=testset.get_focal_length(r["index"]))
    if stats["images"]:
        _logger.info(f"Avg. processing time: {stats['seconds'] / stats['images'] * 1000:4.1f}ms")


if __name__ == "__main__":
    main()
