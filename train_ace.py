#!/usr/bin/env python3
"""ACE mapping stage — drop-in for the reference's `train_ace.py` (reference train_ace.py:16-241): same positional
arguments, same flags and defaults (so `ace_zero.py`'s command lines, ace_zero_util.py:63-157, parse unchanged), same
output files (`<map>.pt` fp16 head, `<map>.txt` log, `poses_<map>_preliminary.txt`). The work runs on the sm_100a
kernels through `ace_trainer.TrainerACE`.

Extensions (not in the reference): `--synthetic N` trains on N procedurally rendered frames instead of `rgb_files`
(no image I/O dependency); `--encoder_seed S` uses deterministic random encoder weights when no checkpoint exists.
"""
import argparse
import logging
from pathlib import Path


def _strtobool(x):
    return str(x).lower() in ("1", "true", "yes", "y", "t", "on")


# (flag, type, default) — names and defaults of reference train_ace.py:30-226
_FLAGS = [
    ("--base_seed", int, 2089), ("--pose_files", str, None), ("--use_ace_pose_file", Path, None),
    ("--ace_pose_file_conf_threshold", float, 1000), ("--use_pose_seed", float, -1), ("--depth_files", str, None),
    ("--refine_calibration", _strtobool, False), ("--refine_calibration_lr", float, 0.001),
    ("--use_heuristic_focal_length", _strtobool, False), ("--use_external_focal_length", float, None),
    ("--image_resolution", int, 480), ("--num_data_workers", int, 12),
    ("--encoder_path", Path, Path(__file__).parent / "ace_encoder_pretrained.pt"), ("--load_weights", Path, None),
    ("--num_head_blocks", int, 1), ("--use_half", _strtobool, True), ("--use_homogeneous", _strtobool, True),
    ("--learning_rate_min", float, 0.0005), ("--learning_rate_max", float, 0.005),
    ("--learning_rate_schedule", str, "circle"), ("--learning_rate_warmup_iterations", int, 1000),
    ("--learning_rate_warmup_learning_rate", float, 0.0005), ("--learning_rate_cooldown_iterations", int, 5000),
    ("--learning_rate_cooldown_trigger_px_threshold", int, 10),
    ("--learning_rate_cooldown_trigger_percent_threshold", float, 0.7), ("--max_training_buffer_size", int, 8000000),
    ("--max_dataset_passes", int, 10), ("--samples_per_image", int, 1024), ("--training_buffer_cpu", _strtobool, False),
    ("--batch_size", int, 5120), ("--iterations", int, 25000), ("--iterations_output", int, 300),
    ("--repro_loss_hard_clamp", int, 1000), ("--repro_loss_soft_clamp", int, 50), ("--repro_loss_soft_clamp_min", int, 1),
    ("--repro_loss_type", str, "dyntanh"), ("--repro_loss_schedule", str, "circle"), ("--depth_min", float, 0.1),
    ("--depth_target", float, 10), ("--depth_max", float, 1000), ("--use_aug", _strtobool, True),
    ("--aug_rotation", int, 15), ("--aug_scale", float, 1.5), ("--render_visualization", _strtobool, False),
    ("--render_target_path", Path, Path("renderings")), ("--use_existing_vis_buffer", Path, None),
    ("--render_flipped_portrait", _strtobool, False), ("--render_map_error_threshold", int, 10),
    ("--render_map_depth_filter", int, 100), ("--render_camera_z_offset", int, 4), ("--render_marker_size", float, 0.03),
    ("--pose_refinement", str, "none"), ("--pose_refinement_weight", float, 0.1), ("--pose_refinement_wait", int, 0),
    ("--pose_refinement_lr", float, 0.001), ("--refinement_ortho", str, "gram-schmidt"),
]


def build_parser():
    p = argparse.ArgumentParser(description="Fast training of a scene coordinate regression network (sm_100a).",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("rgb_files", type=str, help="Glob pattern for RGB files, e.g. 'datasets/scene/*.jpg'")
    p.add_argument("output_map_file", type=Path, help="target file for the trained network")
    for flag, typ, default in _FLAGS:
        p.add_argument(flag, type=typ, default=default)
    p.add_argument("--synthetic", type=int, default=0, help="extension: train on N procedural frames")
    p.add_argument("--synthetic_seed", type=int, default=2089)
    p.add_argument("--encoder_seed", type=int, default=None,
                   help="extension: deterministic random encoder weights instead of --encoder_path")
    p.add_argument("--gpus", type=int, default=0,
                   help="extension: GPUs (ranks) for this stage; 0 = $ACEZ_GPUS or 1. More than one: the executable "
                        "re-launches itself under torchrun (acezero_b200/launch.py)")
    return p


def validate(o):
    """reference train_ace.py:232-238"""
    if o.batch_size % 512 != 0:
        raise ValueError("batch_size must be a multiple of 512")
    if o.repro_loss_schedule not in ("circle", "linear"):
        raise ValueError("repro_loss_schedule must be 'circle' or 'linear'")
    if o.pose_refinement not in ("none", "naive", "mlp"):
        raise ValueError("unknown pose_refinement")


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    o = build_parser().parse_args(argv)
    validate(o)
    from acezero_b200 import launch
    # seed trials (ace_zero.py:184-196: one mapping image, several seeds in parallel) stay on one leased GPU each;
    # full mapping stages shard over ACEZ_GPUS / --gpus ranks
    small_job = o.use_pose_seed >= 0
    n_gpus = launch.requested_gpus(o.gpus)
    if o.batch_size % max(n_gpus, 1) != 0:
        raise ValueError(f"batch_size {o.batch_size} is not divisible by {n_gpus} GPUs")
    if n_gpus == 1:
        # opt-in (ACEZ_WORKER): hand the stage to the persistent worker process instead of paying interpreter start, torch import,
        # CUDA context and kernel set-up once per stage (acezero_b200/worker.py); a no-op without the switch or without a worker
        from acezero_b200 import worker
        worker.try_forward("train_ace", argv)
    launch.maybe_self_launch(Path(__file__).resolve(), argv, n_gpus, small_job=small_job)
    import torch
    rank, world = launch.select_device(small_job=small_job)
    if rank != 0:
        logging.getLogger().setLevel(logging.WARNING)
    from ace_trainer import TrainerACE
    dataset = None
    o.encoder_state_dict = None
    if o.encoder_seed is not None:
        from acezero_b200.weights import random_encoder_state
        o.encoder_state_dict = random_encoder_state(o.encoder_seed)
    if o.synthetic > 0:
        from acezero_b200.synthetic import SyntheticDataset
        dataset = SyntheticDataset(o.synthetic, seed=o.synthetic_seed,
                                   focal=o.use_external_focal_length or 525.0,
                                   device=f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu")
        o.num_data_workers = 0
    trainer = TrainerACE(o, dataset=dataset, rank=rank, world_size=world)
    trainer.train()
    if world > 1:
        launch.shutdown_distributed()
    launch.release_gpu()


if __name__ == "__main__":
    main()
