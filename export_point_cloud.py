#!/usr/bin/env python3
"""Point-cloud export — drop-in for the reference's `export_point_cloud.py` (reference export_point_cloud.py:24-132): same
arguments, `.txt` / `.ply` by file ending, OpenGL or OpenCV convention; extraction from a network + pose file runs encoder,
head and the filter metrics on the sm_100a kernels (acezero_b200/pointcloud.py). A visualisation buffer (pickle with
`map_xyz` / `map_clr`) is passed through as in the reference.

Extensions: `--synthetic N` / `--encoder_seed S` as in train_ace.py (no image files needed)."""
import argparse
import logging
import pickle
from pathlib import Path

_logger = logging.getLogger(__name__)


def _strtobool(x):
    return str(x).lower() in ("1", "true", "yes", "y", "t", "on")


def build_parser():
    p = argparse.ArgumentParser(description="Extract point cloud from network (slow) or visualization buffer file (fast). "
                                            "File ending determines output format where txt and ply are supported.",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("output_file", type=Path)
    p.add_argument("--network", type=Path, help="network to extract point cloud from.")
    p.add_argument("--pose_file", type=Path, help="pose file of images that trained the network")
    p.add_argument("--visualization_buffer", type=Path, help="Vis buffer files that contains a pre-calculated point cloud.")
    p.add_argument("--encoder_path", type=Path, default=Path(__file__).parent / "ace_encoder_pretrained.pt")
    p.add_argument("--image_resolution", type=int, default=480)
    p.add_argument("--confidence_threshold", type=int, default=500)
    p.add_argument("--convention", type=str, default="opengl", choices=["opengl", "opencv"])
    p.add_argument("--dense_point_cloud", type=_strtobool, default=False)
    p.add_argument("--synthetic", type=int, default=0)
    p.add_argument("--synthetic_seed", type=int, default=2089)
    p.add_argument("--encoder_seed", type=int, default=None)
    return p


def main(argv=None):
    logging.basicConfig(level=logging.INFO)
    parser = build_parser()
    opt = parser.parse_args(argv)
    if opt.visualization_buffer is None and (opt.network is None or (opt.pose_file is None and opt.synthetic <= 0)):
        parser.error("You must provide either a visualization buffer or network and pose file.")
    if opt.dense_point_cloud and opt.visualization_buffer is not None:
        parser.error("A dense cloud cannot be extracted from a visualization buffer. Please provide network and pose file.")
    if opt.output_file.suffix not in (".txt", ".ply"):
        raise ValueError(f"Output file format {opt.output_file.suffix} not supported.")
    from acezero_b200 import pointcloud
    if opt.visualization_buffer is None:
        import torch
        from torch.utils.data import DataLoader
        from ace_network import Regressor
        _logger.info("Extracting point cloud from network.")
        if opt.encoder_seed is not None:
            from acezero_b200.weights import random_encoder_state
            encoder_state_dict = random_encoder_state(opt.encoder_seed)
        else:
            encoder_state_dict = torch.load(opt.encoder_path, map_location="cpu")
        head_state_dict = torch.load(opt.network, map_location="cpu")
        network = Regressor.create_from_split_state_dict(encoder_state_dict, head_state_dict).to("cuda").eval()
        if opt.synthetic > 0:
            from acezero_b200.synthetic import SyntheticDataset
            dataset = SyntheticDataset(opt.synthetic, seed=opt.synthetic_seed, device="cuda")
            workers = 0
        else:
            try:
                from dataset import CamLocDataset
            except ImportError as e:
                raise RuntimeError(f"the reference's dataset.CamLocDataset cannot be imported ({e}); dataset I/O is outside the "
                                   "hot path — run inside an ACE0 checkout or use --synthetic") from e
            dataset = CamLocDataset(rgb_files=None, image_short_size=opt.image_resolution, ace_pose_file=opt.pose_file,
                                    ace_pose_file_conf_threshold=opt.confidence_threshold)
            workers = 6
        _logger.info(f"Images found: {len(dataset)}")
        loader = DataLoader(dataset, shuffle=False, num_workers=workers)
        pc_xyz, pc_clr = pointcloud.point_cloud_from_network(network, loader, filter_depth=100, dense_cloud=opt.dense_point_cloud)
    else:
        _logger.info("Extracting point cloud from visualization buffer.")
        with open(opt.visualization_buffer, "rb") as f:
            state = pickle.load(f)
        pc_xyz, pc_clr = state["map_xyz"], state["map_clr"]
    if opt.convention == "opencv":     # OpenGL to OpenCV convention
        pc_xyz[:, 1] = -pc_xyz[:, 1]
        pc_xyz[:, 2] = -pc_xyz[:, 2]
    if opt.output_file.suffix == ".txt":
        pointcloud.write_txt(opt.output_file, pc_xyz, pc_clr)
    else:
        pointcloud.write_ply(opt.output_file, pc_xyz, pc_clr)
    _logger.info(f"Done. Wrote point cloud to: {opt.output_file}")


if __name__ == "__main__":
    main()
