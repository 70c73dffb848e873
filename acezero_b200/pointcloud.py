"""Point-cloud export from a trained network (reference ace_vis_util.py:431-592 `get_point_cloud_from_network`,
export_point_cloud.py): encoder + head inference on the mapping images, then per image keep the cells whose scene coordinate
re-projects within 1 px of its pixel under the mapping pose, is closer than `filter_depth` and locally smooth — with the
reference's relaxation ladder (gradient thresholds 0.1 / 0.5 / 1 m, at least `pc_points_min / #images` points per image, at
most `pc_points_max / #images`). The per-cell quantities come from ONE fused kernel per micro-batch
(`acez_pointcloud_metrics`); the selection works on those three maps. Returns OpenGL-convention points like the reference.
"""
import struct

import numpy as np
import torch

from . import _lib


def metrics(scene_coords_n3hw, pose_inv_n34, K_n33, subsample=8):
    """(err, grad, depth), each [n, h*w] float32 on the device."""
    lib = _lib.load()
    sc = scene_coords_n3hw.float().contiguous()
    n, _, h, w = sc.shape
    P = pose_inv_n34.float().contiguous()
    K = K_n33.float().contiguous()
    err = torch.empty((n, h * w), device=sc.device, dtype=torch.float32)
    grad = torch.empty_like(err)
    depth = torch.empty_like(err)
    rc = lib.acez_pointcloud_metrics(_lib.ptr(sc), n, h, w, _lib.ptr(P), _lib.ptr(K), subsample, _lib.ptr(err), _lib.ptr(grad),
                                     _lib.ptr(depth), _lib.stream_ptr())
    _lib.check(rc, "acez_pointcloud_metrics")
    return err, grad, depth


def select_points(err, grad, depth, points_min, points_max, filter_depth=100.0, dense_cloud=False, generator=None):
    """Selection mask of ONE image (1-D tensors over its cells), reference ace_vis_util.py:517-558."""
    grad_thresholds = [float("inf")] if dense_cloud else [0.1, 0.5, 1.0, float("inf")]
    repro_threshold = float("inf") if dense_cloud else 1.0
    for g in grad_thresholds:
        grad_mask = grad < g
        if int(grad_mask.sum()) > points_min:
            break
    mask = grad_mask & (depth < filter_depth)
    if int(mask.sum()) == 0:
        mask = torch.ones_like(mask)
    sel = (err < repro_threshold) & mask
    n_valid = int(sel.sum())
    if n_valid < points_min:
        e = torch.sort(err[mask]).values
        relaxed = e[min(points_min, e.shape[0] - 1)]
        sel = mask & (err < relaxed)
    elif n_valid > points_max:
        keep = points_max / n_valid
        sub = torch.randperm(n_valid, generator=generator) < int(keep * n_valid)
        out = sel.clone()
        out[sel] = sub.to(sel.device)
        sel = out
    return sel


def point_cloud_from_network(network, data_loader, filter_depth=100.0, dense_cloud=False, color_fn=None,
                             pc_points_min=100000, pc_points_max=1000000, device="cuda"):
    """N x 3 points (OpenGL convention, as the reference returns them) and N x 3 colours (0..255).

    color_fn(file, index, image_1hw) -> [3, h8*w8] float colours of the cells; default: the network's own grey input image at
    the cell centres (the reference re-reads the RGB file with skimage, ace_vis_util.py:560-577)."""
    n_img = len(data_loader)
    per_min, per_max = int(pc_points_min / n_img), int(pc_points_max / n_img)
    sub = network.OUTPUT_SUBSAMPLE
    xyz, clr = [], []
    with torch.no_grad():
        for image, _, gt_inv_pose, _, K, _, _, file, idx in data_loader:
            image = image.to(device, non_blocking=True)
            sc = network(image).float()                                       # [B,3,h,w]
            B, _, h, w = sc.shape
            err, grad, depth = metrics(sc, gt_inv_pose[:, :3].to(device), K.to(device), sub)
            for b in range(B):
                sel = select_points(err[b], grad[b], depth[b], per_min, per_max, filter_depth, dense_cloud)
                if color_fn is not None:
                    c = color_fn(file[b] if not isinstance(file, str) else file, int(idx[b]) if torch.is_tensor(idx) else int(idx), image[b])
                else:
                    g = (image[b, 0, sub // 2::sub, sub // 2::sub].float() * 0.25 + 0.4).clamp(0, 1) * 255.0   # dataset.py:150-153
                    g = torch.nn.functional.interpolate(g[None, None], size=(h, w), mode="nearest")[0, 0] if g.shape != (h, w) else g
                    c = g.reshape(1, -1).expand(3, -1)
                xyz.append(sc[b].reshape(3, -1)[:, sel].cpu().numpy())
                clr.append(torch.as_tensor(c).to(sel.device)[:, sel].cpu().numpy())
    pc_xyz = np.concatenate(xyz, axis=1).T.copy()
    pc_clr = np.concatenate(clr, axis=1).T.copy()
    pc_xyz[:, 1] = -pc_xyz[:, 1]          # OpenCV -> OpenGL (ace_vis_util.py:586-588)
    pc_xyz[:, 2] = -pc_xyz[:, 2]
    return pc_xyz, pc_clr


def write_txt(path, pc_xyz, pc_clr):
    """`x y z r g b` lines (export_point_cloud.py:117-123)."""
    with open(path, "w") as f:
        for p, c in zip(pc_xyz, pc_clr):
            f.write(f"{p[0]} {p[1]} {p[2]} {c[0]:.0f} {c[1]:.0f} {c[2]:.0f}\n")


def write_ply(path, pc_xyz, pc_clr):
    """Binary little-endian PLY with float xyz + uchar rgb (what trimesh.PointCloud.export writes for the reference,
    export_point_cloud.py:125-130) — written directly, no trimesh dependency."""
    n = int(pc_xyz.shape[0])
    header = ("ply\nformat binary_little_endian 1.0\n" f"element vertex {n}\n"
              "property float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
    rec = np.empty(n, dtype=[("xyz", "<f4", 3), ("rgb", "u1", 3)])
    rec["xyz"] = pc_xyz.astype(np.float32)
    rec["rgb"] = np.clip(np.rint(pc_clr), 0, 255).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())


def read_ply(path):
    """Inverse of write_ply (tests)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    n = int([l for l in data[:end].decode("ascii").splitlines() if l.startswith("element vertex")][0].split()[-1])
    rec = np.frombuffer(data[end:end + n * 15], dtype=[("xyz", "<f4", 3), ("rgb", "u1", 3)])
    return rec["xyz"].copy(), rec["rgb"].copy()
