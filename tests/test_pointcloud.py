"""Point-cloud export: the fused metrics kernel against a torch restatement of the reference's
formulae (ace_vis_util.py:489-515), the selection ladder, and the writers."""
import numpy as np
import pytest
import torch


def _ref_metrics(sc, pose_inv_n34, K, sub):
    """ace_vis_util.py:484-515 with plain torch ops."""
    n, _, h, w = sc.shape
    X = sc.flatten(2)
    Xh = torch.cat([X, torch.ones_like(X[:, :1])], 1)
    cam = torch.matmul(pose_inv_n34, Xh)
    px = torch.matmul(K, cam)
    px[:, 2].clamp_(min=0.1)
    uv = px[:, :2] / px[:, 2, None]
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    grid = (sub * (torch.stack([xx, yy]) + 0.5)).view(2, -1).to(sc.device)
    err = (uv - grid[None]).abs().sum(1)
    gx = torch.linalg.norm(sc[:, :, :, 1:] - sc[:, :, :, :-1], dim=1)
    gx = torch.nn.functional.pad(gx, (1, 0), mode="reflect")
    gy = torch.linalg.norm(sc[:, :, 1:, :] - sc[:, :, :-1, :], dim=1)
    gy = torch.nn.functional.pad(gy, (0, 0, 1, 0), mode="reflect")
    return err, torch.max(gx, gy).view(n, -1), cam[:, 2]


@pytest.mark.gpu
def test_metrics_kernel_matches_the_reference_formulae():
    from acezero_b200 import pointcloud
    g = torch.Generator().manual_seed(3)
    n, h, w = 3, 60, 80
    sc = (torch.randn((n, 3, h, w), generator=g) * 2 + torch.tensor([0.0, 0.0, 4.0]).view(1, 3, 1, 1)).cuda()
    P = torch.eye(4)[:3].repeat(n, 1, 1)
    P[:, :, 3] = torch.randn((n, 3), generator=g) * 0.2
    K = torch.tensor([[525.0, 0, 320], [0, 525.0, 240], [0, 0, 1]]).repeat(n, 1, 1)
    err, grad, depth = pointcloud.metrics(sc, P.cuda(), K.cuda(), 8)
    e2, g2, d2 = _ref_metrics(sc, P.cuda(), K.cuda(), 8)
    assert torch.allclose(err, e2, rtol=1e-4, atol=1e-2)
    assert torch.allclose(grad, g2, rtol=1e-5, atol=1e-6)
    assert torch.allclose(depth, d2, rtol=1e-5, atol=1e-5)


def test_selection_ladder_and_writers(tmp_path):
    from acezero_b200 import pointcloud
    n = 4800
    g = torch.Generator().manual_seed(1)
    err = torch.rand(n, generator=g) * 5
    grad = torch.rand(n, generator=g) * 0.05
    depth = torch.rand(n, generator=g) * 10
    sel = pointcloud.select_points(err, grad, depth, points_min=100, points_max=2000)
    assert int(sel.sum()) == int((err < 1.0).sum())                          # enough points within 1 px: the plain filter
    sel = pointcloud.select_points(err + 3.0, grad, depth, points_min=100, points_max=2000)
    assert 95 <= int(sel.sum()) <= 101                                      # none within 1 px: the 100 best are kept
    sel = pointcloud.select_points(err * 0.01, grad, depth, points_min=100, points_max=2000)
    assert int(sel.sum()) <= 2000                                           # all within 1 px: sub-sampled to the maximum
    sel = pointcloud.select_points(err, grad + 2.0, depth + 1000.0, points_min=100, points_max=2000)
    assert int(sel.sum()) > 0                                               # nothing survives depth / gradient: keep all, then filter
    xyz = np.random.RandomState(0).randn(50, 3)
    clr = np.random.RandomState(1).uniform(0, 255, (50, 3))
    pointcloud.write_ply(tmp_path / "a.ply", xyz, clr)
    x2, c2 = pointcloud.read_ply(tmp_path / "a.ply")
    assert np.allclose(x2, xyz.astype(np.float32)) and np.array_equal(c2, np.rint(clr).astype(np.uint8))
    pointcloud.write_txt(tmp_path / "a.txt", xyz, clr)
    rows = [l.split() for l in (tmp_path / "a.txt").read_text().strip().splitlines()]
    assert len(rows) == 50 and len(rows[0]) == 6 and abs(float(rows[3][1]) - xyz[3, 1]) < 1e-12


@pytest.mark.gpu
def test_export_cli_on_a_procedural_scene(tmp_path):
    """export_point_cloud.py end to end on a head trained for a few hundred iterations: the cloud is non-empty, finite, inside
    the room, and both formats / conventions are written."""
    import train_ace
    import export_point_cloud
    train_ace.main(["synthetic", str(tmp_path / "map.pt"), "--synthetic", "8", "--encoder_seed", "7", "--iterations", "400",
                    "--use_external_focal_length", "525", "--max_dataset_passes", "4", "--iterations_output", "200"])
    export_point_cloud.main([str(tmp_path / "pc.ply"), "--network", str(tmp_path / "map.pt"), "--synthetic", "8", "--encoder_seed", "7"])
    from acezero_b200 import pointcloud
    xyz, rgb = pointcloud.read_ply(tmp_path / "pc.ply")
    # 8 images x 4800 cells, each image asked for 12500 points: the relaxation keeps (almost) every cell within depth range
    assert 8 * 4000 <= xyz.shape[0] <= 8 * 4800 and np.isfinite(xyz).all() and rgb.shape == (xyz.shape[0], 3)
    export_point_cloud.main([str(tmp_path / "pc.txt"), "--network", str(tmp_path / "map.pt"), "--synthetic", "8", "--encoder_seed", "7",
                             "--convention", "opencv", "--dense_point_cloud", "True"])
    assert (tmp_path / "pc.txt").stat().st_size > 1000
