"""Pose and calibration refinement during mapping (reference refine_poses.py:15-252, refine_calibration.py:7-60).

These are tiny PyTorch-autograd models (a 12->128->12 MLP / one scalar) that stay in PyTorch, as SURVEY.md §2.1 rows 5-6
prescribes; the fused CUDA step hands them dL/dP (b,3,4) and dL/dK00, dL/dK11 (`acez_head_train_fwd_bwd`) exactly where
the reference's autograd graph would. One deliberate restructuring: the refined pose is evaluated once per *image* and
gathered per patch (the reference evaluates the MLP on all 5120 patch rows although only #images distinct inputs exist);
values and gradients are the same function of the same inputs.

`roma` (pinned 1.4.1 in the reference's environment.yml:226, not vendored, not installed here) provides
`special_gramschmidt` / `special_procrustes` in the reference; both are restated below from their documented
definitions — parity for them is unpinned.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import optim


def _capturable(device):
    """torch.optim.AdamW(capturable=True) keeps its step counters on the device: the whole refinement iteration (kernels,
    autograd of the refiners, their optimiser steps) can then be captured into one CUDA graph. Same update rule."""
    return torch.device(device).type == "cuda"


def special_gramschmidt(M):
    """Rotation from the first two columns of M (..., 3, 3): x = M[:, 0] normalised, y = M[:, 1] orthogonalised against
    x and normalised, z = x cross y; columns (x, y, z)."""
    x = F.normalize(M[..., :, 0], dim=-1)
    y = M[..., :, 1]
    y = F.normalize(y - (x * y).sum(-1, keepdim=True) * x, dim=-1)
    z = torch.cross(x, y, dim=-1)
    return torch.stack([x, y, z], dim=-1)


def special_procrustes(M):
    """Closest rotation in the Frobenius sense: U diag(1, 1, det(U V^T)) V^T."""
    U, _, Vh = torch.linalg.svd(M)
    d = torch.det(U @ Vh)
    D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=-1))
    return U @ D @ Vh


class PoseNetwork(nn.Module):
    """reference refine_poses.py:15-72: 1x1-conv MLP 12 -> channels -> 12; module construction order (head_skip, conv1..3,
    blocks, fc1..3) is the reference's so that the default initialisation draws the same random numbers."""

    def __init__(self, num_head_blocks, channels=512):
        super().__init__()
        self.in_channels = 12
        self.head_channels = channels
        self.head_skip = nn.Identity() if self.in_channels == self.head_channels else nn.Conv2d(12, channels, 1, 1, 0)
        self.conv1 = nn.Conv2d(12, channels, 1, 1, 0)
        self.conv2 = nn.Conv2d(channels, channels, 1, 1, 0)
        self.conv3 = nn.Conv2d(channels, channels, 1, 1, 0)
        self.res_blocks = []
        for block in range(num_head_blocks):
            blk = (nn.Conv2d(channels, channels, 1, 1, 0), nn.Conv2d(channels, channels, 1, 1, 0),
                   nn.Conv2d(channels, channels, 1, 1, 0))
            self.res_blocks.append(blk)
            self.add_module(str(block) + 'c0', blk[0])
            self.add_module(str(block) + 'c1', blk[1])
            self.add_module(str(block) + 'c2', blk[2])
        self.fc1 = nn.Conv2d(channels, channels, 1, 1, 0)
        self.fc2 = nn.Conv2d(channels, channels, 1, 1, 0)
        self.fc3 = nn.Conv2d(channels, 12, 1, 1, 0)

    def forward(self, res):
        x = F.relu(self.conv1(res))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        res = self.head_skip(res) + x
        for blk in self.res_blocks:
            x = F.relu(blk[0](res))
            x = F.relu(blk[1](x))
            x = F.relu(blk[2](x))
            res = res + x
        u = F.relu(self.fc1(res))
        u = F.relu(self.fc2(u))
        return self.fc3(u)


class PoseRefiner:
    """reference refine_poses.py:75-252 ('none' / 'naive' / 'mlp')."""

    def __init__(self, dataset, device, options):
        self.dataset = dataset
        self.device = device
        if options.pose_refinement not in ['none', 'naive', 'mlp']:
            raise ValueError(f"Pose refinement strategy {options.pose_refinement} not supported")
        self.refinement_strategy = options.pose_refinement
        self.learning_rate = options.pose_refinement_lr
        self.update_weight = options.pose_refinement_weight
        self.orthonormalization = options.refinement_ortho
        self.pose_buffer = None
        self.pose_buffer_orig = None
        self.pose_network = None
        self.pose_optimizer = None

    @property
    def active(self):
        return self.refinement_strategy != 'none'

    def create_pose_buffer(self):
        self.pose_buffer_orig = torch.zeros(len(self.dataset), 3, 4)
        for i, pose in enumerate(self.dataset.poses):
            self.pose_buffer_orig[i] = torch.as_tensor(pose).float().inverse().clone()[:3]
        self.pose_buffer = self.pose_buffer_orig.contiguous().to(self.device, non_blocking=True)
        if self.refinement_strategy == 'naive':
            self.pose_buffer = self.pose_buffer.detach().requires_grad_()
            self.pose_optimizer = optim.AdamW([self.pose_buffer], lr=self.learning_rate, capturable=_capturable(self.device))
        elif self.refinement_strategy == 'mlp':
            self.pose_network = PoseNetwork(0, 128).to(self.device)
            self.pose_network.train()
            self.pose_optimizer = optim.AdamW(self.pose_network.parameters(), lr=self.learning_rate,
                                              capturable=_capturable(self.device))

    def _orthonormalize_poses(self, poses_b33):
        if self.orthonormalization == 'none':
            return poses_b33
        if self.orthonormalization == 'gram-schmidt':
            return special_gramschmidt(poses_b33)
        return special_procrustes(poses_b33)

    def _predict_pose_updates(self, poses_b34):
        p = poses_b34.reshape(-1, 12, 1, 1)
        upd = (p + self.update_weight * self.pose_network(p)).view(-1, 3, 4)
        return self._orthonormalize_poses(upd[:, :3, :3]), upd[:, :3, 3]

    def current_poses_n34(self):
        """Differentiable current estimate for ALL images (world-to-camera, [N,3,4])."""
        if self.refinement_strategy == 'none':
            return self.pose_buffer
        if self.refinement_strategy == 'naive':
            R = self._orthonormalize_poses(self.pose_buffer[:, :3, :3])
            return torch.cat([R, self.pose_buffer[:, :3, 3:]], dim=2)
        R, t = self._predict_pose_updates(self.pose_buffer)
        return torch.cat([R, t.unsqueeze(2)], dim=2)

    def get_current_poses(self, original_poses_b44, original_poses_indices):
        """Reference signature (refine_poses.py:212-244): refined 4x4 poses for a batch of buffer rows."""
        out = original_poses_b44.clone()
        if self.refinement_strategy == 'none':
            return out
        cur = self.current_poses_n34()[original_poses_indices.view(-1).long()]
        out[:, :3, :4] = cur
        return out

    def get_all_original_poses(self):
        return self.pose_buffer_orig.clone()

    def get_all_current_poses(self):
        with torch.no_grad():
            return self.current_poses_n34().detach().clone().cpu()

    def zero_grad(self, set_to_none=False):
        if self.pose_optimizer is not None:
            self.pose_optimizer.zero_grad(set_to_none=set_to_none)

    def step(self):
        if self.pose_optimizer is not None:
            self.pose_optimizer.step()


class CalibrationRefiner:
    """reference refine_calibration.py:7-60: one relative focal-length scalar shared by all images."""

    def __init__(self, dataset, learning_rate, device):
        import numpy as np
        focal_lengths = [dataset.get_focal_length(i) for i in range(len(dataset))]
        if not np.allclose(focal_lengths, focal_lengths[0]):
            raise ValueError("All images must have the same focal length for calibration refinement")
        self.focal_length_init = focal_lengths[0]
        self.device = device
        self.global_f = torch.zeros(1).to(device).detach().requires_grad_()
        self.optimizer = optim.AdamW([self.global_f], lr=learning_rate, capturable=_capturable(device))

    def get_focal_length(self):
        return (1 + self.global_f) * self.focal_length_init

    def get_refined_calibration_matrices(self, Ks_b33):
        refined_22 = torch.eye(2, 2, device=Ks_b33.device) * self.get_focal_length()
        refined_b22 = refined_22.unsqueeze(0).expand(Ks_b33.shape[0], -1, -1)
        aug_scales = Ks_b33[:, 0, 0] / self.focal_length_init
        refined_scaled_b22 = refined_b22 * aug_scales.detach()[:, None, None]
        out = Ks_b33.clone().detach()
        out[:, :2, :2] = refined_scaled_b22
        return out

    def zero_grad(self):
        self.optimizer.zero_grad()

    def step(self):
        self.optimizer.step()
