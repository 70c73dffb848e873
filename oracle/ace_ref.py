"""CPU oracle for the ACE mapping hot path — TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (CPU) restatement of the reference's head forward (ace_network.py:120-149), training step
(ace_trainer.py:499-640), loss (ace_loss.py:39-90) and optimiser / GradScaler step (ace_schedule.py:106-126), written
on [b,512] rows instead of the fake BCHW view. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path never does.

PARITY STATUS: pinned. tests/golden/ace_train_golden.npz holds the outputs of the *reference's own*
`TrainerACE.training_step`, `Regressor` and `ReproLoss` executed on CPU by oracle/make_golden.py (run in the build
container where /root/reference exists); tests/test_oracle_golden.py checks this restatement against them.

Two arithmetic modes:
  emulate_half=False : everything fp32 — what the reference computes on a CPU (autocast / GradScaler disable
                       themselves without CUDA); this is the mode the golden vectors pin.
  emulate_half=True  : inserts fp16 roundings where CUDA autocast (ace_trainer.py:517, use_half default True) has
                       them: conv weights / bias / outputs / residual sums in fp16 with fp32 accumulation, softplus
                       and the loss in fp32, activation gradients rounded to fp16 on the way back, weight gradients
                       rounded to fp16, GradScaler dynamics (init 65536, x2 / 2000 steps, x0.5 + skip on inf).
                       The CUDA kernels are compared against this mode.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

HEAD_CHANNELS = 512  # ace_network.py:76


# --------------------------------------------------------------------------------------------------------------------
# head parameters
# --------------------------------------------------------------------------------------------------------------------
def head_layer_names(num_head_blocks):
    """Hidden 512->512 layers in forward order (ace_network.py:83-103,122-137)."""
    names = ["res3_conv1", "res3_conv2", "res3_conv3"]
    for b in range(num_head_blocks):
        names += [f"{b}c0", f"{b}c1", f"{b}c2"]
    return names + ["fc1", "fc2"]


def make_head_state(seed, num_head_blocks=1, use_homogeneous=True, mean=(0.0, 0.0, 0.0), scale=1.0):
    """Deterministic (numpy RandomState) head state dict with the reference's key names / shapes
    (ace_network.py:83-118; OIHW 1x1 conv weights). Values follow nn.Conv2d's default init scale
    (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    rs = np.random.RandomState(seed)
    sd = {}
    bound = scale / math.sqrt(HEAD_CHANNELS)
    for n in head_layer_names(num_head_blocks):
        sd[n + ".weight"] = torch.from_numpy(rs.uniform(-bound, bound, (512, 512, 1, 1)).astype(np.float32))
        sd[n + ".bias"] = torch.from_numpy(rs.uniform(-bound, bound, (512,)).astype(np.float32))
    c3 = 4 if use_homogeneous else 3
    sd["fc3.weight"] = torch.from_numpy(rs.uniform(-bound, bound, (c3, 512, 1, 1)).astype(np.float32))
    sd["fc3.bias"] = torch.from_numpy(rs.uniform(-bound, bound, (c3,)).astype(np.float32))
    if use_homogeneous:  # buffers, ace_network.py:108-114
        max_scale, min_scale = 4.0, 0.01
        sd["max_scale"] = torch.tensor([max_scale])
        sd["min_scale"] = torch.tensor([min_scale])
        sd["max_inv_scale"] = 1.0 / sd["max_scale"]
        sd["h_beta"] = math.log(2) / (1.0 - sd["max_inv_scale"])
        sd["min_inv_scale"] = 1.0 / sd["min_scale"]
    sd["mean"] = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    return sd


class _RoundHalf(torch.autograd.Function):
    """x -> fp16 -> fp32 with the gradient rounded the same way (an autocast dtype boundary)."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.half().float()


def _rh(x, on):
    return _RoundHalf.apply(x) if on else x


def head_forward(sd, x_b512, num_head_blocks, use_homogeneous, emulate_half=False):
    """ace_network.py:120-149 on rows. `sd` maps reference state-dict names to (possibly requires_grad) tensors."""
    def conv(x, name, relu=True):
        W = _rh(sd[name + ".weight"].reshape(sd[name + ".weight"].shape[0], -1), emulate_half)
        b = _rh(sd[name + ".bias"], emulate_half)
        y = _rh(x @ W.t() + b, emulate_half)
        return F.relu(y) if relu else y

    res = _rh(x_b512.float(), emulate_half)
    x = conv(res, "res3_conv1")
    x = conv(x, "res3_conv2")
    x = conv(x, "res3_conv3")
    res = _rh(res + x, emulate_half)                      # head_skip is Identity for 512-d features (:81,126)
    for b in range(num_head_blocks):                      # :128-133
        x = conv(res, f"{b}c0")
        x = conv(x, f"{b}c1")
        x = conv(x, f"{b}c2")
        res = _rh(res + x, emulate_half)
    sc = conv(res, "fc1")
    sc = conv(sc, "fc2")
    sc = conv(sc, "fc3", relu=False)
    if use_homogeneous:                                   # :139-144 (softplus runs in fp32 under autocast)
        h = F.softplus(sc[:, 3:4], beta=float(sd["h_beta"])) + sd["max_inv_scale"].float()
        h = torch.clamp(h, max=float(sd["min_inv_scale"]))
        sc = sc[:, :3] / h
    return sc + sd["mean"].float().view(1, 3)             # :147


# --------------------------------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------------------------------
class LossOptions:
    """The train_ace.py defaults the training step reads (train_ace.py:137-176)."""

    def __init__(self, **kw):
        self.repro_loss_type = "dyntanh"
        self.repro_loss_schedule = "circle"
        self.repro_loss_soft_clamp = 50.0
        self.repro_loss_soft_clamp_min = 1.0
        self.repro_loss_hard_clamp = 1000.0
        self.depth_min = 0.1
        self.depth_max = 1000.0
        self.depth_target = 10.0
        self.learning_rate_cooldown_trigger_px_threshold = 10.0
        self.iterations = 25000
        self.use_depth = False
        for k, v in kw.items():
            setattr(self, k, v)


def loss_weight(o, iteration):
    """ace_loss.py:53-69: the tanh weight at `iteration` (python float / numpy double, as the reference)."""
    if o.repro_loss_type == "tanh":
        return o.repro_loss_soft_clamp
    if o.repro_loss_type == "dyntanh":
        w = iteration / o.iterations
        if o.repro_loss_schedule == "circle":
            w = 1 - np.sqrt(1 - w ** 2)
        return float((1 - w) * o.repro_loss_soft_clamp + o.repro_loss_soft_clamp_min)
    return o.repro_loss_soft_clamp


def repro_loss_compute(o, errs, iteration):
    """ace_loss.py:39-90."""
    if errs.nelement() == 0:
        return 0
    if o.repro_loss_type in ("tanh", "dyntanh"):
        w = loss_weight(o, iteration)
        return w * torch.tanh(errs / w).sum()
    m = errs > o.repro_loss_soft_clamp
    l1 = errs[~m].sum()
    if o.repro_loss_type == "l1":
        return l1
    if o.repro_loss_type == "l1+sqrt":
        return l1 + torch.sqrt(o.repro_loss_soft_clamp * errs[m]).sum()
    return l1 + torch.log(1 + o.repro_loss_soft_clamp * errs[m]).sum()


def training_loss(o, sc_b3, target_px_b2, inv_aug_poses_b34, inv_poses_b44, Ks_b33, invKs_b33, target_crds_b3,
                  iteration, P_b34=None):
    """ace_trainer.py:521-613. Returns (loss, batch_inliers, n_valid)."""
    b = sc_b3.shape[0]
    pred_b31 = sc_b3.unsqueeze(-1).float()
    pred_b41 = torch.cat([pred_b31, torch.ones_like(pred_b31[:, :1])], dim=1)              # :524
    gt_inv_poses_b34 = torch.bmm(inv_aug_poses_b34, inv_poses_b44) if P_b34 is None else P_b34  # :530
    cam_b31 = torch.bmm(gt_inv_poses_b34, pred_b41)                                        # :533
    px_b31 = torch.bmm(Ks_b33, cam_b31)                                                    # :540
    px_b31[:, 2].clamp_(min=o.depth_min)                                                   # :545
    px_b21 = px_b31[:, :2] / px_b31[:, 2, None]                                            # :548
    err_b2 = px_b21.squeeze() - target_px_b2                                               # :551
    err_b1 = torch.norm(err_b2, dim=1, keepdim=True, p=1)                                  # :552
    invalid = (cam_b31[:, 2] < o.depth_min) | (err_b1 > o.repro_loss_hard_clamp) | (cam_b31[:, 2] > o.depth_max)
    if o.use_depth:                                                                        # :567-574
        far = (torch.linalg.norm(target_crds_b3 - pred_b31.squeeze(), dim=1) > 0.1).unsqueeze(1)
        avail = (target_crds_b3.abs().sum(dim=1) > 0.00001).unsqueeze(1)
        invalid = invalid | (far & avail)
    valid = ~invalid
    if valid.sum() > 0:                                                                    # :578-589
        verr = err_b1[valid]
        loss_valid = repro_loss_compute(o, verr, iteration)
        inl = float((verr < o.learning_rate_cooldown_trigger_px_threshold).sum()) / b
    else:
        loss_valid, inl = 0, 0
    if not o.use_depth:                                                                    # :592-600
        grid_b31 = torch.cat([target_px_b2.unsqueeze(2), torch.ones_like(target_px_b2[:, :1]).unsqueeze(2)], dim=1)
        tgt_b31 = o.depth_target * torch.bmm(invKs_b33, grid_b31)
        loss_invalid = torch.abs(tgt_b31 - cam_b31).masked_select(invalid.unsqueeze(2)).sum()
    else:                                                                                  # :602-609
        if invalid.sum() > 0:
            inv2 = invalid & avail
            li = torch.linalg.norm(target_crds_b3 - pred_b31.squeeze(), dim=1)
            loss_invalid = li[inv2.squeeze()].sum()
        else:
            loss_invalid = 0
    loss = (loss_valid + loss_invalid) / b                                                 # :612-613
    return loss, inl, int(valid.sum())


# --------------------------------------------------------------------------------------------------------------------
# one training iteration: forward + loss + backward + GradScaler + AdamW
# --------------------------------------------------------------------------------------------------------------------
class TrainerRef:
    """Restates ace_trainer.py:499-640 + ace_schedule.py:106-126 for the head (no pose / calibration refinement).

    lr_fn(iteration) -> learning rate applied at that iteration (the scheduler is a pure function of the iteration
    count; ace_schedule.py steps it unconditionally every iteration)."""

    def __init__(self, sd, num_head_blocks, use_homogeneous, opts, lr_fn, emulate_half=False, use_scaler=None):
        self.names = head_layer_names(num_head_blocks) + ["fc3"]
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.params = []
        for n in self.names:
            for s in (".weight", ".bias"):
                p = self.sd[n + s].clone().float().requires_grad_(True)
                self.sd[n + s] = p
                self.params.append(p)
        self.nb, self.homog, self.o, self.lr_fn = num_head_blocks, use_homogeneous, opts, lr_fn
        self.emulate_half = emulate_half
        self.use_scaler = emulate_half if use_scaler is None else use_scaler
        self.opt = torch.optim.AdamW(self.params, lr=lr_fn(0))  # defaults: betas (.9,.999), eps 1e-8, wd 0.01
        self.scale, self.growth_tracker = 65536.0, 0
        self.iteration = 0
        self.skipped = 0

    def forward(self, features_b512):
        return head_forward(self.sd, features_b512, self.nb, self.homog, self.emulate_half)

    def step(self, features_b512, target_px_b2, inv_aug_b34, inv_poses_b44, Ks_b33, invKs_b33, target_crds_b3=None):
        sc = self.forward(features_b512)
        loss, inl, n_valid = training_loss(self.o, sc, target_px_b2, inv_aug_b34, inv_poses_b44, Ks_b33, invKs_b33,
                                           target_crds_b3, self.iteration)
        self.opt.zero_grad(set_to_none=True)
        S = self.scale if self.use_scaler else 1.0
        (loss * S).backward()
        for g in self.opt.param_groups:
            g["lr"] = self.lr_fn(self.iteration)
        found_inf = False
        if self.use_scaler:
            for p in self.params:
                if p.grad is None:
                    continue
                g16 = p.grad.half().float() if self.emulate_half else p.grad  # autocast: weight grads are fp16
                found_inf |= not bool(torch.isfinite(g16).all())
                p.grad = g16 / S
        if not found_inf:
            self.opt.step()
        else:
            self.skipped += 1
        if self.use_scaler:  # GradScaler.update()
            if found_inf:
                self.scale *= 0.5
                self.growth_tracker = 0
            else:
                self.growth_tracker += 1
                if self.growth_tracker == 2000:
                    self.scale *= 2.0
                    self.growth_tracker = 0
        self.iteration += 1
        return float(loss), inl, n_valid, sc.detach()


# --------------------------------------------------------------------------------------------------------------------
# learning-rate schedules as pure functions of the iteration (ace_schedule.py:22-69)
# --------------------------------------------------------------------------------------------------------------------
def one_cycle_lr(max_lr, total_steps, pct_start=0.3, div_factor=25.0, final_div_factor=1e4):
    """torch.optim.lr_scheduler.OneCycleLR(cos, three_phase=False): lr used at iteration i."""
    initial, minimum = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up_end = float(pct_start * total_steps) - 1
    down_end = total_steps - 1

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)

    def fn(i):
        if i <= up_end:
            return cos(initial, max_lr, i / up_end)
        return cos(max_lr, minimum, (i - up_end) / (down_end - up_end))
    return fn


# --------------------------------------------------------------------------------------------------------------------
# synthetic batches: rows of a patch buffer with consistent geometry
# --------------------------------------------------------------------------------------------------------------------
def synth_batch(seed, b, n_images=5, f=525.0, W=640, H=480, frac_far=0.1, with_depth=False):
    """Patch-buffer rows (ace_trainer.py:330-340 layout). Features are random; geometry is consistent so that
    a fraction of rows is valid / invalid under the masks of ace_trainer.py:558-565."""
    rs = np.random.RandomState(seed)
    feats = (rs.standard_normal((b, 512)) * 0.5).astype(np.float16)
    img = rs.randint(0, n_images, size=b)
    poses_inv = np.tile(np.eye(4, dtype=np.float32), (n_images, 1, 1))
    for i in range(n_images):
        ang = rs.uniform(-0.5, 0.5, 3)
        Rx = np.array([[1, 0, 0], [0, math.cos(ang[0]), -math.sin(ang[0])], [0, math.sin(ang[0]), math.cos(ang[0])]])
        Ry = np.array([[math.cos(ang[1]), 0, math.sin(ang[1])], [0, 1, 0], [-math.sin(ang[1]), 0, math.cos(ang[1])]])
        Rz = np.array([[math.cos(ang[2]), -math.sin(ang[2]), 0], [math.sin(ang[2]), math.cos(ang[2]), 0], [0, 0, 1]])
        poses_inv[i, :3, :3] = (Rz @ Ry @ Rx).astype(np.float32)
        # scene (around the head's mean = origin) 2-4 m in front of the camera; the last image looks away from it so
        # that its rows exercise the invalid-point branch (ace_trainer.py:558, 592-600)
        poses_inv[i, :3, 3] = (rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5), rs.uniform(2, 4))
        if i == n_images - 1 and n_images > 1:
            poses_inv[i, 2, 3] = -1.0
    aug = np.tile(np.eye(4, dtype=np.float32)[:3], (b, 1, 1))
    a = rs.uniform(-0.26, 0.26, b)  # +-15 deg in-plane rotation (dataset.py aug_rotation)
    aug[:, 0, 0], aug[:, 0, 1], aug[:, 1, 0], aug[:, 1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    scale = rs.uniform(2 / 3, 3 / 2, b).astype(np.float32)
    K = np.tile(np.eye(3, dtype=np.float32), (b, 1, 1))
    K[:, 0, 0] = K[:, 1, 1] = f * scale
    K[:, 0, 2] = W * scale / 2
    K[:, 1, 2] = H * scale / 2
    Kinv = np.linalg.inv(K).astype(np.float32)
    px = np.stack([8 * (rs.randint(0, 80, b) + 0.5), 8 * (rs.randint(0, 60, b) + 0.5)], 1).astype(np.float32)
    crds = rs.uniform(-3, 3, (b, 3)).astype(np.float32)
    if not with_depth:
        crds[:] = 0
    else:
        crds[rs.uniform(size=b) < 0.3] = 0
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    return {
        "features": t(feats), "target_px": t(px), "aug_poses_inv": t(aug), "poses_inv": t(poses_inv[img]),
        "intrinsics": t(K), "intrinsics_inv": t(Kinv), "target_crds": t(crds),
        "pose_idx": t(img.astype(np.int16).reshape(-1, 1)),
    }


# --------------------------------------------------------------------------------------------------------------------
# encoder (ace_network.py:14-59)
# --------------------------------------------------------------------------------------------------------------------
ENCODER_LAYERS = [  # name, cin, cout, k, stride, pad — state-dict order of ace_encoder_pretrained.pt
    ("conv1", 1, 32, 3, 1, 1), ("conv2", 32, 64, 3, 2, 1), ("conv3", 64, 128, 3, 2, 1), ("conv4", 128, 256, 3, 2, 1),
    ("res1_conv1", 256, 256, 3, 1, 1), ("res1_conv2", 256, 256, 1, 1, 0), ("res1_conv3", 256, 256, 3, 1, 1),
    ("res2_conv1", 256, 512, 3, 1, 1), ("res2_conv2", 512, 512, 1, 1, 0), ("res2_conv3", 512, 512, 3, 1, 1),
    ("res2_skip", 256, 512, 1, 1, 0),
]


def make_encoder_state(seed):
    """Deterministic (numpy RandomState) encoder state dict with the reference's names / shapes (ace_network.py:26-39);
    He-style scale so that activations keep O(1) magnitude through the 11 layers like the pretrained weights do."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, cin, cout, k, _, _ in ENCODER_LAYERS:
        std = math.sqrt(2.0 / (cin * k * k))
        sd[name + ".weight"] = torch.from_numpy((rs.standard_normal((cout, cin, k, k)) * std).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy(rs.uniform(-0.1, 0.1, (cout,)).astype(np.float32))
    return sd


def encoder_forward(sd, image_b1hw, emulate_half=False):
    """ace_network.py:41-59. Returns [B,512,h,w]."""
    spec = {n: (s, p) for n, _, _, _, s, p in ENCODER_LAYERS}

    def conv(x, name, relu=True):
        s, p = spec[name]
        y = F.conv2d(x, _rh(sd[name + ".weight"], emulate_half), _rh(sd[name + ".bias"], emulate_half), stride=s, padding=p)
        y = _rh(y, emulate_half)
        return F.relu(y) if relu else y

    x = _rh(image_b1hw.float(), emulate_half)
    x = conv(x, "conv1"); x = conv(x, "conv2"); x = conv(x, "conv3")
    res = conv(x, "conv4")
    x = conv(res, "res1_conv1"); x = conv(x, "res1_conv2"); x = conv(x, "res1_conv3")
    res = _rh(res + x, emulate_half)
    x = conv(res, "res2_conv1"); x = conv(x, "res2_conv2"); x = conv(x, "res2_conv3")
    return _rh(conv(res, "res2_skip", relu=False) + x, emulate_half)


def synth_image(seed, h=480, w=640):
    """Normalised grayscale image (dataset.py:150-153: (g - 0.4) / 0.25), numpy-seeded."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(((rs.uniform(0, 1, (1, 1, h, w)) - 0.4) / 0.25).astype(np.float32))
