"""Drop-in for the reference's `ace_network` module (reference ace_network.py): the same classes, constructor
arguments, factory methods and `state_dict()` keys / shapes, with the compute running in the sm_100a kernels of
libacez.so (tcgen05 encoder convolutions, tcgen05 head GEMM chain + fused tail) instead of cuDNN/cuBLAS.

There is no PyTorch compute fallback: a forward pass on a non-CUDA tensor raises.
"""
import logging
import math
import re

import torch
import torch.nn as nn

_logger = logging.getLogger(__name__)


class Encoder(nn.Module):
    """FCN encoder (reference ace_network.py:14-59). Parameters live in ordinary `nn.Conv2d` containers so that
    `state_dict()` / `load_state_dict()` match `ace_encoder_pretrained.pt`; `forward` runs the CUDA plan."""

    def __init__(self, out_channels=512):
        super().__init__()
        self.out_channels = out_channels
        self.conv1 = nn.Conv2d(1, 32, 3, 1, 1)
        self.conv2 = nn.Conv2d(32, 64, 3, 2, 1)
        self.conv3 = nn.Conv2d(64, 128, 3, 2, 1)
        self.conv4 = nn.Conv2d(128, 256, 3, 2, 1)
        self.res1_conv1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.res1_conv2 = nn.Conv2d(256, 256, 1, 1, 0)
        self.res1_conv3 = nn.Conv2d(256, 256, 3, 1, 1)
        self.res2_conv1 = nn.Conv2d(256, 512, 3, 1, 1)
        self.res2_conv2 = nn.Conv2d(512, 512, 1, 1, 0)
        self.res2_conv3 = nn.Conv2d(512, self.out_channels, 3, 1, 1)
        self.res2_skip = nn.Conv2d(256, self.out_channels, 1, 1, 0)
        self._engine = None
        self._engine_version = None

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self):
        """The CUDA plan over the current (frozen) weights; rebuilt if they were replaced or modified."""
        from acezero_b200.encoder import EncoderEngine
        v = self._weights_version()
        if self._engine is None or v != self._engine_version:
            self._engine = EncoderEngine(self.state_dict(), device=next(self.parameters()).device)
            self._engine_version = v
        return self._engine

    def forward_nhwc(self, x):
        """[B,1,H,W] -> NHWC fp16 [B,H/8,W/8,C]: the kernels' native layout (rows of the patch buffer)."""
        if not x.is_cuda:
            raise RuntimeError("ace_network.Encoder runs on CUDA (sm_100a) only; move the module and input to the GPU")
        return self.engine().forward_nhwc(x)

    def forward(self, x):
        # BCHW view of the NHWC result (channels_last memory format), fp16 like the reference under autocast
        return self.forward_nhwc(x).permute(0, 3, 1, 2)


class _HeadFunction(torch.autograd.Function):
    """Autograd bridge so that a PyTorch training loop (e.g. the reference's ace_trainer.py:516-518,627) can
    differentiate through `Head.forward`: parameter gradients are produced by the CUDA backward into the flat gradient
    buffer and handed to autograd as views."""

    @staticmethod
    def forward(ctx, head, rows_f16, *params):
        eng = head._bound_engine()
        rows = rows_f16.shape[0]
        if rows > eng.max_rows:
            eng.resize(rows)
        sc = torch.empty((rows, 3), device=rows_f16.device, dtype=torch.float32)
        from acezero_b200 import _lib
        eng.sync_weights()
        rc = eng.lib.acez_head_forward_train(eng.plan, _lib.ptr(rows_f16), rows, _lib.ptr(sc), _lib.stream_ptr())
        _lib.check(rc, "acez_head_forward_train")
        ctx.head, ctx.rows = head, rows
        return sc

    @staticmethod
    def backward(ctx, d_sc):
        from acezero_b200 import _lib
        eng = ctx.head._bound_engine()
        d_sc = d_sc.contiguous().float()
        rc = eng.lib.acez_head_backward(eng.plan, ctx.rows, _lib.ptr(d_sc), _lib.ptr(eng.found_inf), _lib.stream_ptr())
        _lib.check(rc, "acez_head_backward")
        gv = eng.grad_views()
        grads = []
        for name in ctx.head._param_names():
            grads.append(gv[name].clone())
        return (None, None) + tuple(grads)


class Head(nn.Module):
    """MLP head (reference ace_network.py:62-149): same constructor, sub-module names and buffers."""

    def __init__(self, mean, num_head_blocks, use_homogeneous, homogeneous_min_scale=0.01, homogeneous_max_scale=4.0,
                 in_channels=512):
        super().__init__()
        self.use_homogeneous = use_homogeneous
        self.in_channels = in_channels
        self.head_channels = 512
        if in_channels != self.head_channels:
            raise NotImplementedError("the sm_100a head supports the 512-d encoder only (head_skip = Identity, "
                                      "reference ace_network.py:81)")
        self.head_skip = nn.Identity()
        self.num_head_blocks = num_head_blocks
        self.res3_conv1 = nn.Conv2d(512, 512, 1, 1, 0)
        self.res3_conv2 = nn.Conv2d(512, 512, 1, 1, 0)
        self.res3_conv3 = nn.Conv2d(512, 512, 1, 1, 0)
        self.res_blocks = []
        for block in range(num_head_blocks):
            blk = (nn.Conv2d(512, 512, 1, 1, 0), nn.Conv2d(512, 512, 1, 1, 0), nn.Conv2d(512, 512, 1, 1, 0))
            self.res_blocks.append(blk)
            self.add_module(str(block) + "c0", blk[0])
            self.add_module(str(block) + "c1", blk[1])
            self.add_module(str(block) + "c2", blk[2])
        self.fc1 = nn.Conv2d(512, 512, 1, 1, 0)
        self.fc2 = nn.Conv2d(512, 512, 1, 1, 0)
        if self.use_homogeneous:
            self.fc3 = nn.Conv2d(512, 4, 1, 1, 0)
            self.register_buffer("max_scale", torch.tensor([homogeneous_max_scale]))
            self.register_buffer("min_scale", torch.tensor([homogeneous_min_scale]))
            self.register_buffer("max_inv_scale", 1. / self.max_scale)
            self.register_buffer("h_beta", math.log(2) / (1. - self.max_inv_scale))
            self.register_buffer("min_inv_scale", 1. / self.min_scale)
        else:
            self.fc3 = nn.Conv2d(512, 3, 1, 1, 0)
        self.register_buffer("mean", mean.clone().detach().view(1, 3, 1, 1))
        self._engine = None
        self._engine_version = None

    # ---- engine plumbing -------------------------------------------------------------------------------------
    def _layer_names(self):
        from acezero_b200.head import head_layer_names
        return head_layer_names(self.num_head_blocks) + ["fc3"]

    def _param_names(self):
        return [n + s for n in self._layer_names() for s in (".weight", ".bias")]

    def _params(self):
        mods = dict(self.named_modules())
        return [getattr(mods[n], s) for n in self._layer_names() for s in ("weight", "bias")]

    def _homogeneous_buffers(self):
        return (self.h_beta, self.max_inv_scale, self.min_inv_scale) if self.use_homogeneous else ()

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self._params()) + (self.mean.data_ptr(), self.mean._version) + \
            tuple((b.data_ptr(), b._version) for b in self._homogeneous_buffers())

    def engine(self, training=False, max_rows=5120, peer_group=None):
        """HeadEngine holding a copy of the current weights (rebuilt / reloaded when the module's tensors change).
        peer_group: torch.distributed group of the GPUs of this box -> parameters / gradient / workspace in symmetric memory
        (data-parallel optimiser over NVLink peer memory, acezero_b200/csrc/adamw_dp.cu)."""
        from acezero_b200.head import HeadEngine
        dev = self.mean.device
        if dev.type != "cuda":
            raise RuntimeError("ace_network.Head runs on CUDA (sm_100a) only; move the module to the GPU")
        hom = tuple(float(b) for b in self._homogeneous_buffers())
        if self._engine is None or (training and not self._engine.training) or \
                (hom and hom != (self._engine.h_beta, self._engine.max_inv_scale, self._engine.min_inv_scale)):
            # the engine reads the de-homogenisation constants from the module's BUFFERS, as the reference's forward does
            # (ace_network.py:139-144) — for an fp16 head file these are the fp16-rounded values stored in it
            kw = dict(h_beta=hom[0], max_inv_scale=hom[1], min_inv_scale=hom[2]) if hom else {}
            self._engine = HeadEngine(self.num_head_blocks, self.use_homogeneous, self.mean.reshape(3).cpu(),
                                      max_rows=max_rows, training=training, device=dev, peer_group=peer_group, **kw)
            self._engine_version = None
        v = self._weights_version()
        if v != self._engine_version:
            sd = {k: t for k, t in self.state_dict().items()}
            self._engine.load_state(sd)
            self._engine_version = v
        return self._engine

    def _bound_engine(self):
        return self.engine(training=True)

    def export_engine_weights(self):
        """Copy the engine's (trained) flat parameters back into the module's tensors."""
        if self._engine is None:
            return
        v = self._engine.views()
        mods = dict(self.named_modules())
        with torch.no_grad():
            for n in self._layer_names():
                mods[n].weight.copy_(v[n + ".weight"])
                mods[n].bias.copy_(v[n + ".bias"])
        self._engine_version = self._weights_version()

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward_rows(self, rows_f16):
        """[rows,512] fp16 -> [rows,3] fp32 scene coordinates (no autograd)."""
        return self.engine().forward(rows_f16)

    def forward(self, res):
        """res: [B,512,H,W] -> [B,3,H,W] (reference ace_network.py:120-149)."""
        B, C, H, W = res.shape
        rows = res.permute(0, 2, 3, 1).reshape(-1, C)
        if rows.dtype != torch.float16:
            rows = rows.half()
        rows = rows.contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params()):
            sc = _HeadFunction.apply(self, rows, *self._params())
        else:
            sc = self.forward_rows(rows)
        return sc.view(B, H, W, 3).permute(0, 3, 1, 2)


class Regressor(nn.Module):
    """FCN architecture for scene coordinate regression (reference ace_network.py:152-270)."""

    OUTPUT_SUBSAMPLE = 8

    def __init__(self, mean, num_head_blocks, use_homogeneous, num_encoder_features=512):
        super().__init__()
        self.feature_dim = num_encoder_features
        self.encoder = Encoder(out_channels=self.feature_dim)
        self.heads = Head(mean, num_head_blocks, use_homogeneous, in_channels=self.feature_dim)

    @classmethod
    def create_from_encoder(cls, encoder_state_dict, mean, num_head_blocks, use_homogeneous):
        num_encoder_features = encoder_state_dict['res2_conv3.weight'].shape[0]
        _logger.info(f"Creating Regressor using pretrained encoder with {num_encoder_features} feature size.")
        regressor = cls(mean, num_head_blocks, use_homogeneous, num_encoder_features)
        regressor.encoder.load_state_dict(encoder_state_dict)
        return regressor

    @classmethod
    def create_from_state_dict(cls, state_dict):
        mean = torch.zeros((3,))
        pattern = re.compile(r"^heads\.\d+c0\.weight$")
        num_head_blocks = sum(1 for k in state_dict.keys() if pattern.match(k))
        use_homogeneous = state_dict["heads.fc3.weight"].shape[0] == 4
        num_encoder_features = state_dict['encoder.res2_conv3.weight'].shape[0]
        _logger.info(f"Creating regressor from pretrained state_dict:"
                     f"\n\tNum head blocks: {num_head_blocks}"
                     f"\n\tHomogeneous coordinates: {use_homogeneous}"
                     f"\n\tEncoder feature size: {num_encoder_features}")
        regressor = cls(mean, num_head_blocks, use_homogeneous, num_encoder_features)
        regressor.load_state_dict(state_dict)
        return regressor

    @classmethod
    def create_from_split_state_dict(cls, encoder_state_dict, head_state_dict):
        merged_state_dict = {}
        for k, v in encoder_state_dict.items():
            merged_state_dict[f"encoder.{k}"] = v
        for k, v in head_state_dict.items():
            merged_state_dict[f"heads.{k}"] = v
        return cls.create_from_state_dict(merged_state_dict)

    def load_encoder(self, encoder_dict_file):
        self.encoder.load_state_dict(torch.load(encoder_dict_file))

    def get_features(self, inputs):
        return self.encoder(inputs)

    def get_scene_coordinates(self, features):
        return self.heads(features)

    def forward(self, inputs):
        # fused path: keep the encoder's NHWC rows, skip the BCHW round trip
        f = self.encoder.forward_nhwc(inputs)
        B, H, W, C = f.shape
        sc = self.heads.forward_rows(f.view(-1, C)) if not (torch.is_grad_enabled() and any(
            p.requires_grad for p in self.heads._params())) else None
        if sc is None:
            return self.heads(f.permute(0, 3, 1, 2))
        return sc.view(B, H, W, 3).permute(0, 3, 1, 2)
