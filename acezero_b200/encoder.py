"""Host side of the ACE encoder (C ABI `acez_encoder_*`): frozen weights, NHWC fp16 features.
Mirrors `ace_network.Encoder` (reference ace_network.py:14-59)."""
import ctypes as C

import torch

from . import _lib

ENCODER_KEYS = ["conv1", "conv2", "conv3", "conv4", "res1_conv1", "res1_conv2", "res1_conv3", "res2_conv1",
                "res2_conv2", "res2_conv3", "res2_skip"]
ENCODER_SHAPES = {"conv1": (32, 1, 3, 3), "conv2": (64, 32, 3, 3), "conv3": (128, 64, 3, 3), "conv4": (256, 128, 3, 3),
                  "res1_conv1": (256, 256, 3, 3), "res1_conv2": (256, 256, 1, 1), "res1_conv3": (256, 256, 3, 3),
                  "res2_conv1": (512, 256, 3, 3), "res2_conv2": (512, 512, 1, 1), "res2_conv3": (512, 512, 3, 3),
                  "res2_skip": (512, 256, 1, 1)}


def out_hw(H, W):
    d = lambda n: (n - 1) // 2 + 1
    return d(d(d(H))), d(d(d(W)))


class EncoderEngine:
    def __init__(self, state_dict, max_n=1, max_h=480, max_w=640, device="cuda"):
        self.lib = _lib.load()
        _lib.check(self.lib.acez_device_check(), "acez_device_check")
        self.device = torch.device(device)
        if state_dict["res2_conv3.weight"].shape[0] != 512:
            raise ValueError("only the 512-d encoder of ace_encoder_pretrained.pt is supported")
        self.weights = []
        for k in ENCODER_KEYS:
            w = state_dict[k + ".weight"]
            if tuple(w.shape) != ENCODER_SHAPES[k]:
                raise ValueError(f"encoder weight {k} has shape {tuple(w.shape)}, expected {ENCODER_SHAPES[k]}")
            self.weights.append(w.detach().to(self.device, torch.float32).contiguous())
            self.weights.append(state_dict[k + ".bias"].detach().to(self.device, torch.float32).contiguous())
        self.plan = None
        self._cap = (0, 0, 0)
        self._ensure(max_n, max_h, max_w)

    def _ensure(self, n, h, w):
        cn, ch, cw = self._cap
        if self.plan is not None and n <= cn and h * w <= ch * cw and h <= ch and w <= cw:
            return
        n, h, w = max(n, cn), max(h, ch), max(w, cw)
        if self.plan is not None:
            torch.cuda.synchronize()
            self.lib.acez_encoder_plan_destroy(self.plan)
            self.plan = None
        ws_bytes = int(self.lib.acez_encoder_workspace_bytes(n, h, w))
        self.workspace = torch.empty(ws_bytes, device=self.device, dtype=torch.uint8)
        ptrs = (C.c_void_p * 22)(*[t.data_ptr() for t in self.weights])
        plan = C.c_void_p()
        rc = self.lib.acez_encoder_plan_create(ptrs, n, h, w, _lib.ptr(self.workspace), ws_bytes, _lib.stream_ptr(),
                                               C.byref(plan))
        _lib.check(rc, "acez_encoder_plan_create")
        self.plan = plan
        self._cap = (n, h, w)

    def __del__(self):
        try:
            if self.plan is not None:
                self.lib.acez_encoder_plan_destroy(self.plan)
        except Exception:
            pass

    def forward_nhwc(self, image_b1hw, out=None, stream=None):
        """image: CUDA [n,1,H,W] fp16/fp32. Returns NHWC fp16 features [n, h/8, w/8, 512]."""
        if image_b1hw.dim() != 4 or image_b1hw.shape[1] != 1:
            raise ValueError("expected a [n,1,H,W] image tensor")
        img = image_b1hw.contiguous()
        if img.dtype not in (torch.float16, torch.float32):
            img = img.float()
        n, _, H, W = img.shape
        self._ensure(n, H, W)
        h8, w8 = out_hw(H, W)
        if out is None:
            out = torch.empty((n, h8, w8, 512), device=self.device, dtype=torch.float16)
        rc = self.lib.acez_encoder_forward(self.plan, _lib.ptr(img), int(img.dtype == torch.float16), n, H, W,
                                           _lib.ptr(out), _lib.stream_ptr(stream))
        _lib.check(rc, "acez_encoder_forward")
        return out
