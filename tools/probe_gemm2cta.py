"""GPU probe of the EXPERIMENTAL cta_group::2 GEMM (csrc/gemm2cta.cu) against torch.matmul and, for timing, against the
cta_group::1 kernel on the weight-gradient shape (8 x 512 x 512 x 5120, MN-major operands).

    python tools/probe_gemm2cta.py            # correctness cases, then timing
Run each case in the order printed: the first ones are the smallest (one cluster, one k-block).
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from acezero_b200 import _lib  # noqa: E402
from tools.probe_gemm import ref  # noqa: E402


def run(fn_name, A, B, mn, M, N, K, batch=1, bn=128, bias=None):
    lib = _lib.load()
    out = torch.full((batch, M, N), float("nan"), device="cuda", dtype=torch.float32)
    d = _lib.GemmDesc()
    if bias is not None:
        d.bias_grad, d.bias_grad_zstride = bias.data_ptr(), M
    d.A, d.B = A.data_ptr(), B.data_ptr()
    d.a_mn_major = d.b_mn_major = mn
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_zstride = A.stride(0) if batch > 1 else 0
    d.b_zstride = B.stride(0) if batch > 1 else 0
    d.lda, d.ldb = A.stride(-2), B.stride(-2)
    d.bn = bn
    d.epilogue = 2
    d.out32, d.out32_zstride, d.ldo32 = out.data_ptr(), M * N, N
    _lib.check(getattr(lib, fn_name)(C.byref(d), _lib.stream_ptr()), fn_name)
    return out, d


def case(name, mn, M, N, K, batch=1, bn=0, bias_grad=False):
    g = torch.Generator(device="cuda").manual_seed(1)
    shpA = (K, M) if mn else (M, K)
    shpB = (K, N) if mn else (N, K)
    if batch > 1:
        shpA, shpB = (batch,) + shpA, (batch,) + shpB
    A = (torch.randn(shpA, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(shpB, device="cuda", generator=g) * 0.5).half()
    try:
        bg = torch.full((batch, M), float("nan"), device="cuda") if bias_grad else None
        out, _ = run("acez_gemm2cta_f16", A, B, mn, M, N, K, batch, bn, bg)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"{name}: ERROR {e}", flush=True)
        return False
    r = ref(A, B, mn, mn).reshape(batch, M, N)
    err = (out - r).abs()
    scale = r.abs().max().item()
    ok = bool(err.max().item() <= 2e-3 * max(scale, 1.0)) and bool(torch.isfinite(out).all())
    # where is it wrong? per 128-row half (CTA of the pair) and per 128-column half (B half of the pair)
    loc = ""
    if not ok:
        e0 = err[0]
        loc = " | max err by (row half, col half) of the first 256x256 tile: " + " ".join(
            f"{e0[128 * i:128 * i + 128, 128 * j:128 * j + 128].nan_to_num(9e9).max().item():.3g}" for i in range(2) for j in range(2))
    if bias_grad:
        Af = (A.float().transpose(-1, -2) if mn else A.float()).reshape(batch, M, K)
        e2 = (bg - Af.sum(-1)).abs().nan_to_num(9e9).max().item()
        ok = ok and e2 < 1e-2
        loc += f" bias_grad_err={e2:.4g}"
    print(f"{name}: max_abs_err={err.nan_to_num(9e9).max().item():.4g} (ref max {scale:.4g}) {'OK' if ok else 'FAIL'}{loc}", flush=True)
    return ok


def timing():
    L, C5, rows = 8, 512, 5120
    g = torch.Generator(device="cuda").manual_seed(2)
    A = (torch.randn((L, rows, C5), device="cuda", generator=g) * 0.1).half()   # DZ  [K=rows][M]
    B = (torch.randn((L, rows, C5), device="cuda", generator=g) * 0.1).half()   # ACT [K=rows][N]
    for name, fn, bn in (("cta_group::1 128x128", "acez_gemm_f16", 128), ("cta_group::2 256x256 per pair (64 CTAs)", "acez_gemm2cta_f16", 0),
                         ("cta_group::2 256x128 per pair (128 CTAs)", "acez_gemm2cta_f16", 128)):
        for _ in range(3):
            run(fn, A, B, 1, C5, C5, rows, L, bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            run(fn, A, B, 1, C5, C5, rows, L, bn)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1000
        print(f"wgrad shape, {name}: {us:.1f} us = {2 * L * C5 * C5 * rows / us / 1e6:.0f} TFLOP/s", flush=True)
        if fn == "acez_gemm2cta_f16" and os.environ.get("ACEZ_GEMM2_DBG", "0") == "1":
            import numpy as np
            n_cta = 2 * (C5 // 256) * (C5 // (bn or 256)) * L
            buf = np.zeros(n_cta * 8, dtype=np.int64)
            _lib.check(_lib.load().acez_debug_gemm2_clocks(buf.ctypes.data_as(C.c_void_p), n_cta))
            b = buf.reshape(n_cta, 8)
            lead = b[b[:, 1] > 0]
            print(f"   MMA warp (leaders): loop {np.median(lead[:, 1]):.0f} cycles, of which waiting for operands {np.median(lead[:, 0]):.0f} "
                  f"(max {lead[:, 0].max()}); producers: loop {np.median(b[:, 3]):.0f}, waiting for free stages {np.median(b[:, 2]):.0f}; "
                  f"epilogue {np.median(b[:, 4]):.0f}", flush=True)


def timing_kmajor():
    """The same FLOPs with K-major operands (rows contiguous): is the MN-major operand layout what holds the weight-gradient
    GEMM back?"""
    L, C5, rows = 8, 512, 5120
    g = torch.Generator(device="cuda").manual_seed(3)
    A = (torch.randn((L, C5, rows), device="cuda", generator=g) * 0.1).half()
    B = (torch.randn((L, C5, rows), device="cuda", generator=g) * 0.1).half()
    for bn in (128, 0):
        for _ in range(3):
            run("acez_gemm2cta_f16", A, B, 0, C5, C5, rows, L, bn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            run("acez_gemm2cta_f16", A, B, 0, C5, C5, rows, L, bn)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1000
        print(f"same shape, K-major operands, 256x{bn or 256} tiles: {us:.1f} us = {2 * L * C5 * C5 * rows / us / 1e6:.0f} TFLOP/s", flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    ok = True
    ok &= case("K/K   256x256x64   (one pair, one k-block)", 0, 256, 256, 64)
    ok &= case("K/K   256x256x512", 0, 256, 256, 512)
    ok &= case("K/K   512x512x512  (4 pairs)", 0, 512, 512, 512)
    ok &= case("K/K   ragged M=300 N=320", 0, 300, 320, 128)
    ok &= case("MN/MN 256x256x64", 1, 256, 256, 64)
    ok &= case("MN/MN 512x512x5120", 1, 512, 512, 5120)
    ok &= case("MN/MN batched x8 K=640", 1, 512, 512, 640, batch=8)
    ok &= case("K/K   512x512x512  256x128 tiles", 0, 512, 512, 512, bn=128)
    ok &= case("MN/MN batched x8 K=640, 256x128 tiles", 1, 512, 512, 640, batch=8, bn=128)
    ok &= case("MN/MN batched x8 K=5120, 256x128 tiles + bias column (the weight-gradient launch)", 1, 512, 512, 5120, batch=8, bn=128,
               bias_grad=True)
    ok &= case("MN/MN batched x2 K=640, 256x256 tiles + bias column", 1, 512, 512, 640, batch=2, bn=256, bias_grad=True)
    print("RESULT", "PASS" if ok else "FAIL", flush=True)
    if ok:
        timing()
        timing_kmajor()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
