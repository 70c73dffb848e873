"""GPU probe of the tcgen05 GEMM: every operand-major combination against torch.matmul (fp32 reference on fp16 data).

Run on the GPU box: `python tools/probe_gemm.py`. Prints one line per case; exits non-zero if a default case fails.
For MN-major operands it can also sweep alternative (LBO, SBO, k-step) descriptor constants (--sweep).
"""
import ctypes as C
import itertools
import sys

import torch

sys.path.insert(0, ".")
from acezero_b200 import _lib  # noqa: E402


def run_f32(A, B, a_mn, b_mn, M, N, K, batch=1, bias_grad=False, **ov):
    lib = _lib.load()
    out = torch.full((batch, M, N), float("nan"), device="cuda", dtype=torch.float32)
    bg = torch.full((batch, M), float("nan"), device="cuda", dtype=torch.float32) if bias_grad else None
    d = _lib.GemmDesc()
    d.A, d.B = A.data_ptr(), B.data_ptr()
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_zstride = A.stride(0) if batch > 1 else 0
    d.b_zstride = B.stride(0) if batch > 1 else 0
    d.lda = A.stride(-2)
    d.ldb = B.stride(-2)
    d.bn = 128
    d.epilogue = 2
    d.out32 = out.data_ptr()
    d.out32_zstride = M * N
    d.ldo32 = N
    if bg is not None:
        d.bias_grad = bg.data_ptr()
        d.bias_grad_zstride = M
    for k, v in ov.items():
        setattr(d, k, v)
    _lib.check(lib.acez_gemm_f16(C.byref(d), _lib.stream_ptr()), "acez_gemm_f16")
    torch.cuda.synchronize()
    return out, bg


def ref(A, B, a_mn, b_mn):
    Af = A.float().transpose(-1, -2) if a_mn else A.float()   # -> [.., M, K]
    Bf = B.float() if b_mn else B.float().transpose(-1, -2)   # -> [.., K, N]
    return Af @ Bf


def case(name, a_mn, b_mn, M, N, K, batch=1, bias_grad=False, **ov):
    g = torch.Generator(device="cuda").manual_seed(1)
    shpA = (K, M) if a_mn else (M, K)
    shpB = (K, N) if b_mn else (N, K)
    if batch > 1:
        shpA, shpB = (batch,) + shpA, (batch,) + shpB
    A = (torch.randn(shpA, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(shpB, device="cuda", generator=g) * 0.5).half()
    try:
        out, bg = run_f32(A, B, a_mn, b_mn, M, N, K, batch, bias_grad, **ov)
    except Exception as e:  # noqa: BLE001
        print(f"{name}: ERROR {e}")
        return False
    r = ref(A, B, a_mn, b_mn).reshape(batch, M, N)
    err = (out - r).abs().max().item()
    scale = r.abs().max().item()
    ok = bool(err <= 2e-3 * max(scale, 1.0)) and bool(torch.isfinite(out).all())
    msg = f"{name}: max_abs_err={err:.4g} (ref max {scale:.4g}) {'OK' if ok else 'FAIL'}"
    if bias_grad:
        Af = (A.float().transpose(-1, -2) if a_mn else A.float()).reshape(batch, M, K)
        e2 = (bg - Af.sum(-1)).abs().max().item()
        ok = ok and e2 < 1e-2
        msg += f" bias_grad_err={e2:.4g}"
    print(msg, flush=True)
    return ok


def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    ok = True
    ok &= case("K/K   128x128x64", 0, 0, 128, 128, 64)
    ok &= case("K/K   256x512x512", 0, 0, 256, 512, 512)
    ok &= case("K/K   5120x512x512", 0, 0, 5120, 512, 512)
    ok &= case("K/K   ragged M=200", 0, 0, 200, 128, 128)
    mn_ok = case("K/MN  256x512x512", 0, 1, 256, 512, 512)
    mn_ok &= case("MN/K  256x512x512", 1, 0, 256, 512, 512)
    mn_ok &= case("MN/MN 512x512x5120 (+bias col)", 1, 1, 512, 512, 5120, bias_grad=True)
    mn_ok &= case("MN/MN batched x3 K=640", 1, 1, 512, 512, 640, batch=3, bias_grad=True)
    if not mn_ok and "--sweep" in sys.argv:
        print("sweeping MN-major descriptor constants on K/MN 256x512x512")
        for lbo, sbo, ks in itertools.product([8192, 1024, 128, 64 * 128 * 2], [1024, 8192, 128], [2048, 32, 1024, 4096]):
            case(f"  b_lbo={lbo} b_sbo={sbo} b_kstep={ks}", 0, 1, 256, 512, 512, b_lbo=lbo, b_sbo=sbo, b_kstep=ks)
    print("RESULT", "PASS" if (ok and mn_ok) else "FAIL")
    sys.exit(0 if (ok and mn_ok) else 1)


if __name__ == "__main__":
    main()
