"""GPU parity of the stand-alone fused reprojection-loss kernel and the generic tcgen05 GEMM entry."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ace_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loss_type,use_depth,rows", [("dyntanh", False, 5120), ("tanh", False, 1000), ("l1", False, 777),
                                                      ("l1+sqrt", True, 2048), ("l1+log", False, 2048)])
def test_repro_loss_kernel_matches_autograd(lib, loss_type, use_depth, rows):
    """fp32 kernel vs autograd through the oracle's restatement of ace_trainer.py:521-613 on the same fp32 inputs:
    loss 1e-5 relative, gradients 1e-4 relative to the largest gradient entry."""
    from acezero_b200 import _lib
    from acezero_b200.head import LOSS_TYPES
    it, S = 100, 128.0
    bt = ace_ref.synth_batch(500, rows, with_depth=use_depth)
    rs = np.random.RandomState(9)
    sc = torch.from_numpy(rs.uniform(-1.5, 1.5, (rows, 3)).astype(np.float32)).requires_grad_(True)
    opts = ace_ref.LossOptions(repro_loss_type=loss_type, use_depth=use_depth, iterations=1000)
    P = torch.bmm(bt["aug_poses_inv"], bt["poses_inv"]).requires_grad_(True)
    loss, inl, n_valid = ace_ref.training_loss(opts, sc, bt["target_px"], None, None, bt["intrinsics"],
                                               bt["intrinsics_inv"], bt["target_crds"], it, P_b34=P)
    (loss * S).backward()
    w = ace_ref.loss_weight(opts, it)
    lp = _lib.LossParams(LOSS_TYPES[loss_type], w, 0.1, 1000.0, 1000.0, 10.0, 10.0, int(use_depth), S, rows)
    d = {k: v.cuda() for k, v in bt.items()}
    scd = sc.detach().cuda()
    d_sc = torch.empty_like(scd)
    d_P = torch.empty((rows, 3, 4), device="cuda")
    stats = torch.zeros(4, device="cuda")
    for compose in (True, False):
        stats.zero_()
        Pd = None if compose else P.detach().cuda()
        rc = lib.acez_repro_loss_fwd_bwd(C.byref(lp), rows, _lib.ptr(scd), _lib.ptr(d["target_px"]), _lib.ptr(Pd),
                                         _lib.ptr(d["aug_poses_inv"]), _lib.ptr(d["poses_inv"]),
                                         _lib.ptr(d["intrinsics"]), _lib.ptr(d["intrinsics_inv"]),
                                         _lib.ptr(d["target_crds"]), _lib.ptr(d_sc), _lib.ptr(d_P), None,
                                         _lib.ptr(stats), _lib.stream_ptr())
        _lib.check(rc, "acez_repro_loss_fwd_bwd")
        st = stats.cpu().numpy()
        assert st[2] == n_valid and st[1] == round(inl * rows) and st[3] == 0
        assert abs(st[0] - float(loss)) < 1e-5 * abs(float(loss)) + 1e-6
        g = sc.grad.numpy()
        np.testing.assert_allclose(d_sc.cpu().numpy(), g, rtol=1e-4, atol=1e-4 * np.abs(g).max())
        gp = P.grad.numpy()
        np.testing.assert_allclose(d_P.cpu().numpy(), gp, rtol=1e-4, atol=1e-4 * np.abs(gp).max())


def test_repro_loss_argument_errors(lib):
    from acezero_b200 import _lib
    lp = _lib.LossParams(0, 50.0, 0.1, 1000.0, 1000.0, 10.0, 10.0, 0, 1.0, 0)
    rc = lib.acez_repro_loss_fwd_bwd(C.byref(lp), 4, None, None, None, None, None, None, None, None, None, None, None,
                                     None, None)
    assert rc == 1 and b"null" in lib.acez_last_error()


def _gemm(lib, A, B, a_mn, b_mn, M, N, K, **kw):
    from acezero_b200 import _lib
    d = _lib.GemmDesc()
    d.A, d.B = A.data_ptr(), B.data_ptr()
    d.a_mn_major, d.b_mn_major, d.M, d.N, d.K, d.batch = a_mn, b_mn, M, N, K, 1
    d.lda, d.ldb = A.stride(0), B.stride(0)
    for k, v in kw.items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    _lib.check(lib.acez_gemm_f16(C.byref(d), _lib.stream_ptr()), "acez_gemm_f16")
    torch.cuda.synchronize()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (200, 128, 192), (512, 512, 5120)])
def test_gemm_f32_all_operand_majors(lib, a_mn, b_mn, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g) * 0.5).half()
    B = (torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g) * 0.5).half()
    out = torch.full((M, N), float("nan"), device="cuda")
    bg = torch.full((M,), float("nan"), device="cuda")
    _gemm(lib, A, B, a_mn, b_mn, M, N, K, bn=128, epilogue=2, out32=out, ldo32=N, bias_grad=bg)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float() if b_mn else B.float().t()
    ref = Af @ Bf
    # fp32 accumulation of exact fp16 products: only the summation order differs
    assert (out - ref).abs().max() < 2e-3 * ref.abs().max()
    assert (bg - Af.sum(1)).abs().max() < 1e-2


def test_gemm_forward_epilogue_bias_relu_residual(lib):
    M, N, K = 300, 512, 512
    g = torch.Generator(device="cuda").manual_seed(2)
    A = (torch.randn((M, K), device="cuda", generator=g) * 0.5).half()
    W = (torch.randn((N, K), device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = (torch.randn((M, N), device="cuda", generator=g)).half()
    out = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    out2 = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    for bn in (128, 256):
        _gemm(lib, A, W, 0, 0, M, N, K, bn=bn, epilogue=0, bias=bias, resid=res, out=out, out2=out2, ldo=N, relu=1)
        x = torch.relu((A.float() @ W.float().t() + bias.half().float())).half()
        # one fp16 ulp where the fp32 sum straddles a rounding boundary
        assert (out.float() - x.float()).abs().max() <= 2e-3 * x.float().abs().max()
        assert (out2.float() - (res.float() + out.float()).half().float()).abs().max() == 0


def test_gemm_dgrad_epilogue_mask_addend_nonfinite(lib):
    M, N, K = 256, 512, 512
    g = torch.Generator(device="cuda").manual_seed(3)
    dZ = (torch.randn((M, K), device="cuda", generator=g)).half()
    W = (torch.randn((K, N), device="cuda", generator=g) * 0.05).half()   # [out=K rows, in=N]: MN-major B
    mask = torch.relu(torch.randn((M, N), device="cuda", generator=g)).half()
    add = torch.randn((M, N), device="cuda", generator=g).half()
    out = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    raw = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    _gemm(lib, dZ, W, 0, 1, M, N, K, bn=256, epilogue=1, mask=mask, addend=add, out=out, out2=raw, ldo=N, nonfinite=flag)
    v = ((dZ.float() @ W.float()).half() + add).float()
    assert (raw.float() - v).abs().max() <= 4e-3 * v.abs().max()
    assert torch.equal(out, torch.where(mask > 0, raw, torch.zeros_like(raw)))
    assert int(flag) == 0
    dZ[3, 5] = 60000.0
    W[5, :] = 100.0
    _gemm(lib, dZ, W, 0, 1, M, N, K, bn=256, epilogue=1, mask=mask, out=out, out2=raw, ldo=N, nonfinite=flag)
    assert int(flag) == 1 and not torch.isnan(out).any()
