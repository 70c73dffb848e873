"""CPU-only checks: the C-ABI library builds, loads and exports every symbol of include/acez.h; argument validation
and the no-device error path; host-compilable pieces of the DSAC* solver (RNG, quartic, P3P, Rodrigues) against
numpy / cv2."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib):
    from acezero_b200 import _lib
    header = (ROOT / "include" / "acez.h").read_text()
    declared = set(re.findall(r"\b(acez_[a-z0-9_]+)\s*\(", header))
    declared -= {"acez_stream_t"}
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"libacez.so lacks {missing}"
    assert declared == set(_lib.EXPORTS)
    assert lib.acez_version() == 100


def test_compute_entry_fails_loudly_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from acezero_b200 import _lib
    assert lib.acez_device_check() == 4
    d = _lib.GemmDesc()
    assert lib.acez_gemm_f16(C.byref(d), None) != 0
    with pytest.raises(_lib.AcezError):
        from acezero_b200.head import HeadEngine
        HeadEngine()


def test_head_param_count_and_workspace(lib):
    from acezero_b200 import _lib
    cfg = _lib.HeadConfig()
    cfg.num_res_blocks, cfg.use_homogeneous, cfg.max_rows, cfg.training = 2, 1, 5120, 1
    # 8 x (512*512 + 512) + 4*512 + 4 = 2 103 300
    assert lib.acez_head_param_count(C.byref(cfg)) == 2103300
    ws = lib.acez_head_workspace_bytes(C.byref(cfg))
    act = (9 + 2 + 8 + 1) * 5120 * 512 * 2 + 9 * 5120 * 64   # ACT[L+1], XTRA[nres], DZ[L], GRES; MASKB[L+1] (bits)
    # + fp16 weight shadow + fc3 output-gradient rows + fc3 slab partials (160 x 2052 floats) + bias-gradient partials of the
    # 2-CTA weight-gradient GEMM (8 layers x 2 x 4 x 256 floats) + tail block partials
    assert act < ws < act + 8 * 512 * 512 * 2 + 5120 * 16 + 160 * 2052 * 4 + 8 * 2 * 4 * 256 * 4 + 4096 * 32 + 64 * 1024
    cfg.num_res_blocks = 0
    assert lib.acez_head_param_count(C.byref(cfg)) == 0


def test_rng_matches_oracle(lib):
    from oracle import dsacstar_ref as D
    xy = (C.c_int * 2)()
    for seed, img, hyp, tr, j, w, h in [(1305, 0, 0, 0, 0, 80, 60), (2 ** 40 + 7, 9999, 4095, 999999, 3, 107, 60),
                                          (0, 1, 2, 3, 1, 120, 90)]:
        lib.acez_host_draw_cell(C.c_uint64(seed), img, hyp, tr, j, w, h, xy)
        assert (xy[0], xy[1]) == D.draw_cell(seed, img, hyp, tr, j, w, h)
        assert 0 <= xy[0] < w and 0 <= xy[1] < h


def test_quartic_solver(lib):
    rs = np.random.RandomState(0)
    f = lib.acez_host_solve_quartic
    for _ in range(200):
        r = np.sort(rs.uniform(-3, 3, 4))
        c = np.poly(r)[1:]
        roots = np.zeros(4)
        n = f(np.ascontiguousarray(c).ctypes.data_as(C.c_void_p), roots.ctypes.data_as(C.c_void_p))
        assert n == 4
        np.testing.assert_allclose(np.sort(roots), r, atol=1e-6)
    # two real + two complex roots
    c = np.poly([1.5, -0.5, 0.3 + 1j, 0.3 - 1j])[1:].real
    roots = np.zeros(4)
    n = f(np.ascontiguousarray(c).ctypes.data_as(C.c_void_p), roots.ctypes.data_as(C.c_void_p))
    assert n == 2
    np.testing.assert_allclose(np.sort(roots[:2]), [-0.5, 1.5], atol=1e-8)


def test_p3p_contains_ground_truth_and_matches_cv2(lib):
    import cv2
    rs = np.random.RandomState(1)
    f, cx, cy = 525.0, 320.0, 240.0
    cam = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]])
    for _ in range(50):
        rv = rs.uniform(-0.5, 0.5, 3)
        R, _ = cv2.Rodrigues(rv)
        t = np.array([rs.uniform(-1, 1), rs.uniform(-1, 1), rs.uniform(3, 5)])
        Pw = rs.uniform(-1, 1, (4, 3))
        Pc = Pw @ R.T + t
        img = Pc[:, :2] / Pc[:, 2:] * f + [cx, cy]
        b = np.concatenate([(img[:3] - [cx, cy]) / f, np.ones((3, 1))], 1)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        Rs, ts = np.zeros(36), np.zeros(12)
        n = lib.acez_host_p3p(np.ascontiguousarray(Pw[:3]).ctypes.data_as(C.c_void_p),
                              np.ascontiguousarray(b).ctypes.data_as(C.c_void_p), Rs.ctypes.data_as(C.c_void_p),
                              ts.ctypes.data_as(C.c_void_p))
        assert 1 <= n <= 4
        Rs, ts = Rs.reshape(4, 3, 3)[:n], ts.reshape(4, 3)[:n]
        # every returned solution reprojects the 3 points exactly and is a rotation
        for k in range(n):
            pc = Pw[:3] @ Rs[k].T + ts[k]
            np.testing.assert_allclose(pc[:, :2] / pc[:, 2:] * f + [cx, cy], img[:3], atol=1e-6)
            np.testing.assert_allclose(Rs[k] @ Rs[k].T, np.eye(3), atol=1e-9)
        # the ground truth is among them, and picking by the 4th point gives cv2.solvePnP(P3P)'s answer
        errs = [np.linalg.norm((Pw[3] @ Rs[k].T + ts[k])[:2] / (Pw[3] @ Rs[k].T + ts[k])[2] * f + [cx, cy] - img[3])
                for k in range(n)]
        k = int(np.argmin(errs))
        np.testing.assert_allclose(Rs[k], R, atol=1e-7)
        np.testing.assert_allclose(ts[k], t, atol=1e-6)
        ok, rv2, tv2 = cv2.solvePnP(Pw.reshape(-1, 1, 3), img.reshape(-1, 1, 2), cam, None, flags=cv2.SOLVEPNP_P3P)
        assert ok
        np.testing.assert_allclose(cv2.Rodrigues(rv2)[0], Rs[k], atol=1e-6)
        np.testing.assert_allclose(tv2.ravel(), ts[k], atol=1e-5)


def test_rodrigues_and_jacobian(lib):
    import cv2
    rs = np.random.RandomState(2)
    for r in [rs.uniform(-2, 2, 3) for _ in range(20)] + [np.zeros(3), np.array([1e-9, 0, 0])]:
        R, dR = np.zeros(9), np.zeros(27)
        lib.acez_host_rodrigues(np.ascontiguousarray(r).ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p),
                                dR.ctypes.data_as(C.c_void_p))
        Rcv, Jcv = cv2.Rodrigues(r.reshape(3, 1))
        np.testing.assert_allclose(R.reshape(3, 3), Rcv, atol=1e-12)
        np.testing.assert_allclose(dR.reshape(3, 9), Jcv, atol=1e-9)  # cv2: 3x9, row i = dR/dr_i (row-major R)
        back = np.zeros(3)
        lib.acez_host_rodrigues_inv(R.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
        np.testing.assert_allclose(back, cv2.Rodrigues(Rcv)[0].ravel(), atol=1e-9)


def test_dsac_oracle_recovers_pose():
    from oracle import dsacstar_ref as D
    sc, Tgt, f, px, py = D.synth_scene(1305)
    r = D.forward_rgb(sc, 32, 10.0, f, px, py, 100.0, 100.0, 8, 1305, 16, nan_to_max=True)
    rot, tr = D.pose_error(r["pose"], Tgt)
    assert r["inliers"] > 2500 and rot < 0.5 and tr < 0.02


def test_chain_protocol_model_check():
    """The mbarrier protocol of the fused layer-chain kernel (csrc/head_chain4.cu: cluster of four, cta_group::2) under random
    interleavings of its agents and asynchronous engines: no stale / aliased phase, no box written under a reader, no
    TMEM buffer overwritten before it is drained, no deadlock (tools/sim_chain4_protocol.py)."""
    import importlib.util
    spec4 = importlib.util.spec_from_file_location("sim_chain4_protocol", ROOT / "tools" / "sim_chain4_protocol.py")
    sim4 = importlib.util.module_from_spec(spec4)
    spec4.loader.exec_module(sim4)
    for seed in range(24):
        for n in (1, 2, 3, 8):
            sim4.Sim(n, seed).run()


def test_dp_shard_partition(lib):
    """Shards of the peer-memory optimiser (csrc/adamw_dp.cu): multiples of 8 parameters (one 16-byte fp16 store per group and
    peer), G of them cover the 2 103 300 head parameters, the layer strides keep every group inside one weight / bias block."""
    n = 2103300
    for g in (1, 2, 4, 8):
        s = lib.acez_adamw_dp_shard(n, g)
        assert s % 8 == 0 and g * s >= n and (g - 1) * s < n
    assert (512 * 512 + 512) % 8 == 0 and (512 * 512) % 8 == 0


def test_dp_optimizer_protocol_model_check():
    """The cross-GPU protocol of the one-kernel data-parallel optimiser step (csrc/adamw_dp.cu: epoch signals, verdict exchange,
    weight pushes, fences) under random interleavings and delivery orders of the remote stores: gradients are read only while they
    belong to the iteration, weights are never overwritten under a computing owner and are complete before the next forward, no
    signal wait passes on a stale epoch, the verdict is global, no deadlock (tools/sim_dp_protocol.py). The checker has teeth:
    without the final wait for everybody's "weights written" signal it must find the stale-weights violation."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sim_dp_protocol", ROOT / "tools" / "sim_dp_protocol.py")
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for G in (2, 4):
        for seed in range(12):
            sim.Sim(G, 3, 3, seed).run()
    src = (ROOT / "tools" / "sim_dp_protocol.py").read_text()
    broken = src.replace('            for q in range(G):\n                yield from self.wait_row(me, "applied", q, e)\n', "")
    assert broken != src
    ns = {}
    exec(compile(broken, "sim_dp_protocol_broken", "exec"), ns)
    with pytest.raises(AssertionError):
        for seed in range(8):
            ns["Sim"](3, 3, 3, seed).run()


def test_header_is_plain_c():
    """include/acez.h is the FFI boundary: it must compile as C99 (plain pointers and sizes, no C++ / torch types) and as C++."""
    import shutil
    import subprocess
    hdr = str(ROOT / "include" / "acez.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    for cmd in (["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", hdr],
                ["g++", "-x", "c++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
