"""Stress / scale cases of BASELINE.json configs[4] (10k frames, 1M-patch buffer, RANSAC sweep 64..4096 hypotheses) at
sizes the oracles still finish in seconds."""
import numpy as np
import pytest
import torch

from oracle import ace_ref
from oracle import dsacstar_ref as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hyps", [512, 4096])
def test_dsac_hypothesis_sweep_matches_oracle(hyps):
    """Same contract as test_dsac_gpu.test_full_run_matches_oracle at the top of the sweep range."""
    from acezero_b200 import dsac
    seed = 1305
    sc, Tgt, f, px, py = D.synth_scene(seed)
    ref = D.forward_rgb(sc, hyps, 10.0, f, px, py, 100.0, 100.0, 8, seed, 16, nan_to_max=True)
    t = torch.from_numpy(np.ascontiguousarray(sc)).cuda()
    poses, inl, dbg = dsac.forward_rgb_batch(t, f, px, py, hyps, 10.0, 100.0, 100.0, 8, seed, 16, debug=True)
    torch.cuda.synchronize()
    assert np.array_equal(dbg["hyp_tries"].cpu().numpy()[0], ref["tries"])
    assert int(dbg["best"].cpu()[0]) == ref["best"]
    assert abs(int(inl[0]) - ref["inliers"]) <= max(2, ref["inliers"] // 100)
    rot, tr = D.pose_error(poses[0].cpu().numpy(), ref["pose"].astype(np.float64))
    assert rot < 0.05 and tr < 1e-3, (rot, tr)


def test_dsac_many_images_one_call_is_batch_independent():
    """4096 images in one call (the per-image RNG key makes every pose independent of the batch it ran in)."""
    from acezero_b200 import dsac
    import bench
    maps = torch.from_numpy(bench.synth_scene_maps(64, 5000)).cuda()
    big = maps.repeat(64, 1, 1, 1)                       # 4096 maps; image i and i + 64k share data but not the RNG key
    kw = dict(hyps=64, inlier_threshold=10.0, inlier_alpha=100.0, max_reproj=100.0, subsample=8, seed=2089, max_tries=16)
    p_big, n_big = dsac.forward_rgb_batch(big, 525.0, 320.0, 240.0, image_index_base=0, **kw)
    p_small, n_small = dsac.forward_rgb_batch(maps, 525.0, 320.0, 240.0, image_index_base=0, **kw)
    torch.cuda.synchronize()
    assert torch.equal(p_big[:64], p_small) and torch.equal(n_big[:64], n_small)
    assert torch.isfinite(p_big).all()


def test_training_on_a_million_row_buffer_epoch_semantics():
    """1 M-row buffer (1.26 GB): the epoch permutation / ragged-tail rule of ace_trainer.py:466-477 at scale, loss finite
    and decreasing over 200 iterations of the graph-captured step."""
    import bench
    from acezero_b200.head import HeadEngine
    from acezero_b200.trainer import TrainLoop
    dev = torch.device("cuda", 0)
    rows, b = 1_000_003, 5120                            # not a multiple of the batch: 195 full batches per epoch
    buf = bench.synth_buffer(rows, dev, 2089)
    head = HeadEngine(1, True, (0.0, 0.0, 0.0), max_rows=b, training=True, device=dev)
    head.load_state(ace_ref.make_head_state(200, 1, True))
    loop = TrainLoop(head, bench.options(b, 5000), buf, use_graph=True)
    losses = []
    assert loop.run_epoch(on_iteration=None)
    assert loop.iteration == rows // b                   # ragged tail dropped
    torch.cuda.synchronize()
    losses.append(float(head.stats[0]))
    assert np.isfinite(losses[-1])
    # one more epoch: a different permutation of the same generator stream
    assert loop.run_epoch(on_iteration=None)
    torch.cuda.synchronize()
    assert loop.iteration == 2 * (rows // b)
    assert np.isfinite(float(head.stats[0]))
