"""Host logic of the registration loop (acezero_b200/registration.py, reference register_mapping.py:201-258) on the CPU: the
network and the DSAC* launch are stand-ins, what is checked is the loop itself - micro-batching of loader chunks of one image size,
both loader forms (default collate and collate_same_size), the rank filter, max_estimates, the per-image RNG keys handed to the
solver, the camera parameters taken from K, and the result records."""
import numpy as np
import pytest
import torch

from acezero_b200 import registration


class _Net:
    OUTPUT_SUBSAMPLE = 8

    def __init__(self):
        self.batches = []

    def __call__(self, imgs):
        self.batches.append(tuple(imgs.shape))
        n, _, h, w = imgs.shape
        # a recognisable map: every cell carries the image's mean value
        return imgs.float().mean(dim=(1, 2, 3)).view(n, 1, 1, 1).expand(n, 3, h // 8, w // 8).contiguous()


def _item(i, h=64, w=96, focal=500.0):
    img = torch.full((1, h, w), float(i))
    K = torch.tensor([[focal + i, 0.0, w / 2], [0.0, focal + i, h / 2], [0.0, 0.0, 1.0]])
    z = torch.zeros(1)
    return img, z, z, z, K, z, z, f"frame_{i:03d}.png", i


def _default_batches(ids, bs, **kw):
    from torch.utils.data.dataloader import default_collate
    return [default_collate([_item(i, **kw) for i in ids[o:o + bs]]) for o in range(0, len(ids), bs)]


@pytest.fixture
def fake_solver(monkeypatch):
    calls = []

    def solve(sc, focal, ppx, ppy, hyps, thr, alpha, maxerr, sub, seed, tries, image_index=None, **_):
        n = sc.shape[0]
        calls.append({"n": n, "keys": list(image_index), "focal": focal.clone(), "ppx": ppx.clone(), "ppy": ppy.clone()})
        poses = torch.eye(4).repeat(n, 1, 1)
        poses[:, 0, 3] = sc[:, 0, 0, 0]                      # the image's value travels through the 'pose'
        return poses, torch.tensor(list(image_index), dtype=torch.int32) + 1000
    monkeypatch.setattr(registration.dsac, "forward_rgb_batch", solve)
    return calls


def test_micro_batches_keys_and_records(fake_solver):
    ids = [5, 2, 9, 0, 7, 1, 8, 3, 6, 4, 10]             # a shuffled loader
    net = _Net()
    res, stats = registration.register(net, _default_batches(ids, 4), micro_batch=8, device="cpu")
    assert [r["index"] for r in res] == ids and stats["images"] == len(ids)
    assert [c["n"] for c in fake_solver] == [8, 3]        # chunks of 4 + 4, then the ragged rest
    assert [k for c in fake_solver for k in c["keys"]] == ids          # RNG key of an image = its dataset index, any order
    for r in res:
        assert r["file"] == f"frame_{r['index']:03d}.png" and r["inliers"] == r["index"] + 1000
        assert r["pose"].shape == (4, 4) and r["pose"][0, 3] == float(r["index"])
        assert r["focal"] == pytest.approx(500.0 + r["index"])
    c = fake_solver[0]
    np.testing.assert_allclose(c["focal"].numpy(), [500.0 + i for i in ids[:8]])
    np.testing.assert_allclose(c["ppx"].numpy(), [48.0] * 8)
    np.testing.assert_allclose(c["ppy"].numpy(), [32.0] * 8)


def test_rank_filter_and_max_estimates(fake_solver):
    ids = list(range(12))
    parts = []
    for rank in range(3):
        res, _ = registration.register(_Net(), _default_batches(ids, 5), rank=rank, world_size=3, device="cpu")
        assert all(r["index"] % 3 == rank for r in res)
        parts += [r["index"] for r in res]
    assert sorted(parts) == ids                            # every image solved by exactly one rank
    res, _ = registration.register(_Net(), _default_batches(ids, 5), max_estimates=7, device="cpu")
    assert [r["index"] for r in res] == ids[:7]


def test_mixed_sizes_and_collate_same_size(fake_solver):
    items = [_item(0), _item(1, h=80, w=64), _item(2), _item(3, h=80, w=64), _item(4)]
    loader = [registration.collate_same_size(items[:3]), registration.collate_same_size(items[3:])]
    net = _Net()
    res, _ = registration.register(net, loader, micro_batch=16, device="cpu")
    assert sorted(r["index"] for r in res) == [0, 1, 2, 3, 4]
    # a micro-batch never mixes image sizes
    assert all(len({b[2:] for b in [shape]}) == 1 for shape in net.batches)
    sizes = {(64, 96): set(), (80, 64): set()}
    for r in res:
        sizes[(64, 96) if r["index"] in (0, 2, 4) else (80, 64)].add(r["index"])
    assert sizes == {(64, 96): {0, 2, 4}, (80, 64): {1, 3}}
    for shape in net.batches:
        assert shape[2:] in ((64, 96), (80, 64))


def test_single_image_loader_form(fake_solver):
    """The reference's loader form: batch_size 1, default collate (register_mapping.py:147)."""
    res, _ = registration.register(_Net(), _default_batches([3, 1, 2], 1), micro_batch=2, device="cpu")
    assert [r["index"] for r in res] == [3, 1, 2]
    assert [c["n"] for c in fake_solver] == [2, 1]
