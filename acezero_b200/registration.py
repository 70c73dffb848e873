"""Registration loop on the GPU (reference register_mapping.py:201-258): encoder + head + DSAC* per image, with the
scene coordinates staying on the device (the reference copies them to the CPU and runs RANSAC there, GPU idle).
Images are grouped into micro-batches of equal size; poses are solved by one batched DSAC* launch per micro-batch and
read back once at the end. With world_size > 1, image i is processed by rank i % world_size; the RNG of each image is
keyed by its dataset index, so the result does not depend on the number of ranks.
"""
import time

import torch

from . import dsac


def collate_same_size(items):
    """DataLoader collate for registration with batch_size > 1 (extension; the reference loads one image per step,
    register_mapping.py:147): images of equal size are stacked into one 9-tuple batch, a mixed batch falls apart into
    single-image batches. Returns a LIST of 9-tuple batches; `register` accepts both forms."""
    from torch.utils.data.dataloader import default_collate
    groups = {}
    for it in items:
        groups.setdefault(tuple(it[0].shape), []).append(it)
    return [default_collate(g) for g in groups.values()]


def register(network, loader, hypotheses=64, threshold=10.0, inlier_alpha=100.0, max_pixel_error=100.0,
             base_seed=1305, max_tries=1000000, max_estimates=-1, micro_batch=16, rank=0, world_size=1,
             device="cuda"):
    """Returns a list of dicts {file, index, pose (4x4 cam->world numpy), inliers, focal} in processing order."""
    pending, results = [], []
    stats = {"images": 0, "seconds": 0.0}
    t0 = time.time()

    def flush():
        if not pending:
            return
        # image by image from the loader's (pinned) tensors straight into the device batch: asynchronous copies, no host-side cat
        imgs = torch.empty((len(pending),) + tuple(pending[0][0].shape[1:]), dtype=pending[0][0].dtype, device=device)
        for k, p in enumerate(pending):
            imgs[k].copy_(p[0][0], non_blocking=True)
        K = torch.stack([p[1] for p in pending], 0)
        f = K[:, 0, 0].contiguous()
        with torch.no_grad():
            sc = network(imgs).float().contiguous()                   # [n,3,h,w] stays on the device
        # ONE DSAC* launch pair per micro-batch in any image order (the loader of register_mapping.py:147 shuffles):
        # the RNG of image j is keyed by its dataset index through the per-image key array of the C ABI
        idx = [int(p[3]) for p in pending]
        poses, inl = dsac.forward_rgb_batch(sc, f, K[:, 0, 2].contiguous(), K[:, 1, 2].contiguous(), hypotheses,
                                            threshold, inlier_alpha, max_pixel_error, network.OUTPUT_SUBSAMPLE,
                                            base_seed, max_tries, image_index=idx)
        for j, p in enumerate(pending):
            results.append({"file": p[2], "index": int(p[3]), "pose": poses[j], "inliers": inl[j], "focal": p[4]})
        pending.clear()

    def batches():
        for item in loader:
            if isinstance(item, list) and item and isinstance(item[0], (list, tuple)):   # collate_same_size: list of batches
                for sub in item:
                    yield sub
            else:
                yield item

    count = 0
    for image, _, _, _, K, _, _, filenames, indices in batches():
        B = image.shape[0]
        for b in range(B):
            i = int(indices[b]) if torch.is_tensor(indices) else int(indices)
            if i % world_size != rank:
                continue
            Kb = K[b]
            assert torch.allclose(Kb[0, 0], Kb[1, 1]), "a single focal length is supported (register_mapping.py:219)"
            # focal from the CPU copy of K (no device read-back per image)
            focal = float(Kb[0, 0])
            item = (image[b:b + 1], Kb.to(device, non_blocking=True),
                    filenames[b] if not isinstance(filenames, str) else filenames, i, focal)
            if pending and (pending[0][0].shape != item[0].shape or len(pending) >= micro_batch):
                flush()
            pending.append(item)
            count += 1
        if 0 < max_estimates <= count:
            break
    flush()
    torch.cuda.synchronize()
    for r in results:   # one device->host transfer at the end
        r["pose"] = r["pose"].cpu().numpy()
        r["inliers"] = int(r["inliers"])
    stats["images"] = len(results)
    stats["seconds"] = time.time() - t0
    return results, stats
