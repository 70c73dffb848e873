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
    # pending: chunks (images [b,1,H,W] on the host - pinned when the loader pins -, K [b,3,3] host, file names, dataset indices)
    # of ONE image size; results: per micro-batch (files, indices, focals, poses [n,4,4] device, inliers [n] device)
    pending, n_pending, done = [], 0, []
    stats = {"images": 0, "seconds": 0.0}
    t0 = time.time()

    def flush():
        nonlocal n_pending
        if not pending:
            return
        n = n_pending
        # the loader's chunks straight into the device batch: one asynchronous copy per chunk, no host-side cat of the images
        imgs = torch.empty((n,) + tuple(pending[0][0].shape[1:]), dtype=pending[0][0].dtype, device=device)
        o = 0
        for img, _, _, _ in pending:
            imgs[o:o + img.shape[0]].copy_(img, non_blocking=True)
            o += img.shape[0]
        K = torch.cat([p[1] for p in pending], 0).float()                 # host, n x 3 x 3
        assert torch.allclose(K[:, 0, 0], K[:, 1, 1]), "a single focal length is supported (register_mapping.py:219)"
        # focal / principal point from the HOST copy of K: one small upload per micro-batch, no device read-back per image
        cam = torch.stack([K[:, 0, 0], K[:, 0, 2], K[:, 1, 2]], 0).contiguous().to(device, non_blocking=True)
        with torch.no_grad():
            sc = network(imgs).float().contiguous()                       # [n,3,h,w] stays on the device
        # ONE DSAC* launch pair per micro-batch in any image order (the loader of register_mapping.py:147 shuffles):
        # the RNG of image j is keyed by its dataset index through the per-image key array of the C ABI
        files = [f for p in pending for f in p[2]]
        idx = [i for p in pending for i in p[3]]
        poses, inl = dsac.forward_rgb_batch(sc, cam[0], cam[1], cam[2], hypotheses, threshold, inlier_alpha, max_pixel_error,
                                            network.OUTPUT_SUBSAMPLE, base_seed, max_tries, image_index=idx)
        done.append((files, idx, [float(x) for x in K[:, 0, 0]], poses, inl))
        pending.clear()
        n_pending = 0

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
        ind = [int(x) for x in indices] if torch.is_tensor(indices) else [int(indices)]
        names = [filenames] * B if isinstance(filenames, str) else list(filenames)
        keep = [b for b in range(B) if ind[b] % world_size == rank]
        if 0 < max_estimates and count + len(keep) > max_estimates:
            keep = keep[:max(0, max_estimates - count)]
        if keep:
            if len(keep) < B:
                sel = torch.tensor(keep)
                image, K = image[sel], K[sel]
            if pending and (pending[0][0].shape[1:] != image.shape[1:] or n_pending + len(keep) > micro_batch):
                flush()
            pending.append((image, K, [names[b] for b in keep], [ind[b] for b in keep]))
            n_pending += len(keep)
            count += len(keep)
        if 0 < max_estimates <= count:
            break
    flush()
    results = []
    if done:
        # ONE device->host transfer for all poses / inlier counts at the end
        poses = torch.cat([d[3] for d in done], 0).cpu().numpy()
        inl = torch.cat([d[4] for d in done], 0).cpu().numpy()
        k = 0
        for files, idx, focals, _, _ in done:
            for j in range(len(idx)):
                results.append({"file": files[j], "index": idx[j], "pose": poses[k], "inliers": int(inl[k]), "focal": focals[j]})
                k += 1
    stats["images"] = len(results)
    stats["seconds"] = time.time() - t0
    return results, stats
