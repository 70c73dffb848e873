"""Host-side wrapper of the CUDA DSAC* solver (C ABI `acez_dsac_forward_rgb_batch`).

`forward_rgb_batch` is the batched, device-resident entry the new register_mapping.py uses; `forward_rgb` mirrors
the reference's native operator `dsacstar.forward_rgb` (reference dsacstar/dsacstar.cpp:66-186, binding :898-899,
call site register_mapping.py:229-242) positionally.
"""
import ctypes as C

import torch

from . import _lib

MAX_REF_STEPS = 100  # reference dsacstar.cpp:47


def _as_dev_f32(v, n, device):
    if torch.is_tensor(v):
        t = v.to(device=device, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and n > 1:
            t = t.expand(n)
        return t.contiguous()
    return torch.full((n,), float(v), device=device, dtype=torch.float32)


def forward_rgb_batch(scene_coordinates, focal, ppx, ppy, hyps=64, inlier_threshold=10.0, inlier_alpha=100.0,
                      max_reproj=100.0, subsample=8, seed=0, max_tries=1000000, injected_idx=None,
                      image_index_base=0, max_refine_steps=MAX_REF_STEPS, debug=False, stream=None, image_index=None):
    """scene_coordinates: CUDA float32 [n,3,h,w]. Returns (poses [n,4,4] float32 cam->world, inliers [n] int32)
    on the device (+ a dict of per-hypothesis intermediates when debug=True). No host synchronisation.

    The sampling RNG of image i is keyed by `image_index[i]` (sequence / tensor of n ints, any order — e.g. the dataset
    indices of a shuffled micro-batch) or, without it, by `image_index_base + i`."""
    lib = _lib.load()
    sc = scene_coordinates
    if not (torch.is_tensor(sc) and sc.is_cuda and sc.dim() == 4 and sc.shape[1] == 3):
        raise ValueError("scene_coordinates must be a CUDA tensor of shape [n,3,h,w]")
    if sc.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {sc.dtype}")  # the reference's accessor<float,4> error
    sc = sc.contiguous()
    n, _, h, w = sc.shape
    dev = sc.device
    f_t, px_t, py_t = (_as_dev_f32(v, n, dev) for v in (focal, ppx, ppy))
    poses = torch.empty((n, 4, 4), device=dev, dtype=torch.float32)
    inliers = torch.empty((n,), device=dev, dtype=torch.int32)
    ws_bytes = lib.acez_dsac_workspace_bytes(n, h, w, hyps)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    idx_t = None
    if image_index is not None:
        idx_t = image_index if torch.is_tensor(image_index) else torch.as_tensor(list(image_index), dtype=torch.int32)
        idx_t = idx_t.to(device=dev, dtype=torch.int32, non_blocking=True).reshape(-1).contiguous()
        if idx_t.numel() != n:
            raise ValueError(f"image_index must have {n} entries")
    p = _lib.DsacParams(hyps, inlier_threshold, inlier_alpha, max_reproj, subsample, int(seed) & 0xFFFFFFFFFFFFFFFF,
                        int(min(max_tries, 2**31 - 1)), max_refine_steps, image_index_base,
                        idx_t.data_ptr() if idx_t is not None else None)
    inj = None
    if injected_idx is not None:
        inj = torch.as_tensor(injected_idx, dtype=torch.int32).to(dev).contiguous()
        if tuple(inj.shape) != (n, hyps, 4, 2):
            raise ValueError(f"injected_idx must have shape {(n, hyps, 4, 2)}")
    dbg = None
    out = {}
    if debug:
        out = {
            "hyp_poses": torch.empty((n, hyps, 6), device=dev, dtype=torch.float32),
            "hyp_scores": torch.empty((n, hyps), device=dev, dtype=torch.float32),
            "best": torch.empty((n,), device=dev, dtype=torch.int32),
            "hyp_tries": torch.empty((n, hyps), device=dev, dtype=torch.int32),
            "refine_rounds": torch.empty((n,), device=dev, dtype=torch.int32),
        }
        dbg = _lib.DsacDebug(out["hyp_poses"].data_ptr(), out["hyp_scores"].data_ptr(), out["best"].data_ptr(),
                             out["hyp_tries"].data_ptr(), out["refine_rounds"].data_ptr())
    rc = lib.acez_dsac_forward_rgb_batch(_lib.ptr(sc), n, h, w, _lib.ptr(f_t), _lib.ptr(px_t), _lib.ptr(py_t),
                                         C.byref(p), _lib.ptr(inj), _lib.ptr(poses), _lib.ptr(inliers),
                                         C.byref(dbg) if dbg is not None else None, _lib.ptr(ws), ws_bytes,
                                         _lib.stream_ptr(stream))
    _lib.check(rc, "acez_dsac_forward_rgb_batch")
    if debug:
        return poses, inliers, out
    return poses, inliers


def content_key(sc):
    """31-bit RNG key that is a pure function of a scene-coordinate map's bits (device tensor [n], no host sync)."""
    n = sc.shape[0]
    return (sc.contiguous().view(torch.int32).reshape(n, -1).sum(dim=1, dtype=torch.int64) & 0x7FFFFFFF).to(torch.int32)


def forward_rgb(sceneCoordinates, outPose, ransacHypotheses, inlierThreshold, focalLength, ppointX, ppointY,
                inlierAlpha, maxReproj, subSampling, randomSeed, max_hypotheses_tries, image_index=None):
    """Drop-in for the reference's `dsacstar.forward_rgb` (all-positional call at register_mapping.py:229-242).

    sceneCoordinates: [1,3,H,W] float32, CPU (as the reference passes it) or CUDA; outPose: [4,4] float32 written in
    place (camera->world); returns the inlier count as a Python int. Silent on stdout.

    The reference applies its seed only on the first call of a process (thread_rand.cpp:17) and then continues one RNG
    stream, so its result depends on the call history. Here the result is a pure function of the arguments: the
    sampling RNG is keyed by (randomSeed, image key), the key being `image_index` when the caller passes it (extension,
    keyword only) and otherwise a checksum of the scene-coordinate bits — the same image and seed give the same pose
    in any call order, in any process, on any number of GPUs.
    """
    if sceneCoordinates.dim() != 4 or sceneCoordinates.shape[0] != 1:
        raise RuntimeError("sceneCoordinates must be 1x3xHxW (only batch size 1 is supported by this entry point)")
    sc = sceneCoordinates if sceneCoordinates.is_cuda else sceneCoordinates.cuda(non_blocking=True)
    if sc.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {sc.dtype}")
    key = content_key(sc) if image_index is None else [int(image_index)]
    poses, inl = forward_rgb_batch(sc, float(focalLength), float(ppointX), float(ppointY), int(ransacHypotheses),
                                   float(inlierThreshold), float(inlierAlpha), float(maxReproj), int(subSampling),
                                   int(randomSeed), int(max_hypotheses_tries), image_index=key)
    outPose.copy_(poses[0].to(outPose.device))
    return int(inl.item())
