"""CPU oracle for the DSAC* pose solver — TEST INFRASTRUCTURE ONLY.

A restatement of the reference's native operator `dsacstar.forward_rgb`
(reference dsacstar/dsacstar.cpp:66-186 and dsacstar/dsacstar_util.h) in numpy + cv2. Like the reference it
delegates the PnP arithmetic to OpenCV (`cv2.solvePnP` P3P / ITERATIVE, `cv2.projectPoints`, `cv2.Rodrigues`;
reference call sites dsacstar_util.h:104-112,199-205,395-401,578-580,762). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product path never does.

PARITY STATUS: *unpinned by the reference* — the reference ships no tests, golden vectors or fixtures for this
operator and its C++ cannot be compiled here (needs the OpenCV C++ SDK, dsacstar/setup.py:35). This oracle is pinned
only to OpenCV's own behaviour (version cv2.__version__ here, 4.4.0 in the reference's environment.yml:115).

Deliberate difference from the reference: minimal-set sampling uses the counter-based RNG of the CUDA kernel
(`draw_cell`, splitmix64 keyed by seed/image/hypothesis/try) instead of `std::mt19937` per OpenMP thread
(thread_rand.cpp:13-42), whose stream depends on thread count, libstdc++ version and call history.
"""
import numpy as np

try:
    import cv2
except Exception as e:  # pragma: no cover
    cv2 = None
    _cv2_error = e

MAX_REF_STEPS = 100  # dsacstar.cpp:47
EPS = 0.00000001     # dsacstar_util.h:45

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    """splitmix64 finaliser on python ints (mod 2^64) — mirrors acez::splitmix64 in csrc/dsac.cu."""
    z = (z + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def draw_cell(seed, image, hyp, tr, j, w, h):
    """Cell (x, y) of draw j — mirrors acez::draw_cell (replaces irand(0,imW), irand(0,imH), dsacstar_util.h:171-172)."""
    s = splitmix64(seed & 0xFFFFFFFFFFFFFFFF)
    s = splitmix64(s ^ (image & 0xFFFFFFFF))
    s = splitmix64(s ^ (hyp & 0xFFFFFFFF))
    s = splitmix64(s ^ (tr & 0xFFFFFFFF))
    r = splitmix64((s + j) & 0xFFFFFFFFFFFFFFFF)
    x = ((r & 0xFFFFFFFF) * w) >> 32
    y = ((r >> 32) * h) >> 32
    return int(x), int(y)


def create_sampling(w, h, sub):
    """dsacstar_util.h:59-76: pixel position of each cell centre (ints)."""
    xs = np.arange(w) * sub + sub // 2
    ys = np.arange(h) * sub + sub // 2
    return xs, ys


def _cam_mat(f, ppx, ppy):
    # dsacstar.cpp:91-95 (float 3x3)
    return np.array([[f, 0, ppx], [0, f, ppy], [0, 0, 1]], dtype=np.float32)


def _safe_solve_pnp(obj, img, cam, rvec, tvec, guess, flag):
    """dsacstar_util.h:91-120."""
    try:
        if guess:
            ok, r, t = cv2.solvePnP(obj, img, cam, None, rvec.copy(), tvec.copy(), True, flag)
        else:
            ok, r, t = cv2.solvePnP(obj, img, cam, None, flags=flag)
    except cv2.error:
        ok = False
    if not ok:
        return False, np.zeros((3, 1)), np.zeros((3, 1))
    return True, np.asarray(r, dtype=np.float64).reshape(3, 1), np.asarray(t, dtype=np.float64).reshape(3, 1)


def get_repro_errs(sc_3hw, rvec, tvec, cam, xs, ys, max_reproj, nan_to_max=False):
    """dsacstar_util.h:356-446 (calcJ = false branch): float error map [h, w].

    nan_to_max=False is the reference: std::min((float) norm, maxReproj) returns NaN for a NaN norm, the NaN reaches
    the score, poisons softMax() and makes draw() return hypothesis 0 (dsacstar_util.h:441,684-752).
    nan_to_max=True is the CUDA kernel's documented divergence: a non-finite error counts as maxReproj."""
    _, h, w = sc_3hw.shape
    pts3 = sc_3hw.reshape(3, -1).T.astype(np.float32)
    gx, gy = np.meshgrid(xs.astype(np.float32), ys.astype(np.float32))
    pts2 = np.stack([gx.ravel(), gy.ravel()], axis=1)
    proj, _ = cv2.projectPoints(pts3.reshape(-1, 1, 3), rvec, tvec, cam, None)
    proj = proj.reshape(-1, 2).astype(np.float32)
    d = (pts2 - proj).astype(np.float32)
    n = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2).astype(np.float32)
    with np.errstate(invalid="ignore"):
        n = np.minimum(n, np.float32(max_reproj))  # std::min((float) norm, maxReproj): NaN propagates
    if nan_to_max:
        n = np.where(np.isfinite(n), n, np.float32(max_reproj)).astype(np.float32)
    return n.reshape(h, w)


def get_hyp_score(errs, thr, alpha):
    """dsacstar_util.h:316-343 for one hypothesis."""
    beta = np.float32(5) / np.float32(thr)
    soft = (beta * (errs.astype(np.float32) - np.float32(thr))).astype(np.float64)
    soft = 1.0 / (1.0 + np.exp(-soft))
    score = np.sum(1.0 - soft)
    h, w = errs.shape
    return score * float(np.float32(alpha) / np.float32(w) / np.float32(h))


def softmax(scores):
    """dsacstar_util.h:684-704."""
    s = np.asarray(scores, dtype=np.float64)
    m = s[0]
    for v in s:
        if v > m:
            m = v
    e = np.exp(s - m)
    return e / e.sum()


def draw_argmax(probs):
    """dsacstar_util.h:727-752 with training = false: first maximum among probs >= EPS."""
    max_prob, max_idx = -1.0, 0
    for i, p in enumerate(probs):
        if p < EPS:
            continue
        if max_prob < 0 or p > max_prob:
            max_prob, max_idx = p, i
    return max_idx


def pose2trans(rvec, tvec):
    """dsacstar_util.h:759-770."""
    R, _ = cv2.Rodrigues(rvec)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = tvec.ravel()
    return np.linalg.inv(T)


def sample_hypothesis(sc_3hw, xs, ys, cam, thr, seed, image, hyp, max_tries, injected=None):
    """dsacstar_util.h:157-220 for one hypothesis. Returns (rvec, tvec, tries, ok)."""
    _, h, w = sc_3hw.shape
    rvec = np.zeros((3, 1))
    tvec = np.zeros((3, 1))
    tries = 1 if injected is not None else max(1, max_tries)
    for t in range(tries):
        obj = np.zeros((4, 3), dtype=np.float32)
        img = np.zeros((4, 2), dtype=np.float32)
        for j in range(4):
            if injected is not None:
                x, y = int(injected[j][0]), int(injected[j][1])
            else:
                x, y = draw_cell(seed, image, hyp, t, j, w, h)
            obj[j] = sc_3hw[:, y, x]
            img[j] = (xs[x], ys[y])
        ok, rvec, tvec = _safe_solve_pnp(obj.reshape(-1, 1, 3), img.reshape(-1, 1, 2), cam, None, None, False,
                                         cv2.SOLVEPNP_P3P)
        if not ok:
            continue
        proj, _ = cv2.projectPoints(obj.reshape(-1, 1, 3), rvec, tvec, cam, None)
        proj = proj.reshape(-1, 2).astype(np.float32)
        d = (img - proj).astype(np.float32)
        n = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
        if np.all(n < thr):
            return rvec, tvec, t + 1, True
    return rvec, tvec, tries, False


def refine_hyp(sc_3hw, errs, xs, ys, cam, thr, max_steps, max_reproj, rvec, tvec, nan_to_max=False):
    """dsacstar_util.h:522-597. Returns (rvec, tvec, inlier_map or None, accepted rounds)."""
    _, h, w = sc_3hw.shape
    local = errs.copy()
    best = 4
    inlier_map = None
    rounds = 0
    gx, gy = np.meshgrid(xs.astype(np.float32), ys.astype(np.float32))
    for _ in range(max_steps):
        mask = local < np.float32(thr)
        n = int(mask.sum())
        if n <= best:
            break
        best = n
        # reference collects x-outer / y-inner; the fit is order independent up to fp summation
        mt = mask.T
        img = np.stack([gx.T[mt], gy.T[mt]], axis=1).astype(np.float32)
        obj = np.stack([sc_3hw[0].T[mt], sc_3hw[1].T[mt], sc_3hw[2].T[mt]], axis=1).astype(np.float32)
        flag = cv2.SOLVEPNP_ITERATIVE if n > 4 else cv2.SOLVEPNP_P3P
        ok, r2, t2 = _safe_solve_pnp(obj.reshape(-1, 1, 3), img.reshape(-1, 1, 2), cam, rvec, tvec, True, flag)
        if not ok:
            break
        rvec, tvec = r2, t2
        inlier_map = mask.astype(np.int32)
        rounds += 1
        local = get_repro_errs(sc_3hw, rvec, tvec, cam, xs, ys, max_reproj, nan_to_max)
    return rvec, tvec, inlier_map, rounds


def forward_rgb(sc_13hw, hyps, thr, f, ppx, ppy, alpha, max_reproj, subsample, seed, max_tries, image_index=0,
                injected=None, max_ref_steps=MAX_REF_STEPS, nan_to_max=False):
    """dsacstar.cpp:66-186. sc_13hw: float32 [1,3,H,W]. Returns a dict with the camera->world pose, the inlier
    count and the per-hypothesis intermediates the parity tests compare."""
    if cv2 is None:  # pragma: no cover
        raise RuntimeError(f"cv2 unavailable: {_cv2_error}")
    sc = np.asarray(sc_13hw, dtype=np.float32)[0]
    _, h, w = sc.shape
    cam = _cam_mat(f, ppx, ppy)
    xs, ys = create_sampling(w, h, subsample)
    rvecs, tvecs, tries, oks = [], [], [], []
    for k in range(hyps):
        inj = None if injected is None else injected[k]
        r, t, n, ok = sample_hypothesis(sc, xs, ys, cam, thr, seed, image_index, k, max_tries, inj)
        rvecs.append(r); tvecs.append(t); tries.append(n); oks.append(ok)
    errs = [get_repro_errs(sc, rvecs[k], tvecs[k], cam, xs, ys, max_reproj, nan_to_max) for k in range(hyps)]
    scores = np.array([get_hyp_score(e, thr, alpha) for e in errs])
    probs = softmax(scores)
    best = draw_argmax(probs)
    r, t, inl, rounds = refine_hyp(sc, errs[best], xs, ys, cam, thr, max_ref_steps, max_reproj, rvecs[best],
                                   tvecs[best], nan_to_max)
    pose = pose2trans(r, t).astype(np.float32)
    return {
        "pose": pose,
        "inliers": 0 if inl is None else int(inl.sum()),
        "inlier_map": inl,
        "best": int(best),
        "scores": scores,
        "hyp_rvecs": np.array([x.ravel() for x in rvecs]),
        "hyp_tvecs": np.array([x.ravel() for x in tvecs]),
        "tries": np.array(tries),
        "ok": np.array(oks),
        "rounds": rounds,
        "rvec": r.ravel(),
        "tvec": t.ravel(),
    }


# ------------------------------------------------------------------------------------------------------------------
# synthetic scene-coordinate maps: random-depth surface + noise + outliers, known GT pose
# ------------------------------------------------------------------------------------------------------------------
def synth_scene(seed, h=60, w=80, f=525.0, sub=8, outlier_frac=0.3, noise=0.02, ppx=None, ppy=None):
    rng = np.random.default_rng(seed)
    ppx = w * sub / 2 if ppx is None else ppx
    ppy = h * sub / 2 if ppy is None else ppy
    # GT pose: scene -> camera (R, t), rotation <= 30 deg, translation <= 1 m
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, 30))
    R, _ = cv2.Rodrigues((axis * ang).reshape(3, 1)) if cv2 is not None else (None, None)
    t = rng.uniform(-1, 1, size=3) / np.sqrt(3)
    xs, ys = create_sampling(w, h, sub)
    gx, gy = np.meshgrid(xs.astype(np.float64), ys.astype(np.float64))
    # smooth random depth 1..5 m
    depth = 3.0 + 1.8 * np.sin(gx / 97.0 + rng.uniform(0, 6)) * np.cos(gy / 71.0 + rng.uniform(0, 6))
    depth = np.clip(depth + rng.uniform(-0.2, 0.2), 1.0, 5.0)
    cam = np.stack([(gx - ppx) / f * depth, (gy - ppy) / f * depth, depth], axis=0)  # [3,h,w]
    world = np.einsum("ij,jhw->ihw", R.T, cam - t.reshape(3, 1, 1))
    world = world + rng.normal(scale=noise, size=world.shape)
    out = rng.uniform(size=(h, w)) < outlier_frac
    world[:, out] = rng.uniform(-5, 5, size=(3, int(out.sum())))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return world.astype(np.float32)[None], np.linalg.inv(T), float(f), float(ppx), float(ppy)


def pose_error(T_est, T_gt):
    """(rotation error in degrees, translation error in metres) between two camera->world 4x4 poses."""
    dR = T_est[:3, :3].astype(np.float64).T @ T_gt[:3, :3]
    c = np.clip((np.trace(dR) - 1) / 2, -1, 1)
    return float(np.rad2deg(np.arccos(c))), float(np.linalg.norm(T_est[:3, 3].astype(np.float64) - T_gt[:3, 3]))
