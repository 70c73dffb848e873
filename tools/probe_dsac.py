"""GPU probe: CUDA DSAC* vs the cv2 oracle on synthetic scenes (prints diagnostics)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import dsacstar_ref as D
from acezero_b200 import dsac

def one(seed, hyps=64, max_tries=16):
    sc, Tgt, f, px, py = D.synth_scene(seed)
    ref = D.forward_rgb(sc, hyps, 10.0, f, px, py, 100.0, 100.0, 8, seed, max_tries)
    t = torch.from_numpy(sc).cuda()
    poses, inl, dbg = dsac.forward_rgb_batch(t, f, px, py, hyps, 10.0, 100.0, 100.0, 8, seed, max_tries, debug=True)
    torch.cuda.synchronize()
    pose = poses[0].cpu().numpy(); n_in = int(inl[0])
    tries = dbg["hyp_tries"][0].cpu().numpy(); scores = dbg["hyp_scores"][0].cpu().numpy().astype(np.float64)
    hp = dbg["hyp_poses"][0].cpu().numpy()
    same_tries = (tries == ref["tries"])
    dr = np.linalg.norm(hp[:, :3] - ref["hyp_rvecs"], axis=1); dt = np.linalg.norm(hp[:, 3:] - ref["hyp_tvecs"], axis=1)
    good = same_tries & ref["ok"]
    rel = np.abs(scores - ref["scores"]) / np.maximum(ref["scores"], 1e-6)
    print(f"seed {seed}: tries equal {same_tries.mean():.3f} | ok hyps {ref['ok'].sum()} | among equal&ok: max dr {dr[good].max():.2e} max dt {dt[good].max():.2e} max score rel {rel[good].max():.2e}")
    print(f"   best gpu {int(dbg['best'][0])} ref {ref['best']} | inliers gpu {n_in} ref {ref['inliers']} | rounds gpu {int(dbg['refine_rounds'][0])} ref {ref['rounds']}")
    e_ref = D.pose_error(ref["pose"], Tgt); e_gpu = D.pose_error(pose, Tgt); e_x = D.pose_error(pose, ref["pose"].astype(np.float64))
    print(f"   err vs GT: ref {e_ref[0]:.4f}deg {e_ref[1]*1000:.2f}mm | gpu {e_gpu[0]:.4f}deg {e_gpu[1]*1000:.2f}mm | gpu vs ref {e_x[0]:.5f}deg {e_x[1]*1000:.4f}mm")
    gb, rb = int(dbg['best'][0]), ref['best']
    if gb != rb:
        print(f"   MISMATCH scores: ref[{rb}]={ref['scores'][rb]:.6f} gpu[{rb}]={scores[rb]:.6f} | ref[{gb}]={ref['scores'][gb]:.6f} gpu[{gb}]={scores[gb]:.6f} | ok ref[{rb}]={ref['ok'][rb]} tries {ref['tries'][rb]}/{tries[rb]} dr {dr[rb]:.2e} dt {dt[rb]:.2e}")
        print("   ref top", np.argsort(-ref['scores'])[:4], np.sort(ref['scores'])[::-1][:4], "gpu top", np.argsort(-scores)[:4], np.sort(scores)[::-1][:4])
    bad = np.where(~same_tries)[0]
    if len(bad): print("   tries mismatch at", bad[:10], tries[bad[:10]], ref["tries"][bad[:10]])

for s in [1305, 1306, 1307, 7, 8]:
    one(s)
