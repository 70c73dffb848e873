import sys; sys.path.insert(0,'.')
import torch, numpy as np
from oracle import ace_ref
from acezero_b200.head import HeadEngine
rows=1024
sd=ace_ref.make_head_state(200,1,True)
eng=HeadEngine(1,True,(0,0,0),max_rows=rows,training=True); eng.load_state(sd)
opts=ace_ref.LossOptions(iterations=1000); lr_fn=ace_ref.one_cycle_lr(0.005,1000)
for it in range(5):
    bt=ace_ref.synth_batch(400+it,rows); g={k:v.cuda() for k,v in bt.items()}
    eng.set_hyper(lr_fn(it))
    lp=eng.loss_params("dyntanh", ace_ref.loss_weight(opts,it), rows)
    eng.train_fwd_bwd(rows, lp, g["target_px"], g["intrinsics"], g["intrinsics_inv"], aug_inv=g["aug_poses_inv"], pose_inv=g["poses_inv"], features=g["features"])
    torch.cuda.synchronize()
    gr=eng.grads
    print(it, "flag", int(eng.found_inf), "scale", float(eng.scaler_state[0]), "max|g|", float(gr.abs().max()), "finite", bool(torch.isfinite(gr).all()), "stats", eng.stats.tolist(), "cnt", eng.scaler_state[3].view(torch.int32).item())
    eng.adamw_step(use_scaler=True); torch.cuda.synchronize()
    print("   after adamw: scale", float(eng.scaler_state[0]), "step", float(eng.scaler_state[2]))

# ---- inspect the tail kernel's completion counter / block partials
import ctypes as C
ws = eng.workspace
total = int(eng.lib.acez_head_workspace_bytes(C.byref(eng.cfg))) - 1024
base_off = (-ws.data_ptr()) % 1024
off = base_off + total - 132096
part = ws[off:off + 131072].view(torch.float32).view(4096, 8)
cnt = ws[off + 131072: off + 131072 + 4].view(torch.int32)
print("counter now", int(cnt), "partials[0..2]", part[:3].tolist())
bt = ace_ref.synth_batch(999, rows); g = {k: v.cuda() for k, v in bt.items()}
lp = eng.loss_params("dyntanh", 50.0, rows)
eng.train_fwd_bwd(rows, lp, g["target_px"], g["intrinsics"], g["intrinsics_inv"], aug_inv=g["aug_poses_inv"], pose_inv=g["poses_inv"], features=g["features"])
torch.cuda.synchronize()
print("after another call: counter", int(cnt), "partials[0..2]", part[:3].tolist(), "stats", eng.stats.tolist())
