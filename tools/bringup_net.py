"""GPU bring-up of the whole RCAN path against the CPU oracle (prints per-tensor errors)."""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rcan_oracle as O
from synth import MODEL_CASES, synth_hr
from sres_b200 import nn as snn

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="tiny_x4")
ap.add_argument("--B", type=int, default=0)
a = ap.parse_args()
over, B, S, C, loss_name, smooth, _ = MODEL_CASES[a.case]
if a.B: B = a.B
cfg = O.model_cfg(**over); scale = O.scale_of(cfg)
torch.set_num_threads(os.cpu_count())
sd = O.make_state_dict(cfg, C, C)
hr = synth_hr(B, C, S * scale, smooth=smooth)
t0 = time.time()
loss_o, prd_o, grads_o = O.loss_and_grads(hr, sd, cfg, loss_name)
lr_o = O.downsample(hr, scale)
print(f"oracle: loss={loss_o:.6f} ({time.time()-t0:.1f}s on {os.cpu_count()} cores)")
dev = torch.device("cuda:0")
m = snn.RCAN(nchannels_in=C, nchannels_out=C, nfeatures=64, nlayers=cfg["nlayers"], nblocks=cfg["nblocks"],
             cbottleneck=cfg["cbottleneck"], scale=scale, device=dev)
m.load_state_dict(sd)
hr_d = hr.to(dev)
lr_d = snn.bicubic_resize(hr_d, 1.0 / scale)
print("bicubic down max abs err", (lr_d.cpu() - lr_o).abs().max().item())
up_d = snn.bicubic_resize(lr_d, scale); print("bicubic up max abs err", (up_d.cpu() - O.upsample(lr_o, scale)).abs().max().item())
prd = m(lr_d)
loss = snn.loss(prd, hr_d, loss_name)
loss.backward()
torch.cuda.synchronize()
def rel(a_, b_): return ((a_ - b_).norm() / (b_.norm() + 1e-30)).item()
print(f"RESULT case={a.case} B={B} out_rel_l2={rel(prd.detach().cpu(), prd_o):.4e} loss={loss.item():.6f} vs {loss_o:.6f} nan={torch.isnan(prd).sum().item()}")
errs = []
num = den = 0.0
for k, p in m.named_parameters():
    g = p.grad.detach().cpu(); go = grads_o[k]
    num += (g - go).double().pow(2).sum().item(); den += go.double().pow(2).sum().item()
    errs.append((rel(g, go), k, go.norm().item()))
print(f"RESULT grad_global_rel_l2={np.sqrt(num/den):.4e}")
errs.sort(reverse=True)
for e, k, nrm in errs[:12]: print(f"   {e:.3e}  {k}  |g|={nrm:.3e}")
# one fused Adam step vs oracle Adam
opt = snn.FusedAdam(m, lr=1e-4)
opt.step(); torch.cuda.synchronize()
adam = O.AdamState(sd, lr=1e-4)
with torch.no_grad(): adam.step(sd, grads_o)
num = den = 0.0
for k, p in m.named_parameters():
    num += (p.detach().cpu() - sd[k]).double().pow(2).sum().item(); den += sd[k].double().pow(2).sum().item()
print(f"RESULT post_adam_param_rel_l2={np.sqrt(num/den):.4e}")
# inference path (no grad) must agree with the training forward
with torch.no_grad():
    m2 = m(lr_d)
print("eval-vs-train forward after step: finite", torch.isfinite(m2).all().item())
