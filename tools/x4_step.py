"""One (or a few) x4 training steps, eager, for `ncu` checks of BASELINE config 2 (RCAN-full x4, 2-ch 48x48, batch 64).
    python tools/x4_step.py [batch] [nblocks]"""
import os, sys
os.environ.setdefault("SRES_CUDA_GRAPHS", "0")
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 20
m = snn.RCAN(nchannels_in=2, nchannels_out=2, nfeatures=64, nlayers=10, nblocks=R, cbottleneck=16, scale=4, device=dev)
opt = snn.FusedAdam(m, lr=1e-4)
hr = torch.randn(B, 2, 192, 192, device=dev)
for it in range(2):
    opt.zero_grad()
    loss = snn.loss(m(snn.bicubic_resize(hr, 0.25).requires_grad_(True)), hr, "l2")
    loss.backward(); opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
