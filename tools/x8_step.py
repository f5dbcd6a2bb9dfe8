"""One (or a few) x8 training steps, eager, for `ncu` launch lists of BASELINE config 5 (RCAN-full x8, 4-ch 96x96, batch 8)."""
import os, sys
os.environ.setdefault("SRES_CUDA_GRAPHS", "0")
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = snn.RCAN(nchannels_in=4, nchannels_out=4, nfeatures=64, nlayers=10, nblocks=20, cbottleneck=16, scale=8, device=dev)
opt = snn.FusedAdam(m, lr=1e-4)
hr = torch.randn(B, 4, 768, 768, device=dev)
for it in range(2):
    opt.zero_grad()
    loss = snn.loss(m(snn.bicubic_resize(hr, 1.0 / 8).requires_grad_(True)), hr, "l2")
    loss.backward(); opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
