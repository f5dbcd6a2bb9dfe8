"""In-situ per-category GPU time of one RCAN-full training step (SRES_PROFILE=1, eager launches)."""
import ctypes as C, os, sys
os.environ["SRES_PROFILE"] = "1"; os.environ["SRES_CUDA_GRAPHS"] = "0"
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn, _lib as L
dev = torch.device("cuda:0")
m = snn.RCAN(nchannels_in=2, nchannels_out=2, nfeatures=64, nlayers=10, nblocks=20, cbottleneck=16, scale=4, device=dev)
opt = snn.FusedAdam(m, lr=1e-4)
hr = torch.randn(64, 2, 192, 192, device=dev)
buf = C.create_string_buffer(8192)
for it in range(4):
    opt.zero_grad()
    loss = snn.loss(m(snn.bicubic_resize(hr, 0.25).requires_grad_(True)), hr, "l2")
    loss.backward(); opt.step()
    L.check(L.lib().sres_profile_report(buf, C.c_size_t(8192)), "report")
print(buf.value.decode())
