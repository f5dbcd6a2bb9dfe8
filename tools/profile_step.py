"""In-situ per-category GPU time of one training step (SRES_PROFILE=1, eager launches, side stream off).
    python tools/profile_step.py            # BASELINE config 2: RCAN-full x4, 2-ch 48x48, batch 64
    python tools/profile_step.py x8         # BASELINE config 5: RCAN-full x8, 4-ch 96x96, batch 8"""
import ctypes as C, os, sys
os.environ["SRES_PROFILE"] = "1"; os.environ["SRES_CUDA_GRAPHS"] = "0"; os.environ["SRES_SIDE_STREAM"] = "0"
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn, _lib as L
dev = torch.device("cuda:0")
x8 = len(sys.argv) > 1 and sys.argv[1] == "x8"
ch, scale, B, S = (4, 8, 8, 768) if x8 else (2, 4, 64, 192)
m = snn.RCAN(nchannels_in=ch, nchannels_out=ch, nfeatures=64, nlayers=10, nblocks=20, cbottleneck=16, scale=scale, device=dev)
opt = snn.FusedAdam(m, lr=1e-4)
hr = torch.randn(B, ch, S, S, device=dev)
buf = C.create_string_buffer(8192)
for it in range(4):
    opt.zero_grad()
    loss = snn.loss(m(snn.bicubic_resize(hr, 1.0 / scale).requires_grad_(True)), hr, "l2")
    loss.backward(); opt.step()
    L.check(L.lib().sres_profile_report(buf, C.c_size_t(8192)), "report")
print(buf.value.decode())
