"""RCAN-full x4 inference throughput (bicubic down + forward, no grad) vs tile-batch size."""
import os, sys, time, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn
dev = torch.device("cuda:0")
m = snn.RCAN(nchannels_in=2, nchannels_out=2, nfeatures=64, nlayers=10, nblocks=20, cbottleneck=16, scale=4, device=dev)
for B in (32, 64, 128, 256):
    hr = torch.randn(B, 2, 192, 192, device=dev)
    with torch.no_grad():
        for _ in range(3): m(snn.bicubic_resize(hr, 0.25))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = max(4, 512 // B)
        for _ in range(n): m(snn.bicubic_resize(hr, 0.25))
        torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
    print(f"B={B:4d}: {t*1e3:7.2f} ms per batch, {B/t:7.0f} tiles/s, {B/t*36864/1e6:6.1f} output MP/s")
