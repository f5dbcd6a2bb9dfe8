"""Config 4: inference over a whole region through the mirrored controller API -- extract tiles, lnorm,
bicubic down, RCAN-full forward (no grad), denorm, stitch -- on the synthetic stand-in for the swot_20-20e roi
(3000 x 17280, 15 x 90 grid, 1350 tiles per variable, ~20 % land).  Also times one EDSR-16 and one RCAN-full training
step for the record.  Prints output megapixels per second (36 864 px per tile and variable)."""
import os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres.base.util.config import ConfigContext
from sres.controller.config import ResultStructure
from sres.controller.workflow import WorkflowController
from sres_b200 import nn as snn

dev = torch.device("cuda:0")
tmp = tempfile.mkdtemp()
wc = WorkflowController("sres", dict(task="SSS_SST-tiles-48", dataset="synthetic_20-20e", platform="local"), seed=1)
wc.initialize("sres", "rcan-10-20-64", **{"model.cbottleneck": 16, "task.batch_size": 64, "task.tile_order": "corrected",
                                           "dataset.ntimes": 2, "platform.results": tmp})
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    images, losses = wc.inference(0, ResultStructure.Image)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    img = images["SSS"]["model"]
    ntiles = int(np.isfinite(img[::192, ::192]).sum())
    mp = 2 * ntiles * 192 * 192 / 1e6
    print(f"region inference pass {rep}: {dt:.3f} s, {ntiles} tiles x 2 vars, image {img.shape} -> {mp / dt:.1f} output MP/s end to end "
          f"(host images included)")
ConfigContext.deactivate()

def step_time(model, B=64, n=8):
    opt = snn.FusedAdam(model, lr=1e-4)
    hr = torch.randn(B, 2, 192, 192, device=dev)
    for it in range(n + 3):
        if it == 3:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        opt.zero_grad()
        snn.loss(model(snn.bicubic_resize(hr, 0.25).requires_grad_(True)), hr, "l2").backward()
        opt.step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

edsr = snn.EDSR(nchannels_in=2, nchannels_out=2, nfeatures=64, nlayers=16, scale=4, device=dev)
t = step_time(edsr)
print(f"EDSR-16 x4 train step, B=64: {t * 1e3:.2f} ms = {64 / t:.0f} tiles/s")
with torch.no_grad():
    x = torch.randn(64, 2, 48, 48, device=dev)
    for _ in range(3): edsr(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): edsr(x)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 20
print(f"EDSR-16 x4 forward, B=64: {t * 1e3:.2f} ms = {64 / t:.0f} tiles/s = {64 / t * 36864 / 1e6:.0f} output MP/s")
