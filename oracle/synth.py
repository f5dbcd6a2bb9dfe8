"""Seeded synthetic inputs + case tables shared by oracle/gen_golden.py and the tests.

TEST INFRASTRUCTURE (see oracle/rcan_oracle.py).  Nothing here touches /root/reference.
"""
import hashlib

import numpy as np
import torch

TASK = dict(  # config/task/SSS_SST-tiles-48.yaml
    batch_size=36, lr=5e-5, xyflip=True, origin=dict(x=0, y=0), tile_grid=dict(x=-1, y=-1),
    tile_size=dict(x=48, y=48), batch_domain="tiles", norm="lnorm", upsample_mode="cubic",
    downsample_mode="cubic", input_variables=dict(SSS="s", SST="t"), target_variables=["SSS", "SST"],
)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_hr(B, C, S, seed=4456, smooth=False):
    """Seeded HR batch.  smooth=True: low-frequency field + small noise, per-tile normalised,
    mimicking lnorm'd SSS/SST tiles."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, S, S, generator=g)
    if smooth:
        yy, xx = torch.meshgrid(torch.linspace(0, 1, S), torch.linspace(0, 1, S), indexing="ij")
        ph = torch.rand(B, C, 4, generator=g) * 6.2831853
        fld = sum(torch.sin((k + 1) * 3.1 * xx[None, None] + ph[..., k, None, None]) *
                  torch.cos((k + 1) * 2.3 * yy[None, None] + ph[..., (k + 1) % 4, None, None]) for k in range(4))
        x = fld + 0.05 * x
        x = (x - x.mean((2, 3), keepdim=True)) / x.std((2, 3), keepdim=True)
    return x


MODEL_CASES = {
    # name: (cfg overrides, B, LR size, channels, loss, smooth, full_output)
    "tiny_x4": (dict(nlayers=2, nblocks=2), 2, 12, 2, "l2", False, True),
    "tiny_x4_r16_charb": (dict(nlayers=2, nblocks=2, cbottleneck=16, loss_fn="charbonnier"), 2, 12, 2, "charbonnier", False, True),
    "tiny_x2_1ch": (dict(nlayers=1, nblocks=2, downscale_factors=[2]), 3, 10, 1, "l2", True, True),
    "tiny_x8_4ch": (dict(nlayers=1, nblocks=1, downscale_factors=[2, 2, 2]), 1, 8, 4, "l2", False, True),
    "tiny_x3": (dict(nlayers=1, nblocks=1, downscale_factors=[3]), 2, 9, 2, "l2", False, True),
    "small_x4": (dict(nlayers=4, nblocks=4), 4, 48, 2, "l2", True, False),
    # BASELINE config 2 depth (10 groups x 20 RCABs, reduction 16) on a batch the reference's CPU path finishes in seconds
    "full_x4_r16": (dict(nlayers=10, nblocks=20, cbottleneck=16), 2, 48, 2, "l2", True, False),
    # EDSR through the same factory (sres/model/edsr/network.py)
    "edsr_tiny_x4": (dict(name="edsr", nlayers=3), 2, 12, 2, "l2", False, True),
    "edsr_x2_rs01_charb": (dict(name="edsr", nlayers=2, res_scale=0.1, downscale_factors=[2], loss_fn="charbonnier"), 3, 10, 1,
                           "charbonnier", True, True),
    "edsr_16_x4": (dict(name="edsr", nlayers=16), 2, 48, 2, "l2", True, False),
}


def golden_file(name: str) -> str:
    return f"{name}.npz" if name.startswith("edsr_") else f"rcan_{name}.npz"


def synth_region(C, Y, X, seed, nan_frac=0.2):
    """Seeded (C,Y,X) float32 region with rectangular NaN 'land' patches (per variable for C>1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(Y, dtype=np.float32), np.arange(X, dtype=np.float32), indexing="ij")
    out = []
    for c in range(C):
        f = (np.sin(xx / 37.0 + c) * np.cos(yy / 29.0) * (3.0 + c) + 20.0 * (c + 1)).astype(np.float32)
        f += rng.standard_normal((Y, X), dtype=np.float32) * 0.1
        npatch = max(1, int(nan_frac * (Y // 192) * (X // 192) / 2))
        for _ in range(npatch):
            y0, x0 = int(rng.integers(0, Y)), int(rng.integers(0, X))
            f[y0:y0 + int(rng.integers(8, 260)), x0:x0 + int(rng.integers(8, 260))] = np.nan
        out.append(f[None])
    return out


TILE_CASES = {
    # name: (C, Y, X, tile, scale, seed, same_mask)
    "c1_1200": (1, 1200, 1200, 48, 4, 11, True),
    "c2_1200": (2, 1200, 1200, 48, 4, 12, True),
    "c1_odd": (1, 1000, 1423, 48, 4, 13, True),
    "c1_s2": (1, 500, 700, 24, 2, 14, True),
    "c2_diffmask": (2, 1200, 1200, 48, 4, 12, False),
}




LLC_CASES = {
    # name: (nx, roi, seed, land fraction)
    "nx24_roi": (24, dict(y0=13, ys=40, x0=5, xs=81), 21, 0.3),
    "nx16_full": (16, None, 22, 0.5),
    "nx40_west": (40, dict(y0=0, ys=120, x0=70, xs=90), 23, 0.1),
}


def synth_llc_files(folder, nx, seed, land_fraction, nvars=2, ntimes=2):
    """Write a template (hFacC-like: 0 on land, a positive fraction on ocean, a few -0.0 land points) and
    ocean-only big-endian float32 data files `raw/V{v}/V{v}.000{t}.shrunk` like config/dataset/swot_*.yaml names them."""
    import os
    rng = np.random.default_rng(seed)
    n = 13 * nx * nx
    tmpl = rng.random(n, dtype=np.float32) * 0.9 + 0.1
    land = rng.random(n) < land_fraction
    tmpl[land] = 0.0
    tmpl[np.nonzero(land)[0][::7]] = -0.0          # negative zero is land too (template != 0 is False)
    os.makedirs(os.path.join(folder, "meta"), exist_ok=True)
    tmpl.astype(">f4").tofile(os.path.join(folder, "meta", "hFacC_k0.data"))
    nocean = int((tmpl != 0).sum())
    for v in range(nvars):
        os.makedirs(os.path.join(folder, "raw", f"V{v}"), exist_ok=True)
        for t in range(ntimes):
            data = (rng.standard_normal(nocean).astype(np.float32) * 3.0 + 20.0 * (v + 1))
            data[::11] = np.float32("nan")                      # missing values inside the ocean stay NaN
            data.astype(">f4").tofile(os.path.join(folder, "raw", f"V{v}", f"V{v}.000{t + 3}.shrunk"))
    return dict(dataset_root=folder, dataset_files="raw/${dataset.varname}/${dataset.varname}.000${dataset.index}.shrunk",
                template="meta/hFacC_k0.data", nocean=nocean)
