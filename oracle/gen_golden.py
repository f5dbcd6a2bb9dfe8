"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The reference (`/root/reference/sres`) is imported behind stub third-party modules
(oracle/ref_import.py) and executed on seeded synthetic inputs; the vectors it produces pin the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Inputs are regenerated from
seeds by the tests (same torch/numpy build), so the fixtures hold outputs only and stay small.
"""
import hashlib
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as R  # noqa: E402
import rcan_oracle as O  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")

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
}


def gen_model_case(name, over, B, S, C, loss_name, smooth, full_out):
    cfg = O.model_cfg(**over)
    R.set_cfg(cfg, TASK)
    from sres.model.rcan.network import get_model
    from sres.base.util import array as ref_array
    from sres.controller.stats import l2loss as ref_l2
    from sres.controller.dual_trainer import ModelTrainer
    import sres.base.gpu as ref_gpu
    ref_gpu.get_device = lambda: torch.device("cpu")
    ref_array.get_device = ref_gpu.get_device
    torch.set_num_threads(8)
    scale = O.scale_of(cfg)
    model = get_model(nchannels_in=C, nchannels_out=C, device=torch.device("cpu"))
    sd = O.make_state_dict(cfg, C, C)
    assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict key order differs from the oracle"
    model.load_state_dict(sd)
    hr = synth_hr(B, C, S * scale, smooth=smooth)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=0.0)  # dual_trainer.py:126
    opt.zero_grad()
    target = hr.clone().requires_grad_(True)                    # array2tensor: requires_grad=True
    lr_in = ref_array.downsample(target)                       # dual_trainer.py:569
    prd = model(lr_in)                                          # :570
    fake = types.SimpleNamespace(eps=1e-6)
    if loss_name == "l2":
        loss = ref_l2(prd, ModelTrainer.conform_to_product(fake, prd, target))
    else:
        loss = ModelTrainer.charbonnier(fake, prd, ModelTrainer.conform_to_product(fake, prd, target))
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    post = {k: v.detach().clone() for k, v in model.state_dict().items()}
    interp = ref_array.upsample(lr_in.detach())
    out = dict(
        loss=np.float64(loss.item()),
        lr_input=lr_in.detach().numpy() if full_out else lr_in.detach().numpy()[:1],
        interp_sample=interp.numpy()[:1, :, ::3, ::3],
        output=prd.detach().numpy() if full_out else prd.detach().numpy()[:, :, ::8, ::8],
        output_norm=np.float64(prd.detach().double().norm().item()),
        grad_names=np.array(list(grads.keys())),
        grad_norms=np.array([g.double().norm().item() for g in grads.values()]),
        post_norms=np.array([post[k].double().norm().item() for k in grads.keys()]),
        grad_global_norm=np.float64(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())).item()),
    )
    keep = ["head.0.weight", "head.0.bias", "tail.1.weight", "tail.1.bias", "body.0.body.0.body.0.bias",
            "body.0.body.0.body.3.conv_du.0.weight", "body.0.body.0.body.3.conv_du.2.bias", f"body.{cfg['nlayers']}.bias"]
    for k in keep:
        out["grad::" + k] = grads[k].numpy()
        out["post::" + k] = post[k].numpy()
    w = grads["body.0.body.0.body.2.weight"].numpy()
    out["grad::body.0.body.0.body.2.weight[:8,:8]"] = w[:8, :8]
    np.savez_compressed(os.path.join(GOLD, f"rcan_{name}.npz"), **out)
    print(f"rcan_{name}: loss={out['loss']:.6f} |out|={out['output_norm']:.4f} |g|={out['grad_global_norm']:.4e}")


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
}


def gen_tiles_case(name, C, Y, X, tile, scale, seed, same_mask):
    task = dict(TASK, tile_size=dict(x=tile, y=tile), batch_size=7)
    dfs = {2: [2], 4: [2, 2], 8: [2, 2, 2]}[scale]
    cfg = R.set_cfg(O.model_cfg(downscale_factors=dfs), task)
    from sres.base.source.swot.raw import SWOTRawDataLoader
    from sres.data.tiles import TileGrid, TileBatchIterator
    from sres.base.source.batch import xyflip as ref_xyflip
    from sres.controller.dual_trainer import ModelTrainer, denorm as ref_denorm
    var = synth_region(C, Y, X, seed)
    if C > 1 and same_mask:  # identical land mask for every variable (the case the reference handles)
        m = np.isnan(var[0])
        for v in var[1:]:
            v[m] = np.nan
    names = ["SSS", "SST"][:C]
    fake = types.SimpleNamespace(tile_grid=TileGrid(), varnames=names, time_index=0)
    ts = SWOTRawDataLoader.get_tiles(fake, var)
    fake.timeslice = ts
    out = dict(tiles_sha=sha(ts.values), tiles_shape=np.array(ts.shape), tile_ids=np.asarray(ts.coords["tiles"].values),
               grid_shape=np.array([ts.attrs["grid_shape"]["y"], ts.attrs["grid_shape"]["x"]]),
               tiles_sample=ts.values[:2, :, ::24, ::24])
    nb = SWOTRawDataLoader.select_batch(fake, (7, 14))
    out.update(norm_sha=sha(nb.values), norm_sample=nb.values[:, :, ::48, ::48], norm_mean=nb.attrs["mean"], norm_std=nb.attrs["std"])
    assert SWOTRawDataLoader.select_batch(fake, (ts.shape[0], ts.shape[0] + 7)) is None
    last = SWOTRawDataLoader.select_batch(fake, (ts.shape[0] - 3, ts.shape[0] + 4))
    out["last_batch_shape"] = np.array(last.shape)
    # xyflip, all 8 orientations drawn through the reference's own random.randint call
    flips = []
    for sd in range(40):
        random.seed(sd)
        fb = ref_xyflip(nb.copy())
        flips.append((sd, fb.attrs["xyflip"], sha(fb.values)))
    out["flip_seeds"] = np.array([f[0] for f in flips]); out["flip_idx"] = np.array([f[1] for f in flips])
    out["flip_sha"] = np.array([f[2] for f in flips])
    # batch iterator
    random.seed(99)
    it = TileBatchIterator(ntiles=int(ts.shape[0]), randomize=True)
    out["batch_starts_shuffled"] = np.array([b["start"] for b in iter(it)])
    it = TileBatchIterator(ntiles=int(ts.shape[0]))
    out["batch_starts"] = np.array([b["start"] for b in iter(it)])
    # denorm + assemble_images over all batches (identity "model": products are the tiles themselves)
    batches = []
    for b in iter(TileBatchIterator(ntiles=int(ts.shape[0]))):
        bd = SWOTRawDataLoader.select_batch(fake, (b["start"], b["end"]))
        t = torch.from_numpy(np.ascontiguousarray(bd.values))
        lo = torch.from_numpy(np.ascontiguousarray(bd.values[:, :, ::scale, ::scale]))
        batches.append(dict(input=ref_denorm(lo, bd.attrs), target=ref_denorm(t, bd.attrs)))
    for ivar in range(C):
        imgs = ModelTrainer.assemble_images(None, batches, ivar, ts.coords["tiles"].values, ts.attrs["grid_shape"])
        for k, da in imgs.items():
            out[f"image_{ivar}_{k}_sha"] = sha(da.values)
            out[f"image_{ivar}_{k}_shape"] = np.array(da.values.shape)
            out[f"image_{ivar}_{k}_dtype"] = str(da.values.dtype)
            out[f"image_{ivar}_{k}_nan"] = np.int64(np.isnan(da.values).sum())
    np.savez_compressed(os.path.join(GOLD, f"tiles_{name}.npz"), **out)
    print(f"tiles_{name}: tiles{tuple(ts.shape)} grid={ts.attrs['grid_shape']} ids[:6]={ts.coords['tiles'].values[:6]}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["model", "tiles"]
    if "model" in which:
        for name, case in MODEL_CASES.items():
            gen_model_case(name, *case)
    if "tiles" in which:
        for name, case in TILE_CASES.items():
            gen_tiles_case(name, *case)


if __name__ == "__main__":
    main()
