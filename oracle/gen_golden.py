"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The reference (`/root/reference/sres`) is imported behind stub third-party modules
(oracle/ref_import.py) and executed on seeded synthetic inputs; the vectors it produces pin the
oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Inputs are regenerated from
seeds by the tests (same torch/numpy build), so the fixtures hold outputs only and stay small.
"""
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
from synth import LLC_CASES, MODEL_CASES, TASK, TILE_CASES, golden_file, sha, synth_hr, synth_llc_files, synth_region  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")

def gen_model_case(name, over, B, S, C, loss_name, smooth, full_out):
    cfg = O.model_cfg(**over)
    R.set_cfg(cfg, TASK)
    import importlib
    get_model = importlib.import_module(f"sres.model.{cfg['name']}.network").get_model  # manager.py:93-95
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
    if O.is_edsr(cfg):
        keep = ["head.0.weight", "head.0.bias", "tail.1.weight", "tail.1.bias", "body.0.body.0.bias", "body.0.body.2.bias",
                f"body.{cfg['nlayers']}.bias"]
        w2 = "body.0.body.2.weight"
    else:
        keep = ["head.0.weight", "head.0.bias", "tail.1.weight", "tail.1.bias", "body.0.body.0.body.0.bias",
                "body.0.body.0.body.3.conv_du.0.weight", "body.0.body.0.body.3.conv_du.2.bias", f"body.{cfg['nlayers']}.bias"]
        w2 = "body.0.body.0.body.2.weight"
        if cfg["nlayers"] > 4:   # full depth: also the far end of the body (last group, last RCAB, body-tail conv)
            g, r = cfg["nlayers"] - 1, cfg["nblocks"] - 1
            keep += [f"body.{g}.body.{r}.body.3.conv_du.0.weight", f"body.{g}.body.{r}.body.3.conv_du.2.bias",
                     f"body.{g}.body.{r}.body.0.bias", f"body.{g}.body.{cfg['nblocks']}.bias", "tail.0.0.bias"]
    for k in keep:
        out["grad::" + k] = grads[k].numpy()
        out["post::" + k] = post[k].numpy()
    w = grads[w2].numpy()
    out[f"grad::{w2}[:8,:8]"] = w[:8, :8]
    np.savez_compressed(os.path.join(GOLD, golden_file(name)), **out)
    print(f"{golden_file(name)}: loss={out['loss']:.6f} |out|={out['output_norm']:.4f} |g|={out['grad_global_norm']:.4e}")


def gen_tiles_case(name, C, Y, X, tile, scale, seed, same_mask):
    task = dict(TASK, tile_size=dict(x=tile, y=tile), batch_size=7)
    dfs = {2: [2], 4: [2, 2], 8: [2, 2, 2]}[scale]
    task.update(name="golden", dataset="synthetic")
    cfg = R.set_cfg(O.model_cfg(downscale_factors=dfs), task, platform=dict(cache="/tmp/sres_golden_cache"))
    from sres.base.source.swot.raw import SWOTRawDataLoader
    from sres.data.tiles import TileGrid, TileBatchIterator
    from sres.base.source.batch import xyflip as ref_xyflip
    from sres.controller.dual_trainer import ModelTrainer, denorm as ref_denorm
    var = synth_region(C, Y, X, seed)
    if C > 1 and same_mask:  # identical land mask for every variable (the case the reference handles)
        m = np.isnan(var[0])
        for v in var[1:]:
            v[np.isnan(v)] = 0.5
            v[m] = np.nan
    names = ["SSS", "SST"][:C]
    fake = object.__new__(SWOTRawDataLoader)  # the reference class, constructor (file I/O setup) skipped
    fake.tile_grid, fake.varnames, fake.time_index = TileGrid(), names, 0
    if not same_mask:
        # differing land masks + odd survivor count: the reference's final reshape raises
        try:
            SWOTRawDataLoader.get_tiles(fake, var)
            raised = ""
        except ValueError as e:
            raised = str(e)
        np.savez_compressed(os.path.join(GOLD, f"tiles_{name}.npz"), raised=raised)
        print(f"tiles_{name}: reference raised: {raised!r}")
        return
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


def gen_llc_case(name, nx, roi, seed, land):
    """The reference's file reader on synthetic LLC files: load_file's own lines (raw.py:136-143) around the reference's
    mds2d (called with the small test grid size; load_file itself hard-codes the default nx = 4320) and subset_roi."""
    import tempfile
    task = dict(TASK, name="golden", dataset="synthetic")
    with tempfile.TemporaryDirectory() as tmp:
        files = synth_llc_files(tmp, nx, seed, land)
        R.set_cfg(O.model_cfg(), task, dataset=dict(files, roi=roi) if roi is not None else dict(files), platform=dict(cache=tmp))
        from sres.base.source.swot.raw import filepath, subset_roi, template
        from sres.base.source.swot.util import mds2d
        from sres.base.util.config import cfg as ref_cfg
        out = {}
        for v in range(2):
            for t in (3, 4):
                for cparm, value in dict(varname=f"V{v}", index=t).items():          # raw.py:134-135
                    ref_cfg().dataset[cparm] = value
                path = filepath().replace("${dataset.varname}", f"V{v}").replace("${dataset.index}", str(t))
                var_template = np.fromfile(template(), ">f4")
                var_data = np.fromfile(path, ">f4")
                mask = (var_template != 0)
                var_template[mask] = var_data
                var_template[~mask] = np.nan
                sss_east, sss_west = mds2d(var_template, nx)
                result = np.expand_dims(np.c_[sss_east, sss_west.T[::-1, :]], 0)
                roi_data = subset_roi(result)
                out[f"sha_V{v}_{t}"] = sha(np.ascontiguousarray(roi_data))
                out[f"shape_V{v}_{t}"] = np.array(roi_data.shape)
                out[f"nan_V{v}_{t}"] = np.int64(np.isnan(roi_data).sum())
        out["sample"] = np.ascontiguousarray(roi_data[0, ::5, ::7])
        np.savez_compressed(os.path.join(GOLD, f"llc_{name}.npz"), **out)
        print(f"llc_{name}: roi{tuple(roi_data.shape)} nan={int(np.isnan(roi_data).sum())}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["model", "tiles", "llc"]   # e.g. `gen_golden.py model edsr_tiny_x4` regenerates one model case
    if "model" in which:
        only = [a for a in which if a in MODEL_CASES]
        for name, case in MODEL_CASES.items():
            if not only or name in only:
                gen_model_case(name, *case)
    if "tiles" in which:
        for name, case in TILE_CASES.items():
            gen_tiles_case(name, *case)
    if "llc" in which:
        for name, case in LLC_CASES.items():
            gen_llc_case(name, *case)


if __name__ == "__main__":
    main()
