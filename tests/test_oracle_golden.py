"""The oracle (oracle/*.py) against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py, run in the build container).  CPU only."""
import os
import random

import numpy as np
import pytest
import torch

import rcan_oracle as O
import tiles_oracle as T
from synth import MODEL_CASES, TILE_CASES, golden_file, sha, synth_hr, synth_region

torch.set_num_threads(8)


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_rcan_oracle_matches_reference(name, golden_dir):
    over, B, S, C, loss_name, smooth, full_out = MODEL_CASES[name]
    gold = np.load(os.path.join(golden_dir, golden_file(name)))
    cfg = O.model_cfg(**over)
    scale = O.scale_of(cfg)
    sd = O.make_state_dict(cfg, C, C)
    hr = synth_hr(B, C, S * scale, smooth=smooth)
    adam = O.AdamState(sd, lr=1e-4)
    lr_in = O.downsample(hr, scale)
    loss, prd, grads = O.train_step(hr, sd, cfg, adam, loss_name)
    # same torch build, same ops, same thread count: bit-for-bit except where the op order of the
    # restated Adam differs from torch.optim's foreach kernels (1 ulp)
    ref_lr = gold["lr_input"]
    np.testing.assert_array_equal(lr_in.numpy()[: ref_lr.shape[0]], ref_lr)
    out = prd.numpy() if full_out else prd.numpy()[:, :, ::8, ::8]
    np.testing.assert_allclose(out, gold["output"], rtol=0, atol=1e-6)
    assert abs(loss - float(gold["loss"])) <= 1e-7 * max(1.0, abs(float(gold["loss"])))
    names = [str(n) for n in gold["grad_names"]]
    assert names == list(sd.keys())
    gn = np.array([grads[k].double().norm().item() for k in names])
    np.testing.assert_allclose(gn, gold["grad_norms"], rtol=2e-5, atol=1e-12)
    pn = np.array([sd[k].double().norm().item() for k in names])
    np.testing.assert_allclose(pn, gold["post_norms"], rtol=1e-6)
    for key in gold.files:
        if key.startswith("grad::") and "[" not in key:
            k = key[6:]
            g = grads[k].numpy()
            np.testing.assert_allclose(g, gold[key], rtol=1e-4, atol=1e-6 * np.abs(gold[key]).max())
        if key.startswith("post::"):
            np.testing.assert_allclose(sd[key[6:]].numpy(), gold[key], rtol=0, atol=2e-7)
    interp = O.upsample(lr_in, scale)
    np.testing.assert_array_equal(interp.numpy()[:1, :, ::3, ::3], gold["interp_sample"])


def _tile_inputs(C, Y, X, seed, same_mask):
    var = synth_region(C, Y, X, seed)
    if C > 1 and same_mask:
        m = np.isnan(var[0])
        for v in var[1:]:
            v[np.isnan(v)] = 0.5
            v[m] = np.nan
    return var


@pytest.mark.parametrize("name", [n for n in TILE_CASES if TILE_CASES[n][6]])
def test_tiles_oracle_matches_reference(name, golden_dir):
    C, Y, X, tile, scale, seed, same_mask = TILE_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"tiles_{name}.npz"))
    var = _tile_inputs(C, Y, X, seed, same_mask)
    ts = dict(x=tile, y=tile)
    tiles, ids, gs = T.get_tiles(var, ts, scale)
    assert list(tiles.shape) == list(gold["tiles_shape"])
    assert sha(tiles) == str(gold["tiles_sha"])
    np.testing.assert_array_equal(ids, gold["tile_ids"])
    assert [gs["y"], gs["x"]] == list(gold["grid_shape"])
    nb, stats = T.lnorm(T.select_batch(tiles, 7, 14))
    assert sha(nb) == str(gold["norm_sha"])
    np.testing.assert_array_equal(stats["mean"], gold["norm_mean"])
    np.testing.assert_array_equal(stats["std"], gold["norm_std"])
    assert T.select_batch(tiles, tiles.shape[0], tiles.shape[0] + 7) is None
    assert list(T.select_batch(tiles, tiles.shape[0] - 3, tiles.shape[0] + 4).shape) == list(gold["last_batch_shape"])
    for sd_, fi, fsha in zip(gold["flip_seeds"], gold["flip_idx"], gold["flip_sha"]):
        random.seed(int(sd_))
        assert random.randint(0, 7) == int(fi)
        assert sha(T.xyflip(nb, int(fi))) == str(fsha)
    assert set(int(f) for f in gold["flip_idx"]) == set(range(8))
    random.seed(99)
    sh = [b["start"] for b in T.tile_batches(tiles.shape[0], 7, randomize=True)]
    np.testing.assert_array_equal(sh, gold["batch_starts_shuffled"])
    np.testing.assert_array_equal([b["start"] for b in T.tile_batches(tiles.shape[0], 7)], gold["batch_starts"])
    batches = []
    for b in T.tile_batches(tiles.shape[0], 7):
        raw = T.select_batch(tiles, b["start"], b["end"])
        bd, st = T.lnorm(raw)
        batches.append(dict(input=T.denorm(np.ascontiguousarray(bd[:, :, ::scale, ::scale]), st), target=T.denorm(bd, st)))
    for ivar in range(C):
        imgs = T.assemble_images(batches, ivar, ids, gs)
        for k, img in imgs.items():
            assert list(img.shape) == list(gold[f"image_{ivar}_{k}_shape"])
            assert str(img.dtype) == str(gold[f"image_{ivar}_{k}_dtype"])
            assert int(np.isnan(img).sum()) == int(gold[f"image_{ivar}_{k}_nan"])
            assert sha(img) == str(gold[f"image_{ivar}_{k}_sha"])


def test_tiles_oracle_reproduces_reference_reshape_failure(golden_dir):
    C, Y, X, tile, scale, seed, same_mask = TILE_CASES["c2_diffmask"]
    gold = np.load(os.path.join(golden_dir, "tiles_c2_diffmask.npz"))
    assert str(gold["raised"]).startswith("cannot reshape")
    var = _tile_inputs(C, Y, X, seed, same_mask)
    with pytest.raises(ValueError, match="cannot reshape"):
        T.get_tiles(var, dict(x=tile, y=tile), scale)
    tiles, ids, gs = T.get_tiles(var, dict(x=tile, y=tile), scale, mode="corrected")
    assert tiles.shape[1] == 2 and np.isfinite(tiles.mean(axis=(2, 3))).all()


def test_flops_formula():
    cfg = O.model_cfg()
    assert abs(O.flops_per_tile(cfg, 2, 2) / 1e9 - 73.31) < 0.01
    assert abs(O.flops_per_tile(O.model_cfg(nlayers=4, nblocks=4), 2, 2) / 1e9 - 9.773) < 0.001
    n = sum(int(np.prod(s)) for s in O.param_shapes(cfg, 2, 2).values())
    assert n == 16313602 and len(O.param_shapes(cfg, 2, 2)) == 1630


def test_llc_reader_oracle_matches_reference(golden_dir, tmp_path):
    """oracle/tiles_oracle.py:llc_load_file (restatement of raw.py:133-145 + util.py:3-55 + subset_roi) against the hashes the
    reference's own mds2d / subset_roi produced on the same seeded synthetic LLC files (oracle/gen_golden.py:gen_llc_case)."""
    import tiles_oracle as T
    from synth import LLC_CASES, sha, synth_llc_files
    for name, (nx, roi, seed, land) in LLC_CASES.items():
        gold = np.load(os.path.join(golden_dir, f"llc_{name}.npz"))
        folder = str(tmp_path / name)
        files = synth_llc_files(folder, nx, seed, land)
        for v in range(2):
            for t in (3, 4):
                data = os.path.join(folder, "raw", f"V{v}", f"V{v}.000{t}.shrunk")
                got = T.llc_load_file(os.path.join(folder, files["template"]), data, nx, roi)
                assert list(got.shape) == list(gold[f"shape_V{v}_{t}"]) and got.dtype == np.float32
                assert int(np.isnan(got).sum()) == int(gold[f"nan_V{v}_{t}"])
                assert sha(np.ascontiguousarray(got)) == str(gold[f"sha_V{v}_{t}"])
