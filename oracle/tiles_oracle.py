"""CPU oracle of the tile extraction / normalisation / stitching either side of the RCAN model.

TEST INFRASTRUCTURE, not product code (see oracle/rcan_oracle.py for the rules).  Plain numpy
restatement of the reference's index arithmetic, pinned by oracle/gen_golden.py against the
reference's own functions executed in the build container (tests/golden/tiles_*.npz).

  grid_shape / active_region   sres/data/tiles.py:110-127 (TileGrid)
  get_tiles                    sres/base/source/swot/raw.py:216-233
  tile_batches                 sres/data/tiles.py:48-74 (TileBatchIterator)
  select_batch + lnorm         sres/base/source/swot/raw.py:160-167, :169-183, :211-214
  xyflip                       sres/base/source/batch.py:33-49
  denorm                       sres/controller/dual_trainer.py:67-77
  assemble_images              sres/controller/dual_trainer.py:449-480
"""
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def full_tile_size(tile_size: Dict[str, int], scale: int) -> Dict[str, int]:
    return {d: tile_size[d] * scale for d in ("x", "y")}  # tiles.py:134-135


def grid_shape(image_shape: Dict[str, int], tile_size: Dict[str, int], scale: int,
               cfg_tile_grid: Optional[Dict[str, int]] = None) -> Dict[str, int]:
    """tiles.py:110-121: tiles per axis = image // (tile_size*scale) unless fixed (>= 0) in the task."""
    ts = full_tile_size(tile_size, scale)
    glob = {d: image_shape[d] // ts[d] for d in ("x", "y")}
    cfg_tile_grid = cfg_tile_grid or {"x": -1, "y": -1}
    return {d: (cfg_tile_grid[d] if cfg_tile_grid[d] >= 0 else glob[d]) for d in ("x", "y")}


def active_region(origin: Dict[str, int], tile_size, scale, gshape) -> Dict[str, Tuple[int, int]]:
    ts = full_tile_size(tile_size, scale)  # tiles.py:123-127
    return {d: (origin[d], origin[d] + ts[d] * gshape[d]) for d in ("x", "y")}


def get_tiles(var_data: Sequence[np.ndarray], tile_size: Dict[str, int], scale: int,
              origin: Optional[Dict[str, int]] = None, cfg_tile_grid=None, mode: str = "reference"):
    """raw.py:216-233.  var_data: list of (1,Y,X) arrays, one per variable.

    Returns (tiles (N,C,T,T), tile_ids (N,), grid_shape).  mode="reference" reproduces the
    reference bit for bit, including its channel-major flattening (for C > 1 `result[i, ch]` is flat
    tile C*i+ch of the concatenated variable list).  mode="corrected" keeps tile i of
    every variable together and drops a tile when ANY variable is non-finite there."""
    origin = origin or {"x": 0, "y": 0}
    raw = np.concatenate(list(var_data), axis=0)
    c, Y, X = raw.shape
    ts = full_tile_size(tile_size, scale)
    gs = grid_shape(dict(x=X, y=Y), tile_size, scale, cfg_tile_grid)
    roi = active_region(origin, tile_size, scale, gs)
    region = raw[..., roi["y"][0]:roi["y"][1], roi["x"][0]:roi["x"][1]]
    tile_data = region.reshape(c, gs["y"], ts["y"], gs["x"], ts["x"])
    if mode == "reference":
        tiles = np.swapaxes(tile_data, 2, 3).reshape(c * gs["y"] * gs["x"], ts["y"], ts["x"])
        msk = np.isfinite(tiles.mean(axis=-1).mean(axis=-1))
        ctiles = np.compress(msk, tiles, 0)
        idxs = np.compress(msk, np.arange(tiles.shape[0]), 0)
        result = ctiles.reshape(ctiles.shape[0] // c, c, ts["y"], ts["x"])
        return result, idxs[0:result.shape[0]], gs
    if mode == "corrected":
        tiles = np.swapaxes(tile_data, 2, 3).reshape(c, gs["y"] * gs["x"], ts["y"], ts["x"])
        msk = np.isfinite(tiles.mean(axis=-1).mean(axis=-1)).all(axis=0)
        result = np.ascontiguousarray(np.swapaxes(tiles[:, msk], 0, 1))
        return result, np.nonzero(msk)[0], gs
    raise ValueError(mode)


def tile_batches(ntiles: int, batch_size: int, randomize: bool = False, rng: Optional[random.Random] = None):
    """tiles.py:48-74: start indices range(0,ntiles,batch_size), optionally random.shuffle'd."""
    starts = list(range(0, ntiles, batch_size))
    if randomize:
        (rng or random).shuffle(starts)
    return [dict(start=s, end=s + batch_size) for s in starts]


def select_batch(timeslice: np.ndarray, start: int, end: int) -> Optional[np.ndarray]:
    """raw.py:160-167 (slice only; the caller normalises)."""
    n = timeslice.shape[0]
    if start < n:
        return timeslice[start:min(end, n)]
    return None


def lnorm(batch: np.ndarray):
    """raw.py:176-183, 211-214: per tile, per channel (x - mean)/std over (y,x), NaN-skipping,
    population std, in the array's own dtype.  Returns (normalised, dict(mean,std) each (B,C,1,1))."""
    mean = np.nanmean(batch, axis=(2, 3))
    std = np.nanstd(batch, axis=(2, 3))
    bdims = (batch.shape[0], batch.shape[1], 1, 1)
    out = (batch - mean.reshape(bdims)) / std.reshape(bdims)
    return out, dict(mean=mean.reshape(bdims), std=std.reshape(bdims))


def xyflip(batch: np.ndarray, flip_index: int) -> np.ndarray:
    """batch.py:37-49: bit0 flips x, bit1 flips y, bit2 transposes y<->x."""
    if flip_index % 2 == 1:
        batch = np.flip(batch, axis=-1)
    if (flip_index // 2) % 2 == 1:
        batch = np.flip(batch, axis=-2)
    if flip_index // 4 == 1:
        batch = np.swapaxes(batch, -1, -2)
    return batch


def denorm(normed: np.ndarray, norm_data: Dict[str, np.ndarray]) -> np.ndarray:
    """dual_trainer.py:67-77."""
    if "mean" in norm_data:
        normed = (normed * norm_data["std"]) + norm_data["mean"]
    if "max" in norm_data:
        normed = (normed * (norm_data["max"] - norm_data["min"])) + norm_data["min"]
    return normed


def assemble_images(batches: List[Dict[str, np.ndarray]], ivar: int, tile_ids: np.ndarray,
                    gshape: Dict[str, int]) -> Dict[str, np.ndarray]:
    """dual_trainer.py:449-480: place tile `tid` at (tid // gx, tid % gx), NaN elsewhere, np.block.
    (dtype follows np.block: float64 when an unfilled float64 NaN cell remains, else the tiles'.)"""
    out = {}
    for image_type in batches[0].keys():
        tidx0, grid = 0, None
        for b in batches:
            batch = b[image_type][:, ivar, :, :]
            if grid is None:
                empty = np.full(list(batch.shape[-2:]), np.nan)
                grid = [[empty] * gshape["x"] for _ in range(gshape["y"])]
            for bidx in range(batch.shape[0]):
                tid = int(tile_ids[tidx0 + bidx])
                grid[tid // gshape["x"]][tid % gshape["x"]] = batch[bidx].squeeze()
            tidx0 += batch.shape[0]
        out[image_type] = np.block(grid)
    return out


# ------------------------------------------------------------------------------------------------
# raw LLC4320 reader (sres/base/source/swot/raw.py:133-145, :38-45; sres/base/source/swot/util.py:3-55)
# ------------------------------------------------------------------------------------------------
def llc_rearrange(d: np.ndarray, nx: int):
    """util.py:3-7: faces 1-3 | faces 4-6 side by side (east), faces 8-13 (west); face 7 (Arctic) is dropped."""
    deast = np.c_[d[:nx * nx * 3].reshape(3 * nx, nx), d[nx * nx * 3:nx * nx * 6].reshape(3 * nx, nx)]
    dwest = d[nx * nx * 7:].reshape(nx * 2, nx * 3)
    return deast, dwest


def llc_load_file(template_path: str, data_path: str, nx: int, roi=None) -> np.ndarray:
    """raw.py:133-145 with the grid size as a parameter (the reference hard-codes mds2d's default nx = 4320) and
    subset_roi (raw.py:38-45) applied: (1, ys, xs) float32."""
    var_template = np.fromfile(template_path, ">f4")
    var_data = np.fromfile(data_path, ">f4")
    mask = (var_template != 0)
    var_template[mask] = var_data
    var_template[~mask] = np.nan
    east, west = llc_rearrange(var_template, nx)
    result = np.expand_dims(np.c_[east, west.T[::-1, :]], 0)
    if roi is not None:
        x0, xs = roi.get("x0", 0), roi.get("xs", result.shape[-1])
        y0, ys = roi.get("y0", 0), roi.get("ys", result.shape[-2])
        result = result[..., y0:y0 + ys, x0:x0 + xs]
    return result
