"""Tile iteration and grid arithmetic, mirror of sres/data/tiles.py:8-158 (pure host logic)."""
import math
import random
from typing import Dict, List, Optional, Tuple

import numpy as np

from sres.base.util.config import cfg


class TileIterator(object):

    def __init__(self, **kwargs):
        self.randomize: bool = kwargs.get("randomize", False)
        self._batch_losses = {}
        self.index: int = 0
        self.next_index = 0

    def batch_losses(self, ltype) -> List[float]:
        return self._batch_losses.setdefault(ltype, [])

    def clear_batch_losses(self, ltype):
        self._batch_losses[ltype] = []

    def register_loss(self, ltype: str, loss: float):
        self.batch_losses(ltype).append(loss)

    def accumulate_loss(self, ltype: str):
        accum_loss = np.array(self.batch_losses(ltype)).mean()
        self.clear_batch_losses(ltype)
        return accum_loss

    @classmethod
    def get_iterator(cls, **kwargs):
        domain = cfg().task.get("batch_domain", "tiles")
        if domain == "tiles":
            return TileBatchIterator(**kwargs)
        if domain == "time":
            return TileGridIterator(**kwargs)
        return None


class TileBatchIterator(TileIterator):
    """Batch start indices range(0, ntiles, batch_size), optionally random.shuffle'd (tiles.py:48-74).
    The end index is not clipped; the loader clips it (swot/raw.py:163)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.batch_size: int = cfg().task.batch_size
        self.ntiles: int = kwargs.get("ntiles", 0)
        assert self.ntiles > 0, "Must provide ntiles for TileBatchIterator"
        self.batch_start_idxs: List[int] = list(range(0, self.ntiles, self.batch_size))
        if self.randomize:
            random.shuffle(self.batch_start_idxs)

    def __iter__(self):
        self.next_index = 0
        return self

    @property
    def active(self):
        return (self.ntiles == 0) or (self.next_index < len(self.batch_start_idxs))

    def __next__(self) -> Dict[str, int]:
        if not self.active:
            raise StopIteration()
        self.index = self.next_index
        bstart = self.batch_start_idxs[self.index]
        self.next_index = self.index + 1
        return dict(start=bstart, end=bstart + self.batch_size)


class TileGridIterator(TileIterator):

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.grid = TileGrid()
        self.regular_grid: List[Dict[str, int]] = list(self.grid.get_tile_locations(**kwargs).values())

    def __iter__(self):
        if self.randomize:
            random.shuffle(self.regular_grid)
        self.next_index = 0
        return self

    @property
    def active(self):
        return self.next_index < len(self.regular_grid)

    def __next__(self) -> Dict[str, int]:
        if not self.active:
            raise StopIteration()
        self.index = self.next_index
        self.next_index = self.index + 1
        return self.regular_grid[self.index]


class TileGrid(object):

    def __init__(self):
        self.origin: Dict[str, int] = cfg().task.get("origin", {})
        self.tile_grid: Optional[Dict[str, int]] = None
        self.tile_size: Dict[str, int] = cfg().task.tile_size
        self.tlocs: Dict[Tuple[int, int], Dict[str, int]] = {}
        self.upsample_factor = math.prod(cfg().model.downscale_factors)

    def get_global_grid_shape(self, **kwargs) -> Dict[str, int]:
        image_shape = kwargs.get("image_shape", cfg().task.get("image_shape", None))
        if image_shape is None:
            return dict(x=1, y=1)
        ts = self.get_full_tile_size()
        return {dim: image_shape[dim] // ts[dim] for dim in ["x", "y"]}

    def get_grid_shape(self, **kwargs) -> Dict[str, int]:
        glob = self.get_global_grid_shape(**kwargs)
        cgs = cfg().task.tile_grid
        self.tile_grid = {dim: (cgs[dim] if (cgs[dim] >= 0) else glob[dim]) for dim in ["x", "y"]}
        return self.tile_grid

    def get_active_region(self, **kwargs) -> Dict[str, Tuple[int, int]]:
        ts = self.get_full_tile_size()
        gs = self.get_grid_shape(**kwargs)
        return {d: (self.origin[d], self.origin[d] + ts[d] * gs[d]) for d in ["x", "y"]}

    def get_tile_size(self, highres: bool = False) -> Dict[str, int]:
        sf = self.upsample_factor if highres else 1
        return {d: self.tile_size[d] * sf for d in ["x", "y"]}

    def get_full_tile_size(self) -> Dict[str, int]:
        return {d: self.tile_size[d] * self.upsample_factor for d in ["x", "y"]}

    def get_tile_origin(self, ix: int, iy: int, highres: bool = False) -> Dict[str, int]:
        sf = self.upsample_factor if highres else 1
        return {d: self.origin[d] + self.cdim(ix, iy, d) * self.tile_size[d] * sf for d in ["x", "y"]}

    def get_tile_locations(self, **kwargs) -> Dict[Tuple[int, int], Dict[str, int]]:
        highres: bool = kwargs.get("highres", False)
        selected_tile = kwargs.get("selected_tile", None)
        if len(self.tlocs) == 0:
            if self.tile_grid is None:
                self.get_grid_shape(**kwargs)
            for ix in range(self.tile_grid["x"]):
                for iy in range(self.tile_grid["y"]):
                    if (selected_tile is None) or ((ix, iy) == selected_tile):
                        self.tlocs[(ix, iy)] = self.get_tile_origin(ix, iy, highres)
        return self.tlocs

    @classmethod
    def cdim(cls, ix: int, iy: int, dim: str) -> int:
        return ix if dim == "x" else iy
