"""Tile-batch schedule and tile-grid geometry for the tile path.

What the trainer needs from the reference's sres/data/tiles.py, restated as two small value types:

  BatchSchedule  the order in which the tiles of a timeslice are visited: batch start indices
                 0, B, 2B, ... (B = task.batch_size), shuffled with Python's `random` when asked
                 (reference: TileBatchIterator, tiles.py:48-74), plus the per-timeslice loss log the
                 training loop keeps next to it (tiles.py:17-29).
  TileGrid       how many full high-resolution tiles fit the region and which part of it they cover
                 (reference: TileGrid.get_grid_shape / get_active_region, tiles.py:110-127).

`TileIterator.get_iterator` and `TileBatchIterator` keep the reference's entry-point names; the `time`
batch domain (one tile position across time, tiles.py:76-98) is not part of the hot path and raises.
"""
import math
import random
from collections import defaultdict
from typing import Dict, Iterator, List, Tuple

from sres.base.util.config import cfg


class BatchSchedule:
    """Iterable of {'start': s, 'end': s + batch_size} over the tiles of one timeslice.  `end` is not clipped
    to ntiles: the loader clips it when it slices (swot/raw.py:163)."""

    def __init__(self, ntiles: int = 0, randomize: bool = False, batch_size: int = None, **_ignored):
        if ntiles <= 0:
            raise AssertionError("Must provide ntiles for TileBatchIterator")
        self.ntiles = int(ntiles)
        self.batch_size = int(batch_size if batch_size is not None else cfg().task.batch_size)
        self.batch_start_idxs: List[int] = list(range(0, self.ntiles, self.batch_size))
        if randomize:
            random.shuffle(self.batch_start_idxs)   # the reference draws the order from Python's RNG, so do we
        self._losses: Dict[str, List[float]] = defaultdict(list)

    def __iter__(self) -> Iterator[Dict[str, int]]:
        return (dict(start=s, end=s + self.batch_size) for s in self.batch_start_idxs)

    def __len__(self) -> int:
        return len(self.batch_start_idxs)

    # per-timeslice loss log ------------------------------------------------------------------------
    def batch_losses(self, ltype: str) -> List[float]:
        return self._losses[ltype]

    def register_loss(self, ltype: str, loss: float) -> None:
        self._losses[ltype].append(float(loss))

    def accumulate_loss(self, ltype: str) -> float:
        """Mean of the losses registered since the last call (nan when there were none), then reset."""
        vals = self._losses.pop(ltype, [])
        return sum(vals) / len(vals) if vals else float("nan")


TileBatchIterator = BatchSchedule


class TileIterator:
    """Factory under the reference's name: `TileIterator.get_iterator(ntiles=..., randomize=...)`."""

    @staticmethod
    def get_iterator(**kwargs) -> BatchSchedule:
        domain = cfg().task.get("batch_domain", "tiles")
        if domain != "tiles":
            raise NotImplementedError(f"sres (B200 build): batch_domain '{domain}' is outside the tile path (tiles only)")
        return BatchSchedule(**kwargs)


class TileGrid:
    """Grid of full-resolution tiles over a region.  A tile covers tile_size * prod(downscale_factors) pixels per
    side; `task.tile_grid` entries >= 0 pin the grid, -1 means "as many as fit the image"."""

    def __init__(self):
        task = cfg().task
        self.origin: Dict[str, int] = dict(task.get("origin", {}))
        self.tile_size: Dict[str, int] = dict(task.tile_size)
        self.upsample_factor: int = math.prod(cfg().model.downscale_factors)
        self._pinned: Dict[str, int] = dict(task.tile_grid)

    def get_tile_size(self, highres: bool = False) -> Dict[str, int]:
        f = self.upsample_factor if highres else 1
        return {d: self.tile_size[d] * f for d in ("x", "y")}

    def get_full_tile_size(self) -> Dict[str, int]:
        return self.get_tile_size(highres=True)

    def get_grid_shape(self, image_shape: Dict[str, int] = None, **_ignored) -> Dict[str, int]:
        if image_shape is None:
            image_shape = cfg().task.get("image_shape", None)
        full = self.get_full_tile_size()
        fits = {d: (image_shape[d] // full[d] if image_shape is not None else 1) for d in ("x", "y")}
        return {d: (self._pinned[d] if self._pinned[d] >= 0 else fits[d]) for d in ("x", "y")}

    def get_active_region(self, **kwargs) -> Dict[str, Tuple[int, int]]:
        full, grid = self.get_full_tile_size(), self.get_grid_shape(**kwargs)
        return {d: (self.origin[d], self.origin[d] + full[d] * grid[d]) for d in ("x", "y")}
