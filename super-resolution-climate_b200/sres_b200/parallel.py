"""Data parallelism over tile batches: one process per GPU, NCCL all-reduce of the flat gradient
buffer, issued segment by segment on a side stream so it overlaps the rest of backward.

Not in the reference (single process, single device: sres/base/gpu.py:6-15); SURVEY.md 8e.
Tiles are independent units (no inter-tile halo, no BatchNorm), weights and optimizer state are
replicated, the only exchange per step is the gradient sum (+ one scalar for the RMSE loss).
"""
from typing import Optional

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous share of `n_items` for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class SegmentAllReduce:
    def __init__(self, engine, process_group=None, average: bool = False):
        if not dist.is_initialized():
            raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.average = average
        self.on_gpu = torch.device(engine.device).type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=engine.device) if self.on_gpu else None
        self.segments = [engine.segment_params(s) for s in range(engine.num_segments())]

    def _reduce(self, engine, off, cnt):
        bucket = engine.flat_grad[off:off + cnt]
        dist.all_reduce(bucket, group=self.group)
        if self.average:
            bucket.div_(self.world)

    def backward(self, engine, xin, dout, accumulate: bool):
        """Run backward segment by segment; the all-reduce of segment s (NCCL, side stream) overlaps the
        kernels of segment s+1.  Each segment completes the gradients of one contiguous parameter range."""
        if accumulate:
            raise RuntimeError("data-parallel backward needs fresh gradients (optimizer.zero_grad() each step)")
        main = torch.cuda.current_stream(engine.device) if self.on_gpu else None
        for seg, (off, cnt) in enumerate(self.segments):
            engine.backward(xin, dout, accumulate=False, seg_begin=seg, seg_end=seg + 1)
            if self.on_gpu:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    self._reduce(engine, off, cnt)
            else:  # host-side logic only (gloo tests)
                self._reduce(engine, off, cnt)
        if self.on_gpu:
            main.wait_stream(self.comm_stream)
        engine.launches += engine.launches_backward()
