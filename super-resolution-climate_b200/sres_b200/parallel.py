"""Data parallelism over tile batches: one process per GPU, NCCL all-reduce of the flat gradient
buffer, issued bucket by bucket (groups of backward segments) on a side stream so it overlaps the rest of backward.

Not in the reference (single process, single device: sres/base/gpu.py:6-15); SURVEY.md 8e.
Tiles are independent units (no inter-tile halo, no BatchNorm), weights and optimizer state are
replicated, the only exchange per step is the gradient sum (+ one scalar for the RMSE loss).
"""
from typing import Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous share of `n_items` for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def dp_schedule(n_batches: int, rank: int, world: int):
    """Which tile batch `rank` runs at every global step of a timeslice, and with which loss weight.

    Every global step consumes `world` consecutive batches of the (shuffled) order; the last step of a timeslice may
    have fewer than `world` batches left.  The ranks without a batch of their own still have to take part in the
    loss and gradient collectives, so they re-run one of the step's batches with weight 0 (it then adds nothing to the
    global sum, the global element count or the gradient) instead of dropping the trailing batches.  Returns a list
    of (batch_index, weight) with the same length on every rank; over all ranks every batch appears exactly once
    with weight 1."""
    steps = []
    for i0 in range(0, n_batches, world):
        left = min(world, n_batches - i0)
        steps.append((i0 + rank, 1.0) if rank < left else (i0 + rank % left, 0.0))
    return steps


def gather_rows(local: Optional[torch.Tensor], counts: Sequence[int], row_shape: Sequence[int], device, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks in rank order on every rank: rank r contributes `counts[r]` rows of shape
    `row_shape` (fp32; `local` may be None / empty when counts[rank] == 0).  One all_gather of blocks padded to the
    longest.  Inference sharding: ranks run the forward on contiguous tile ranges, the gathered
    products are stitched on rank 0."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if len(counts) != world:
        raise ValueError("gather_rows: one count per rank")
    n_local = 0 if local is None else int(local.shape[0])
    if n_local != counts[rank]:
        raise ValueError(f"gather_rows: rank {rank} holds {n_local} rows, expected {counts[rank]}")
    pad = torch.zeros((max(counts),) + tuple(row_shape), dtype=torch.float32, device=device)
    if n_local:
        pad[:n_local] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, counts)], dim=0)


def gather_ranges(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of `shard_range` for an (n_total, ...) tensor."""
    world = dist.get_world_size(group)
    counts = [e - s for s, e in (shard_range(n_total, r, world) for r in range(world))]
    return gather_rows(local, counts, local.shape[1:], local.device, group)


class SegmentAllReduce:
    def __init__(self, engine, process_group=None, average: bool = False, n_buckets: Optional[int] = None):
        if not dist.is_initialized():
            raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.average = average
        if n_buckets is None:
            import os
            # Measured on 8 x B200 (profiles/r02_dp8_buckets.txt): one all-reduce per backward segment (12 NCCL launches whose
            # CTAs displace the persistent tensor-core kernels' CTAs, and 12 graph replays) 28.65 ms per step; ONE bucket after
            # the whole backward (65 MB, exposed) 27.95 ms; TWO buckets (the first overlaps the second half of backward)
            # 27.81 ms against 27.07 ms on one GPU of the same box.
            n_buckets = int(os.environ.get("SRES_DP_BUCKETS", "2"))
        self.on_gpu = torch.device(engine.device).type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=engine.device) if self.on_gpu else None
        self.segments = [engine.segment_params(s) for s in range(engine.num_segments())]
        self.buckets = self._make_buckets(self.segments, n_buckets)
        self.trace = None   # development aid (tools/dp_timeline.py): list of (label, stream name, event) when not None

    @staticmethod
    def _make_buckets(segments, n_buckets):
        """Group consecutive backward segments into `n_buckets` all-reduce buckets of roughly equal parameter count.
        Consecutive segments complete adjacent, descending parameter ranges (segment 0 = the end of the flat buffer), so a
        group is one contiguous range: (seg_begin, seg_end, offset, count).  None / 0 = one bucket per segment."""
        nseg = len(segments)
        if not n_buckets or n_buckets >= nseg:
            return [(s, s + 1, off, cnt) for s, (off, cnt) in enumerate(segments)]
        total = sum(c for _, c in segments)
        buckets, begin, acc = [], 0, 0
        for s, (off, cnt) in enumerate(segments):
            acc += cnt
            left = n_buckets - len(buckets) - 1          # buckets still to open after this one
            if s == nseg - 1 or (acc >= total * (len(buckets) + 1) / n_buckets and nseg - 1 - s >= left):
                lo = min(o for o, _ in segments[begin:s + 1])
                hi = max(o + c for o, c in segments[begin:s + 1])
                if hi - lo != sum(c for _, c in segments[begin:s + 1]):
                    raise RuntimeError("backward segments of one bucket do not form a contiguous parameter range")
                buckets.append((begin, s + 1, lo, hi - lo))
                begin = s + 1
        return buckets

    def _mark(self, label, stream, name):
        if self.trace is not None and self.on_gpu:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((label, name, ev))

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
        self._mark("bwd begin", main, "main")
        for seg, seg_end, off, cnt in self.buckets:
            engine.backward(xin, dout, accumulate=False, seg_begin=seg, seg_end=seg_end)
            if self.on_gpu:
                ev = torch.cuda.Event()
                ev.record(main)
                self._mark(f"seg {seg} done", main, "main")
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    self._mark(f"ar {seg} begin", self.comm_stream, "comm")
                    self._reduce(engine, off, cnt)
                    self._mark(f"ar {seg} end", self.comm_stream, "comm")
            else:  # host-side logic only (gloo tests)
                self._reduce(engine, off, cnt)
        if self.on_gpu:
            main.wait_stream(self.comm_stream)
            self._mark("bwd joined", main, "main")
        engine.launches += engine.launches_backward()
