"""World-size-2 checks of the data-parallel host logic on CPU (gloo): segment-wise gradient all-reduce,
global-batch RMSE from per-rank sums, tile-batch sharding.  The CUDA engine is replaced by a stub that
exposes the same surface (flat_grad, segments) -- no compute kernels run here."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


class _StubEngine:
    """Same surface as RcanEngine for SegmentAllReduce: segments partition flat_grad; backward(seg) writes
    a rank- and segment-dependent pattern into that segment only."""
    device = torch.device("cpu")

    def __init__(self, rank, nseg=5, n=1000):
        self.rank, self.nseg, self.n = rank, nseg, n
        self.flat_grad = torch.full((n,), float("nan"))
        self.launches = 0
        bounds = np.linspace(0, n, nseg + 1).astype(int)
        order = list(range(nseg))
        # like the real network: segment 0 = tail of the buffer, last segment = head
        self._spans = [(int(bounds[nseg - 1 - s]), int(bounds[nseg - s] - bounds[nseg - 1 - s])) for s in order]
        self.calls = []

    def num_segments(self):
        return self.nseg

    def segment_params(self, seg):
        return self._spans[seg]

    def launches_backward(self):
        return 7

    def backward(self, x, dout, accumulate, seg_begin=0, seg_end=None):
        for seg in range(seg_begin, seg_end):
            off, cnt = self._spans[seg]
            self.flat_grad[off:off + cnt] = torch.arange(cnt, dtype=torch.float32) * (self.rank + 1) + seg
            self.calls.append(seg)


def _worker(rank, world, port, out_q):
    sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sres_b200.parallel import SegmentAllReduce, dp_schedule, gather_ranges, gather_rows, shard_range
        eng = _StubEngine(rank)
        ddp = SegmentAllReduce(eng, None, average=False)
        ddp.backward(eng, None, None, accumulate=False)
        expect = torch.empty(eng.n)
        for seg, (off, cnt) in enumerate(eng._spans):
            expect[off:off + cnt] = torch.arange(cnt, dtype=torch.float32) * sum(r + 1 for r in range(world)) + seg * world
        ok_sum = bool(torch.equal(eng.flat_grad, expect)) and eng.calls == list(range(eng.nseg)) and eng.launches == 7
        with pytest.raises(RuntimeError):
            ddp.backward(eng, None, None, accumulate=True)
        # fewer, larger all-reduce buckets (SRES_DP_BUCKETS): consecutive segments merge into contiguous ranges, same sums
        for nb in (1, 2, 3):
            eng_b = _StubEngine(rank)
            ddp_b = SegmentAllReduce(eng_b, None, average=False, n_buckets=nb)
            ddp_b.backward(eng_b, None, None, accumulate=False)
            ok_sum = ok_sum and len(ddp_b.buckets) == nb and bool(torch.equal(eng_b.flat_grad, expect)) and eng_b.calls == list(range(eng_b.nseg))
            ok_sum = ok_sum and sum(c for _, _, _, c in ddp_b.buckets) == eng_b.n and ddp_b.buckets[0][0] == 0 and ddp_b.buckets[-1][1] == eng_b.nseg
        # global-batch RMSE: sqrt(sum_r SSE_r / (world * n_local)) == RMSE of the concatenated batch
        g = torch.Generator().manual_seed(1234)
        full = torch.randn(world * 6, 3, generator=g, dtype=torch.float64)
        s, e = shard_range(world * 6, rank, world)
        stat = (full[s:e] ** 2).sum().reshape(1)
        dist.all_reduce(stat)
        rmse = float(torch.sqrt(stat / full.numel()))
        ok_loss = abs(rmse - float(torch.sqrt((full ** 2).mean()))) < 1e-12
        # tile-batch sharding used by ModelTrainer.train (ntiles % batch_size != 0 AND nbatches % world != 0: 101 tiles in
        # batches of 7 = 15 batches, the last one short, and an odd batch left over for two ranks): every rank walks the
        # same number of global steps, every batch is trained exactly once, the rank without a batch in the ragged last
        # step rides along with weight 0
        batches = list(range(0, 101, 7))
        sched = dp_schedule(len(batches), rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, sched)
        live = sorted(i for part in gathered for i, w in part if w == 1.0)
        ok_shard = live == list(range(len(batches))) and len(set(map(len, gathered))) == 1
        ok_shard = ok_shard and all(0 <= i < len(batches) for part in gathered for i, _ in part)
        ok_shard = ok_shard and sum(1 for part in gathered for _, w in part if w == 0.0) == (-len(batches)) % world
        # global-batch RMSE with UNEQUAL per-rank batches: sum and element count are both reduced (weight-0 ranks add 0, 0)
        sizes = [7 * 3, 3 * 3] + [0] * (world - 2)
        vals = torch.randn(sum(sizes), generator=torch.Generator().manual_seed(5), dtype=torch.float64)
        lo = sum(sizes[:rank])
        mine_v = vals[lo:lo + sizes[rank]]
        st2 = torch.stack([(mine_v ** 2).sum(), torch.tensor(float(mine_v.numel()), dtype=torch.float64)])
        dist.all_reduce(st2)
        ok_loss = ok_loss and abs(float(torch.sqrt(st2[0] / st2[1])) - float(torch.sqrt((vals ** 2).mean()))) < 1e-12
        # inference sharding: ranks hold contiguous tile ranges (uneven, one rank may be empty), gather restores tile order
        tiles = torch.arange(7 * 2 * 3 * 3, dtype=torch.float32).reshape(7, 2, 3, 3)
        s7, e7 = shard_range(7, rank, world)
        ok_gather = bool(torch.equal(gather_ranges(tiles[s7:e7], 7), tiles))
        counts = [5, 0] if world == 2 else [5] + [0] * (world - 1)
        local = tiles[:5] if rank == 0 else None
        ok_gather = ok_gather and bool(torch.equal(gather_rows(local, counts, (2, 3, 3), torch.device("cpu")), tiles[:5]))
        with pytest.raises(ValueError):
            gather_rows(tiles[:1], counts if rank else [4] + counts[1:], (2, 3, 3), torch.device("cpu"))
        out_q.put((rank, ok_sum, ok_loss, ok_shard and ok_gather))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gradient_allreduce_and_global_loss():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
