"""Event timeline of one data-parallel training step (BASELINE config 3) -- where do the N-GPU steps lose time against one GPU?
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dp_timeline.py
Per rank: forward, the scalar loss all-reduce, every backward segment and its gradient all-reduce (NCCL, side stream), the
final join and the optimizer, as CUDA-event times relative to the start of the step (mean over the timed steps).  Rank 0 prints
the slowest and fastest rank per mark, and the step time with the loss all-reduce / the gradient all-reduces switched off."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(4456)
model = snn.RCAN(nchannels_in=2, nchannels_out=2, nfeatures=64, nlayers=10, nblocks=20, cbottleneck=16, scale=4, device=dev)
if world > 1:
    model.enable_data_parallel(); dist.broadcast(model.engine.flat, src=0); model.engine.mark_params_changed()
opt = snn.FusedAdam(model, lr=1e-4)
hr = [torch.randn(64, 2, 192, 192, device=dev) for _ in range(4)]
group = dist.group.WORLD if world > 1 else None
main = torch.cuda.current_stream()

def step(i, marks=None, use_group=True):
    def mark(label):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True); ev.record(main); marks.append((label, "main", ev))
    mark("step begin")
    opt.zero_grad()
    x = snn.bicubic_resize(hr[i % 4], 0.25)
    prd = model(x.requires_grad_(True))
    mark("forward done")
    loss = snn.loss(prd, hr[i % 4], "l2", group if use_group else None)
    mark("loss (+ scalar all-reduce) done")
    loss.backward()
    mark("backward done")
    opt.step()
    mark("optimizer done")
    return loss

def timed(n, **kw):
    for i in range(3): step(i, **kw)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(i, **kw)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

base = timed(20)
no_sse = timed(20, use_group=False) if world > 1 else base
# timeline of 5 steps
acc = {}
for i in range(5):
    marks = []
    if model.ddp is not None: model.ddp.trace = marks
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    step(i, marks)
    torch.cuda.synchronize()
    t0 = marks[0][2]
    for label, sname, ev in marks[1:]:
        acc.setdefault(label, []).append(t0.elapsed_time(ev))
if model.ddp is not None: model.ddp.trace = None
labels = list(acc.keys())
mine = torch.tensor([sum(acc[l]) / len(acc[l]) for l in labels], device=dev, dtype=torch.float64)
if world > 1:
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    allr = torch.stack(allr).cpu()
else:
    allr = mine.cpu()[None]
if rank == 0:
    print(f"world {world}: step {base:.3f} ms (max over ranks, 20 steps); without the scalar loss all-reduce {no_sse:.3f} ms")
    print(f"{'mark':40s} {'min rank':>10s} {'max rank':>10s}   (ms since step begin, mean of 5 isolated steps)")
    for j, l in enumerate(labels):
        print(f"{l:40s} {float(allr[:, j].min()):10.3f} {float(allr[:, j].max()):10.3f}")
if world > 1:
    dist.destroy_process_group()
