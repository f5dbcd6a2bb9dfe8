"""torchrun --nproc-per-node N tools/dp_check.py : DP-N (B tiles per rank) must equal one GPU on the N*B batch."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))
from sres_b200 import nn as snn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, C, S = 4, 2, 12
torch.manual_seed(7)
kw = dict(nchannels_in=C, nchannels_out=C, nfeatures=64, nlayers=3, nblocks=2, cbottleneck=2, scale=4, device=dev)
model = snn.RCAN(**kw)
dist.broadcast(model.engine.flat, src=0); model.engine.mark_params_changed()
model.enable_data_parallel()
g = torch.Generator().manual_seed(11)
hr_all = torch.randn(world * B, C, S * 4, S * 4, generator=g).to(dev)
hr = hr_all[rank * B:(rank + 1) * B]
for it in range(3):       # repeat: graphs for forward, eager segmented backward + NCCL on the side stream
    for p in model.parameters(): p.grad = None
    prd = model(snn.bicubic_resize(hr, 0.25).requires_grad_(True))
    loss = snn.loss(prd, hr, "l2", dist.group.WORLD)
    loss.backward()
torch.cuda.synchronize()
ref = snn.RCAN(**kw)
ref.engine.flat.copy_(model.engine.flat); ref.engine.mark_params_changed()
prd_r = ref(snn.bicubic_resize(hr_all, 0.25).requires_grad_(True))
loss_r = snn.loss(prd_r, hr_all, "l2")
loss_r.backward()
torch.cuda.synchronize()
gd, gr = model.engine.flat_grad, ref.engine.flat_grad
rel = ((gd - gr).norm() / gr.norm()).item()
same = torch.tensor([float((gd - gr).abs().max())], device=dev)
allg = [torch.zeros_like(gd) for _ in range(world)]
dist.all_gather(allg, gd)
ident = all(torch.equal(allg[0], a) for a in allg)
# ---- ragged global step: ranks hold batches of DIFFERENT sizes and the last rank only rides along (weight 0): the reduced
# gradient must equal the single-GPU gradient of the union of the live batches (global RMSE = sqrt(sum SSE / sum N))
sizes = [B if r % 2 == 0 else B - 1 for r in range(world)]
live = [r < world - 1 or world == 1 for r in range(world)]
offs = [sum(sizes[:r]) for r in range(world)]
hr_r = hr_all[offs[rank]:offs[rank] + sizes[rank]].contiguous()
for p in model.parameters(): p.grad = None
prd = model(snn.bicubic_resize(hr_r, 0.25).requires_grad_(True))
loss_g = snn.loss(prd, hr_r, "l2", dist.group.WORLD, 1.0 if live[rank] else 0.0)
loss_g.backward()
torch.cuda.synchronize()
union = torch.cat([hr_all[offs[r]:offs[r] + sizes[r]] for r in range(world) if live[r]]).contiguous()
for p in ref.parameters(): p.grad = None
loss_u = snn.loss(ref(snn.bicubic_resize(union, 0.25).requires_grad_(True)), union, "l2")
loss_u.backward()
torch.cuda.synchronize()
rel_ragged = ((model.engine.flat_grad - ref.engine.flat_grad).norm() / ref.engine.flat_grad.norm()).item()
ragged_ok = rel_ragged < 1e-4 and abs(loss_g.item() - loss_u.item()) < 1e-5 * abs(loss_u.item())
if rank == 0:
    print(f"RESULT ragged DP-{world}: sizes {sizes}, live {live}: gradient rel {rel_ragged:.3e}, loss {loss_g.item():.6f} vs {loss_u.item():.6f} -> {'OK' if ragged_ok else 'FAIL'}")
# ---- sharded inference through the mirrored trainer: DP-N stitched images == single-GPU stitched images ----------
import numpy as np, tempfile
from sres.base.util.config import ConfigContext
from sres.controller.workflow import WorkflowController
from sres.controller.config import ResultStructure, TSet
tmp = tempfile.mkdtemp()
over = {"model.nlayers": 2, "model.nblocks": 2, "task.batch_size": 8, "task.tile_size": dict(x=12, y=12), "task.tile_order": "corrected",
        "dataset.region": dict(ys=480, xs=576), "dataset.ntimes": 2, "platform.results": tmp}
wc = WorkflowController("sres", dict(task="SSS_SST-tiles-48", dataset="synthetic_1200", platform="local"), seed=1)
wc.initialize("sres", "rcan-10-20-64", **over)
tr = wc.trainer
dist.broadcast(tr.model.engine.flat, src=0); tr.model.engine.mark_params_changed()
images, losses = wc.inference(0, ResultStructure.Image)       # stitched on rank 0 only
tr.world, tr.rank = 1, 0           # the same trainer, unsharded
images1, losses1 = tr.process_image(TSet.Validation, 0, interp_loss=True)
same_img = all(np.array_equal(images[v][k], images1[v][k], equal_nan=True) for v in images for k in images[v]) and (rank != 0 or len(images) == len(images1) > 0)
same_loss = all(abs(losses[v]["model"] - losses1[v]["model"]) < 1e-6 for v in losses)
# ---- data-parallel training through the trainer loop: same shuffles on every rank, replicas stay identical ----------
tr.world, tr.rank = world, rank
out = tr.train(2, True, seed=5, verbose=False)   # range(1, nepochs): one epoch, like the reference
flat = tr.model.engine.flat
allw = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(allw, flat)
same_w = all(torch.equal(allw[0], a) for a in allw) and bool(torch.isfinite(flat).all())
ConfigContext.deactivate()
if rank == 0:
    print(f"RESULT dp{world} trainer.train: loss {out['prediction']:.5f}; replicas identical after training = {same_w}")
    assert same_w
if rank == 0:
    print(f"RESULT dp{world} inference: stitched images identical to unsharded = {same_img}; losses equal = {same_loss}")
    assert same_img and same_loss
if rank == 0:
    print(f"RESULT dp{world}: loss {loss.item():.7f} vs single-GPU {loss_r.item():.7f}; grad rel-L2 vs single-GPU = {rel:.3e}; identical across ranks = {ident}")
    assert abs(loss.item() - loss_r.item()) < 1e-6 and rel < 1e-5 and ident
dist.destroy_process_group()
