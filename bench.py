#!/usr/bin/env python
"""Benchmark of the RCAN hot path (BASELINE.json metric: RCAN train tiles/s, 48x48 2-ch, x4).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # data parallel, one rank per GPU

One step = one optimizer step of RCAN-full (10 groups x 20 RCABs, 64 features, reduction 16, x4) on a
batch of 64 synthetic 2-channel 48x48 LR tiles per GPU: bicubic down of the HR batch, forward, RMSE
loss, backward, fused Adam.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200"))

WORKLOAD = dict(nlayers=10, nblocks=20, nfeatures=64, cbottleneck=16, downscale_factors=[2, 2])
TILE, CH, SCALE, BATCH = 48, 2, 4, 64
METRIC = "rcan_train_tiles_per_s"


def flops_per_tile(train=True):
    """Algorithmic FLOPs per LR tile: 2*MACs of all convolutions + CA MLP."""
    Fn, G, R, red, S, s = 64, WORKLOAD["nlayers"], WORKLOAD["nblocks"], WORKLOAD["cbottleneck"], TILE, SCALE
    ups = 4 * Fn * Fn * 1 + 4 * Fn * Fn * 4
    fwd = 2 * 9 * S * S * (CH * Fn + (G * (2 * R + 1) + 1) * Fn * Fn + ups + Fn * CH * s * s) + G * R * 4 * Fn * Fn / red
    return fwd * (3 if train else 1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], src="measured")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.max_mhz, self._halt = gpu_index, [], set(), None, threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def synth_hr(B, C, S, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, S, S, generator=g)


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyTorch CPU path
# ---------------------------------------------------------------------------------------------------
def cpu_train_tiles_per_s(batch, steps, warmup, threads):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rcan_oracle as O
    torch.set_num_threads(threads)
    cfg = O.model_cfg(**WORKLOAD)
    sd = O.make_state_dict(cfg, CH, CH)
    adam = O.AdamState(sd, lr=1e-4)
    hr = synth_hr(batch, CH, TILE * SCALE, 4456)
    for _ in range(warmup):
        O.train_step(hr, sd, cfg, adam, "l2")
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(hr, sd, cfg, adam, "l2")
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    batch = 2
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    tps, spstep = cpu_train_tiles_per_s(batch, steps, warmup, threads)
    sample = f"{steps} timed steps of batch {batch} (of the {BATCH}-tile workload batch), fp32 PyTorch {torch.__version__} CPU, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tiles/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "RCAN-full (10x20 RCAB, 64 feats, reduction 16) x4 train step, 2-ch 48x48 LR tiles; CPU sample batch 2",
                   "tile": TILE, "channels": CH, "scale": SCALE, "batch_per_gpu": BATCH},
        "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from sres_b200 import _lib as L
    from sres_b200 import nn as snn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE {world}; using {world}", file=sys.stderr)
    pk = peaks()
    torch.manual_seed(4456)
    # the public API: the reference's own controller surface (sres.controller.dual_trainer.ModelTrainer)
    from sres.base.util.config import ConfigContext
    from sres.controller.dual_trainer import ModelTrainer
    ConfigContext.set_defaults(task="SSS_SST-tiles-48", dataset="synthetic_1200", platform="local")
    cc = ConfigContext.activate_global("sres", model="rcan-10-20-64", **{
        "model.cbottleneck": WORKLOAD["cbottleneck"], "task.batch_size": BATCH, "task.lr": 1e-4, "pipeline.gpu": local})
    trainer = ModelTrainer(cc)
    model, opt = trainer.model, trainer.optimizer
    model.train()
    if world > 1:
        dist.broadcast(model.engine.flat, src=0)
        model.engine.mark_params_changed()
    eng = model.engine
    B, S = BATCH, TILE * SCALE
    nbuf = 4
    host = [synth_hr(B, CH, S, 4456 + 17 * rank + i).pin_memory() for i in range(nbuf)]
    resident = [h.to(dev) for h in host]
    step = trainer.train_step   # zero_grad, bicubic down, forward, RMSE, backward, Adam (dual_trainer.py:310-323)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    per_step_launches = eng.launches_forward(TILE, TILE) + eng.launches_backward() + 1 + 4 + 1  # + bicubic, loss(sum x2, value, grad), adam

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    for i in range(args.warmup):
        step(resident[i % nbuf])
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(resident[i % nbuf])
    e1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers through the public trainer-style call, H2D + loss D2H inside the timed region
    for i in range(2):
        step(host[i % nbuf]).item()      # pinned host batch in, python float out
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(host[i % nbuf]).item()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)

    # ---- inference (BASELINE config 4, model part): batched forward without gradients, output megapixels/s ------
    model.eval()
    with torch.no_grad():
        for i in range(3):
            trainer.apply_network(resident[i % nbuf])
        sync_all()
        e0.record()
        for i in range(args.steps):
            trainer.apply_network(resident[i % nbuf])
        e1.record()
        sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    infer_tiles_s = world * B * args.steps / (float(t.item()) / 1e3)
    model.train()

    # ---- roofline of the dominant kernel: the 64->64 tensor-core conv, timed live with CUDA events ---
    import ctypes as C
    rows = B * (TILE + 1) * (TILE + 1)
    nact = 24  # rotate over 24 activation buffers (24 x 19.7 MB >> L2) so the operand is not L2-resident
    acts = [torch.randn(rows, 64, device=dev).bfloat16() for _ in range(nact)]
    outb = torch.empty(rows, 64, device=dev, dtype=torch.bfloat16)
    wpack = torch.randn(9 * 64 * 64, device=dev).bfloat16()
    bias = torch.zeros(64, device=dev)
    ca = L.ConvArgs()
    ca.wpack_bf16, ca.bias, ca.out_bf16 = wpack.data_ptr(), bias.data_ptr(), outb.data_ptr()
    ca.B, ca.H, ca.W, ca.n_out, ca.epi_flags = B, TILE, TILE, 64, L.EPI_RELU
    lib = L.lib()
    st = L.cur_stream()
    reps = 48
    for i in range(8):
        ca.in_bf16 = acts[i % nact].data_ptr()
        L.check(lib.sres_conv3x3_igemm(C.byref(ca), st), "conv")
    e0.record()
    for i in range(reps):
        ca.in_bf16 = acts[i % nact].data_ptr()
        lib.sres_conv3x3_igemm(C.byref(ca), st)
    e1.record()
    torch.cuda.synchronize()
    conv_ms = e0.elapsed_time(e1) / reps
    conv_flops = 2.0 * B * TILE * TILE * 64 * 64 * 9
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on a bounded sample ---------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        tps, _ = cpu_train_tiles_per_s(2, 3, 1, threads)
        cpu = {"value": tps, "unit": "tiles/s", "cores": threads, "kind": "port",
               "sample": "3 timed steps of batch 2 of the same RCAN-full x4 train step (oracle/rcan_oracle.py, fp32 PyTorch CPU)"}

    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of the same kernel/shape
    traffic = None
    try:
        m = json.load(open(os.path.join(ROOT, "profiles", "r01_v4_conv_ncu_metrics.json")))
        conv = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        traffic = sum(float(m[k]["value"].replace(",", "")) * conv[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        pass
    step_tflops = value / world * flops_per_tile(True) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "RCAN-full (10 groups x 20 RCABs, 64 feats, reduction 16) x4 train step: bicubic down, fwd, RMSE, bwd, Adam",
                   "tile": TILE, "channels": CH, "scale": SCALE, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "operands": "bf16 tensor-core operands, fp32 accumulate / trunk / grads",
                   "l2": "per-step working set ~12 GB of saved activations >> 126 MB L2; inputs rotate over 4 buffers"},
        "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": B * CH * S * S * 4, "d2h_bytes_per_step": 4},
        "gpu_launches": per_step_launches * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv3x3_igemm_kernel<64,false,17> (64->64 conv + bias + ReLU, bf16 out: RCAB conv1, B=64, 48x48)", "achieved": achieved,
                     "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_burst"], "traffic": traffic,
                     "traffic_note": "DRAM bytes per launch (ncu --set full, profiles/r01_v4_conv_ncu_metrics.json); algorithmic 39.3 MB "
                                     "(19.7 in + 19.7 out): the input is read once, the output is still in L2 when the kernel ends",
                     "peak_source": pk["src"] + " burst (kernel timed alone)", "us_per_launch": conv_ms * 1e3,
                     "step_tflops_per_gpu": step_tflops, "step_frac_of_sustained": step_tflops / pk["tflops_sustained"]},
        "cpu_baseline": cpu,
        "inference": {"value": infer_tiles_s * (TILE * SCALE) ** 2 / 1e6, "unit": "output MP/s", "tiles_per_s": infer_tiles_s,
                      "what": "bicubic down + RCAN-full forward (no grad) on resident 64-tile batches, all GPUs",
                      "frac_of_tensor_peak": infer_tiles_s / world * flops_per_tile(False) / 1e12 / pk["tflops_sustained"]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
