"""Bring-up of the N = 192 convolution kernel on a B200: per-flavour error statistics against the tap-per-MMA kernel and fp32
torch, and the per-CTA timeline (clock stamps) of the conv1 flavour.   python tools/bringup_n192.py [B H W]"""
import ctypes as C, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sres_b200 import _lib as L
from gpu_util import to_ptl, from_ptl, pack, conv_args, run_conv, rel_l2, bf16_round, pads_are_zero
lib = L.lib(); dev = torch.device("cuda:0")
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (5, 48, 48)
torch.manual_seed(0)
rows = lib.sres_ptl_rows(B, H, W); nt = (rows + 125) // 126
print(f"B={B} H={H} W={W} rows={rows} tiles={nt} tile_rows={lib.sres_conv_tile_rows(H, W)}")
x = bf16_round(torch.randn(B, 64, H, W)); w = bf16_round(torch.randn(64, 64, 3, 3) * 0.05); bias = torch.randn(64)
ref = F.conv2d(x, w, bias, padding=1)
xin = to_ptl(x.to(dev), torch.bfloat16); wp = pack(lib, w.to(dev), 0); bd = bias.to(dev)
msk = to_ptl(torch.randn(B, 64, H, W).to(dev), torch.bfloat16)
trunk = to_ptl(torch.randn(B, 64, H, W).to(dev), torch.float32)
def run(dbg, **f):
    out16 = torch.full((rows, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    out32 = trunk.clone() if f.get("rmw") else torch.full((rows, 64), float("nan"), device=dev)
    part = torch.full((nt, 2, 4, 64), float("nan"), device=dev)
    kw = dict(in_bf16=xin, wpack_bf16=wp, B=B, H=H, W=W, n_out=64, epi_flags=f.get("epi_flags", 0), debug_flags=dbg)
    if f.get("bias"): kw["bias"] = bd
    if f.get("mask"): kw["mask_bf16"] = msk
    if f.get("o16"): kw["out_bf16"] = out16
    if f.get("o32"): kw["out_f32"] = out32
    if f.get("rmw"): kw["resid_f32"] = out32
    if f.get("part"): kw["pool_part"] = part
    run_conv(lib, conv_args(**kw))
    return out16, out32, part
flav = {"conv1": dict(bias=1, epi_flags=1, o16=1), "conv2": dict(bias=1, epi_flags=2, o16=1, part=1), "dgrad2": dict(mask=1, o16=1),
        "dgrad1": dict(mask=1, epi_flags=4, o32=1, rmw=1, part=1), "plain32": dict(bias=1, o32=1)}
for name, f in flav.items():
    a16, a32, ap = run(128, **f); b16, b32, bp = run(64, **f)
    msg = name
    if f.get("o16"):
        d = (a16.float() - b16.float()).abs()
        msg += f"  bf16: max|d| {d.max().item():.3e} frac differing {(d > 0).float().mean().item():.3e} nan {int(torch.isnan(a16.float()).sum())} pads0 {pads_are_zero(a16, B, H, W)}"
    if f.get("o32"):
        msg += f"  fp32 rel {rel_l2(a32.cpu(), b32.cpu()):.3e} nan {int(torch.isnan(a32).sum())} pads0 {pads_are_zero(a32, B, H, W)}"
    if name == "plain32":
        msg += f"  vs torch fp32 {rel_l2(from_ptl(a32, B, H, W).cpu(), ref):.3e}"
    print(msg)
# timeline of the conv1 flavour
a = L.ConvArgs(); out16 = torch.zeros(rows, 64, device=dev, dtype=torch.bfloat16)
a.in_bf16 = xin.data_ptr(); a.wpack_bf16 = wp.data_ptr(); a.bias = bd.data_ptr(); a.out_bf16 = out16.data_ptr()
a.B, a.H, a.W, a.n_out, a.epi_flags, a.debug_flags = B, H, W, 64, 1, 128
grid = min(nt, lib.sres_device_sm_count())
tl = torch.zeros(grid, 16, device=dev, dtype=torch.int64); a.debug_timeline = tl.data_ptr()
for _ in range(3):
    L.check(lib.sres_conv3x3_igemm(C.byref(a), L.cur_stream()), "conv"); torch.cuda.synchronize()
t = tl.cpu(); rel = t - t[:, :1]
names = {1: "setup done", 2: "producer past pdl_wait", 4: "first A tile landed", 3: "weights row 0 landed", 6: "first accumulator ready", 5: "last MMA issued",
         7: "last accumulator ready (grp 0)", 8: "last store issued (grp 0)", 9: "last store issued (grp 1)", 10: "stores drained", 12: "exit"}
print("per-CTA timeline, cycles since CTA entry (min / median / max over CTAs), tiles per CTA", -(-nt // grid))
for k, n in names.items():
    v = rel[:, k]; print(f"  {n:28s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
for k, n in {13: "epilogue warp 4 waiting for the tensor core", 14: "MMA warp waiting for a free accumulator", 15: "MMA warp waiting for TMA"}.items():
    v = t[:, k]; print(f"  {n:45s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
