"""Per-CTA timeline of the fused two-convolution launch (forward pair, B = 64, 48 x 48) next to the timing of two launches."""
import ctypes as C, os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sres_b200 import _lib as L
from gpu_util import pack, conv_args
lib = L.lib(); dev = torch.device("cuda:0")
B, H, W = 64, 48, 48
lib.sres_conv_pair_flag_bytes.restype = C.c_size_t
rows = lib.sres_ptl_rows(B, H, W); nt = (rows + 127) // 128
xin = torch.randn(rows, 64, device=dev).bfloat16(); mid = torch.zeros(rows, 64, device=dev, dtype=torch.bfloat16)
out = torch.zeros(rows, 64, device=dev, dtype=torch.bfloat16); part = torch.zeros(nt, 2, 4, 64, device=dev)
w1 = pack(lib, (torch.randn(64, 64, 3, 3) * 0.05).to(dev), 0); w2 = pack(lib, (torch.randn(64, 64, 3, 3) * 0.05).to(dev), 0)
b1 = torch.randn(64, device=dev); flags = torch.zeros(lib.sres_conv_pair_flag_bytes(B, H, W) // 4, dtype=torch.int32, device=dev)
grid = min(nt, lib.sres_device_sm_count())
tl = torch.zeros(grid, 16, device=dev, dtype=torch.int64)
a1 = conv_args(in_bf16=xin, wpack_bf16=w1, bias=b1, out_bf16=mid, B=B, H=H, W=W, n_out=64, epi_flags=L.EPI_RELU)
a2 = conv_args(in_bf16=mid, wpack_bf16=w2, bias=b1, out_bf16=out, pool_part=part, B=B, H=H, W=W, n_out=64, epi_flags=L.EPI_POOL)
st = L.cur_stream()
def timeit(fn, n=40):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
t_pair = timeit(lambda: lib.sres_conv3x3_pair(C.byref(a1), C.byref(a2), L.ptr(flags), st))
def two():
    lib.sres_conv3x3_igemm(C.byref(a1), st); lib.sres_conv3x3_igemm(C.byref(a2), st)
t_two = timeit(two)
print(f"fused pair {t_pair:.1f} us, two launches {t_two:.1f} us")
a1.debug_timeline = tl.data_ptr()
for _ in range(3):
    L.check(lib.sres_conv3x3_pair(C.byref(a1), C.byref(a2), L.ptr(flags), st), "pair"); torch.cuda.synchronize()
t = tl.cpu(); rel = t - t[:, :1]
names = {1: "setup done", 3: "phase-1 weights landed", 4: "phase-1 last MMA issued", 2: "phase-1 MMAs retired (producer)", 5: "phase-1 epilogue done + published",
         6: "phase-2 weights landed", 7: "phase-2 last MMA issued", 8: "phase-2 epilogue done", 9: "exit"}
print("cycles since CTA entry (min / median / max over CTAs)")
for k, n in names.items():
    v = rel[:, k]; print(f"  {n:38s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
for k, n in {12: "producer: flag bookkeeping + spinning", 14: "MMA warp waiting for the epilogue", 15: "MMA warp waiting for TMA"}.items():
    v = t[:, k]; print(f"  {n:38s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
# the same first convolution as a stand-alone launch of the tap-per-MMA kernel, same conditions
tl2 = torch.zeros(grid, 16, device=dev, dtype=torch.int64)
a1.debug_timeline = tl2.data_ptr()
for _ in range(3):
    L.check(lib.sres_conv3x3_igemm(C.byref(a1), st), "conv"); torch.cuda.synchronize()
t = tl2.cpu(); rel = t - t[:, :1]
print("stand-alone conv1 (tap-per-MMA kernel), cycles since CTA entry (min / median / max)")
for k, n in {1: "setup done", 3: "weights landed", 4: "first A tile landed", 6: "first accumulator ready", 5: "last MMA issued", 8: "last store issued", 9: "stores drained", 12: "exit"}.items():
    v = rel[:, k]; print(f"  {n:38s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
for k, n in {13: "epilogue waiting for the tensor core", 14: "MMA warp waiting for the epilogue", 15: "MMA warp waiting for TMA"}.items():
    v = t[:, k]; print(f"  {n:38s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
