"""Per-CTA timeline of the fused BACKWARD pair (dgrad(conv2) + ReLU mask -> dgrad(conv1) + fp32 read-modify-write + sum g*t2) at
B = 64, 48 x 48, with the epilogue operands (mask T1, T2) cold (rotating buffers) or L2-resident."""
import ctypes as C, os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sres_b200 import _lib as L
from gpu_util import pack, conv_args
lib = L.lib(); dev = torch.device("cuda:0")
B, H, W = 64, 48, 48
lib.sres_conv_pair_flag_bytes.restype = C.c_size_t
rows = lib.sres_ptl_rows(B, H, W); nt = (rows + 127) // 128
NB = 10
dt2 = torch.randn(rows, 64, device=dev).bfloat16(); dt1 = torch.zeros(rows, 64, device=dev, dtype=torch.bfloat16)
t1 = [torch.randn(rows, 64, device=dev).bfloat16() for _ in range(NB)]
t2 = [torch.randn(rows, 64, device=dev).bfloat16() for _ in range(NB)]
g32 = torch.randn(rows, 64, device=dev); part = torch.zeros(nt, 2, 4, 64, device=dev)
w1 = pack(lib, (torch.randn(64, 64, 3, 3) * 0.05).to(dev), 1); w2 = pack(lib, (torch.randn(64, 64, 3, 3) * 0.05).to(dev), 1)
flags = torch.zeros(lib.sres_conv_pair_flag_bytes(B, H, W) // 4, dtype=torch.int32, device=dev)
grid = min(nt, lib.sres_device_sm_count())
st = L.cur_stream()
def args(i):
    a1 = conv_args(in_bf16=dt2, wpack_bf16=w2, out_bf16=dt1, mask_bf16=t1[i], B=B, H=H, W=W, n_out=64, epi_flags=0)
    a2 = conv_args(in_bf16=dt1, wpack_bf16=w1, out_f32=g32, resid_f32=g32, mask_bf16=t2[i], pool_part=part, B=B, H=H, W=W, n_out=64, epi_flags=L.EPI_DOT)
    return a1, a2
A = [args(i) for i in range(NB)]
assert lib.sres_conv_pair_supported(C.byref(A[0][0]), C.byref(A[0][1]))
def timeit(rot, n=40):
    for i in range(5): lib.sres_conv3x3_pair(C.byref(A[i % NB if rot else 0][0]), C.byref(A[i % NB if rot else 0][1]), L.ptr(flags), st)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        k = i % NB if rot else 0
        lib.sres_conv3x3_pair(C.byref(A[k][0]), C.byref(A[k][1]), L.ptr(flags), st)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
print(f"backward pair: {timeit(False):.1f} us with L2-resident mask / T2, {timeit(True):.1f} us with cold ones (rotating over {NB} x 2 buffers)")
for rot in (False, True):
    tl = torch.zeros(grid, 16, device=dev, dtype=torch.int64)
    for i in range(4):
        a1, a2 = A[(i + 3) % NB if rot else 0]
        a1.debug_timeline = tl.data_ptr()
        L.check(lib.sres_conv3x3_pair(C.byref(a1), C.byref(a2), L.ptr(flags), st), "pair"); torch.cuda.synchronize()
        a1.debug_timeline = None
    t = tl.cpu(); rel = t - t[:, :1]
    print("cold operands" if rot else "L2-resident operands", "- cycles (min / median / max over CTAs)")
    for k, n in {3: "phase-1 weights landed", 4: "phase-1 last MMA issued", 5: "phase-1 epilogue done + published", 6: "phase-2 weights landed", 7: "phase-2 last MMA issued", 8: "phase-2 epilogue done", 9: "exit"}.items():
        v = rel[:, k]; print(f"  {n:46s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
    for k, n in {12: "producer: flag bookkeeping + spinning", 14: "MMA warp waiting for the epilogue", 15: "MMA warp waiting for TMA"}.items():
        v = t[:, k]; print(f"  {n:46s} {int(v.min()):8d} {int(v.median()):8d} {int(v.max()):8d}")
