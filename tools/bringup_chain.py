"""Timing and per-CTA phase breakdown of the image-resident RCAB chain launch (rcab_chain.cu) at B = 64, 48 x 48, 20 blocks,
next to the same blocks run as fused pair + channel-attention launches.  python tools/bringup_chain.py [B] [nblocks]"""
import ctypes as C, os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "super-resolution-climate_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sres_b200 import _lib as L
from gpu_util import conv_args
lib = L.lib(); dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H = W = 48; hid = 4
RP = (H + 1) * (W + 1); rows = B * RP; KW = 64 * 64 * 9
stride = 2 * (KW + 64) + hid * 64 + hid + 64 * hid + 64
params = (torch.randn(nb, stride, device=dev) * 0.03).contiguous()
wpack = (torch.randn(nb * 2, KW, device=dev) * 0.03).bfloat16()
x0 = torch.randn(rows, 64, device=dev)
xb = torch.zeros(nb + 1, rows, 64, device=dev, dtype=torch.bfloat16); xb[0] = x0.bfloat16()
t1 = torch.zeros(nb, rows, 64, device=dev, dtype=torch.bfloat16); t2 = torch.zeros_like(t1)
xf = torch.zeros(rows, 64, device=dev); mean = torch.zeros(nb, B, 64, device=dev); sv = torch.zeros_like(mean)
scratch = torch.zeros(lib.sres_rcab_chain_scratch_bytes(B) // 4, device=dev)
a = L.ChainArgs()
a.xb_bf16, a.t1_bf16, a.t2_bf16 = xb.data_ptr(), t1.data_ptr(), t2.data_ptr()
a.wpack_bf16, a.params, a.x_in_f32, a.x_f32 = wpack.data_ptr(), params.data_ptr(), x0.data_ptr(), xf.data_ptr()
a.save_mean, a.save_s, a.scratch = mean.data_ptr(), sv.data_ptr(), scratch.data_ptr()
a.rcab_stride, a.save_stride = stride, B * 64
a.B, a.H, a.W, a.n_blocks, a.hidden = B, H, W, nb, hid
a.xb_first, a.xb_ring, a.xb_count = 0, 0, nb + 1
a.t_first, a.t_fixed, a.t_count = 0, 0, nb
st = L.cur_stream()
def timeit(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
t_chain = timeit(lambda: L.check(lib.sres_rcab_chain_fwd(C.byref(a), st), "chain"))
# the tile-parallel path on the same buffers
lib.sres_conv_pair_flag_bytes.restype = C.c_size_t
nt = lib.sres_conv_mtiles(B, H, W)
part = torch.zeros(nt, 2, 4, 64, device=dev)
flags = torch.zeros(2, lib.sres_conv_pair_flag_bytes(B, H, W) // 4, dtype=torch.int32, device=dev)
def tile_parallel():
    for r in range(nb):
        pr = params[r]
        o = 2 * (KW + 64)
        a1 = conv_args(in_bf16=xb[r], wpack_bf16=wpack[2 * r], bias=pr[KW:KW + 64], out_bf16=t1[r], B=B, H=H, W=W, n_out=64, epi_flags=L.EPI_RELU)
        a2 = conv_args(in_bf16=t1[r], wpack_bf16=wpack[2 * r + 1], bias=pr[2 * KW + 64:2 * KW + 128], out_bf16=t2[r], pool_part=part, B=B, H=H, W=W, n_out=64, epi_flags=L.EPI_POOL)
        lib.sres_conv3x3_pair(C.byref(a1), C.byref(a2), L.ptr(flags[r & 1]), st)
        lib.sres_ca_apply_fwd(L.ptr(t2[r]), L.ptr(part), None, L.ptr(pr[o:o + hid * 64]), L.ptr(pr[o + hid * 64:o + hid * 64 + hid]),
                              L.ptr(pr[o + hid * 64 + hid:o + hid * 64 + hid + 64 * hid]), L.ptr(pr[o + hid * 64 + hid + 64 * hid:stride]), hid,
                              L.ptr(x0 if r == 0 else xf), L.ptr(xf), L.ptr(xb[r + 1]), L.ptr(mean[r]), L.ptr(sv[r]), B, H, W, st)
t_tp = timeit(tile_parallel)
print(f"B={B} {nb} RCABs: chain launch {t_chain:.1f} us = {t_chain / nb:.2f} us per RCAB;  pair + channel-attention launches {t_tp:.1f} us = {t_tp / nb:.2f} us per RCAB")
grid = B * 2
tl = torch.zeros(grid, 16, device=dev, dtype=torch.int64)
a.debug_timeline = tl.data_ptr()
for _ in range(2):
    L.check(lib.sres_rcab_chain_fwd(C.byref(a), st), "chain"); torch.cuda.synchronize()
t = tl.cpu().double()
if os.environ.get("SRES_CHAIN_OVERLAP", "0") != "0":
    print(f"cycles per CTA over the launch: median {(t[:, 9] - t[:, 0]).median():.0f}; per RCAB {(t[:, 9] - t[:, 0]).median() / nb:.0f}")
    for rk in (0, 1):
        sel = t[rk::2]
        print(f" cluster rank {rk}: cycles per RCAB, median over CTAs")
        for i, n in ((1, "conv2 phase"), (2, "conv2 + pool exchange + MLP"), (3, "streaming x += t2*s (stream warp 2)"), (4, "gated conv1 of the next block"),
                     (5, "producer waiting for the stream warps"), (6, "whole overlapped phase incl. barrier")):
            print(f"   {n:40s} {sel[:, i].median() / nb:9.0f}")
    sys.exit(0)
names = ["conv1 phase (epilogue thread)", "S1 wait (T1 stored, cluster)", "conv2 phase", "pool exchange + S2", "gate MLP", "apply x += t2*s", "S3 wait"]
tot = (t[:, 9] - t[:, 0])
print(f"cycles per CTA over the launch: median {tot.median():.0f} (min {tot.min():.0f}, max {tot.max():.0f}); per RCAB {tot.median() / nb:.0f}")
for rk in (0, 1):
    sel = t[rk::2]
    print(f" cluster rank {rk} ({10 if rk == 0 else 9} tiles per convolution): cycles per RCAB, median over CTAs")
    for i, n in enumerate(names):
        print(f"   {n:32s} {sel[:, 1 + i].median() / nb:9.0f}")
    for i, n in ((10, "MMA warp waiting for weights"), (11, "MMA warp waiting for the epilogue"), (12, "MMA warp waiting for TMA")):
        print(f"   {n:32s} {sel[:, i].median() / nb:9.0f}")
