"""GPU bring-up check of the tcgen05 wgrad kernel against torch.nn.grad.conv2d_weight."""
import argparse, ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "super-resolution-climate_b200"))
from sres_b200 import _lib as L
from bringup_conv import to_ptl

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=2); ap.add_argument("--H", type=int, default=48); ap.add_argument("--W", type=int, default=48)
ap.add_argument("--iters", type=int, default=0)
a = ap.parse_args()
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
lib = L.lib()
lib.sres_conv_wgrad_workspace_bytes.restype = C.c_size_t
B, H, W = a.B, a.H, a.W
x = torch.randn(B, 64, H, W, device=dev).bfloat16().float()
dy = torch.randn(B, 64, H, W, device=dev).bfloat16().float()
xp, dyp = to_ptl(x, torch.bfloat16), to_ptl(dy, torch.bfloat16)
wsb = lib.sres_conv_wgrad_workspace_bytes()
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
dw = torch.full((64, 64, 3, 3), float("nan"), device=dev)
db = torch.full((64,), float("nan"), device=dev)
def run():
    return lib.sres_conv3x3_wgrad(L.ptr(xp), L.ptr(dyp), B, H, W, L.ptr(dw), L.ptr(db), 64, 1, 0, 0, L.ptr(ws), C.c_size_t(wsb), L.cur_stream())
L.check(run(), "wgrad")
torch.cuda.synchronize()
ref = torch.nn.grad.conv2d_weight(x, (64, 64, 3, 3), dy, padding=1)
rdb = dy.sum((0, 2, 3))
print(f"RESULT wgrad B={B} H={H} W={W} rel_l2={((dw-ref).norm()/ref.norm()).item():.4e} max_abs={(dw-ref).abs().max().item():.4e} nan={torch.isnan(dw).sum().item()}")
for t in range(9):
    ky, kx = divmod(t, 3)
    print(f"  tap {t}: rel {((dw[:,:,ky,kx]-ref[:,:,ky,kx]).norm()/ref[:,:,ky,kx].norm()).item():.3e}")
print(f"RESULT dbias rel_l2={((db-rdb).norm()/rdb.norm()).item():.4e}")
if a.iters:
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): run()
    st.record()
    for _ in range(a.iters): run()
    en.record(); torch.cuda.synchronize()
    ms = st.elapsed_time(en) / a.iters
    print(f"TIMING {ms*1000:.1f} us/wgrad  {2.0*B*H*W*64*64*9/ms/1e9:.1f} TFLOP/s (useful)")
