"""Helpers for the GPU parity tests: padded-tile-layout conversion and ctypes plumbing."""
import ctypes as C

import torch

from sres_b200 import _lib as L


def to_ptl(x_nchw: torch.Tensor, dtype) -> torch.Tensor:
    B, Cc, H, W = x_nchw.shape
    out = torch.zeros(B, H + 1, W + 1, Cc, device=x_nchw.device, dtype=dtype)
    out[:, :H, :W, :] = x_nchw.permute(0, 2, 3, 1).to(dtype)
    return out.reshape(B * (H + 1) * (W + 1), Cc).contiguous()


def from_ptl(p: torch.Tensor, B, H, W) -> torch.Tensor:
    return p.reshape(B, H + 1, W + 1, p.shape[-1])[:, :H, :W, :].permute(0, 3, 1, 2).float().contiguous()


def pads_are_zero(p: torch.Tensor, B, H, W) -> bool:
    full = p.reshape(B, H + 1, W + 1, p.shape[-1]).float()
    return bool((full[:, H] == 0).all() and (full[:, :, W] == 0).all())


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def pack(lib, w: torch.Tensor, mode: int, n_rows=64, stride=1, offset=0) -> torch.Tensor:
    out = torch.empty(9 * n_rows * 64, device=w.device, dtype=torch.bfloat16)
    _KEEP.append(w)
    L.check(lib.sres_pack_conv_weights(L.ptr(w), L.ptr(out), mode, n_rows, w.shape[1], w.shape[0], stride, offset, L.cur_stream()), "pack")
    return out


_KEEP = []  # tensors whose device pointers were handed to the C ABI stay alive until the test session ends


def ptr(t):
    """Device pointer of `t` (None -> NULL) that keeps `t` alive: `L.ptr(x.to(dev))` would free the
    temporary before the kernel runs and let the caching allocator hand its memory to the next tensor."""
    if t is not None:
        _KEEP.append(t)
        if len(_KEEP) > 4096:
            torch.cuda.synchronize()
            del _KEEP[:2048]
    return L.ptr(t)


def conv_args(**kw) -> L.ConvArgs:
    a = L.ConvArgs()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            _KEEP.append(v)
            v = v.data_ptr()
        setattr(a, k, v)
    return a


def run_conv(lib, a: L.ConvArgs):
    L.check(lib.sres_conv3x3_igemm(C.byref(a), L.cur_stream()), "sres_conv3x3_igemm")
    torch.cuda.synchronize()
