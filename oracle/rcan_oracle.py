"""CPU oracle of the RCAN hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (super-resolution-climate_b200/) never does and fails loudly without its
CUDA library.

What it restates (plain fp32 PyTorch on the CPU, the same third-party arithmetic the reference
executes -- the reference pins no torch version, README.md:14; this image has torch 2.11.0):

  edsr_forward            sres/model/edsr/network.py:27-32 (EDSR.forward), sres/model/common/residual.py:50-54
                          (ResBlock: conv, ReLU, conv, .mul(res_scale), += x), sres/model/common/upsample.py:34-66
                          (SPUpsample, same layers as the RCAN Upsampler).  Selected by cfg["name"] == "edsr".
  rcan_forward            sres/model/rcan/network.py:22-27 (RCAN.forward), :44-47 (CALayer),
                          :61-64 (RCAB), :74-77 (ResidualGroup); sres/model/rcan/blocks.py:58-76
                          (Upsampler = [conv F->4F, PixelShuffle(2)] x log2(scale), or conv F->9F +
                          PixelShuffle(3)); sres/model/common/cnn.py:8-9 (3x3 conv, padding k//2).
  state-dict key names    as produced by the reference module tree.
  downsample / upsample   sres/base/util/array.py:72-76, :84-87 (F.interpolate, bicubic,
                          align_corners=False, no antialias).
  l2loss                  sres/controller/stats.py:5-8  (RMSE over the whole batch tensor).
  charbonnier             sres/controller/dual_trainer.py:196-198 (eps = 1e-6, :121).
  l1loss                  NOT in the reference; named by BASELINE.json north_star, kept as a variant.
  apply_network           sres/controller/dual_trainer.py:557-571 (HR batch -> bicubic down -> model).
  train_step              sres/controller/dual_trainer.py:310-323 (zero_grad, forward, loss,
                          backward, Adam.step with torch.optim.Adam defaults, :126).

Pinning: oracle/gen_golden.py imports the UNMODIFIED reference from /root/reference in the build
container (behind stub hydra/omegaconf/xarray modules, oracle/ref_import.py), runs it on seeded
inputs and writes tests/golden/*.npz; tests/test_oracle_golden.py checks this restatement against
those vectors (bit-for-bit on the same torch build).  The reference's own tests hold no golden
vectors or assertions for this path, so the pin is "outputs of the reference
itself run here".
"""
import math
import zlib
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

DEFAULT_MODEL_CFG = dict(  # config/model/rcan-10-20-64.yaml
    name="rcan", nlayers=10, nblocks=20, nfeatures=64, cbottleneck=2, kernel_size=3, bias=True,
    downscale_factors=[2, 2], loss_fn="l2",
)


EDSR_MODEL_CFG = dict(  # config/model/edsr.yaml
    name="edsr", nlayers=16, nfeatures=64, kernel_size=3, bias=True, res_scale=1.0, batch_norm=False,
    downscale_factors=[2, 2], loss_fn="l2",
)


def model_cfg(**over):
    cfg = dict(EDSR_MODEL_CFG if over.get("name") == "edsr" else DEFAULT_MODEL_CFG)
    cfg.update(over)
    return cfg


def is_edsr(cfg) -> bool:
    return cfg.get("name", "rcan") == "edsr"


def scale_of(cfg) -> int:
    return int(math.prod(cfg["downscale_factors"]))  # common.py:24


def upsampler_stages(scale: int) -> List[int]:
    """PixelShuffle factors of the Upsampler (blocks.py:62-73)."""
    if scale & (scale - 1) == 0:
        return [2] * int(round(math.log2(scale)))
    if scale == 3:
        return [3]
    raise NotImplementedError(f"scale {scale}")


def param_shapes(cfg, nchannels_in: int, nchannels_out: int) -> Dict[str, tuple]:
    """Ordered {state_dict key: shape} of the reference RCAN (network.py:9-20) / EDSR (edsr/network.py:14-26)."""
    shapes: Dict[str, tuple] = {}

    def conv(name, cout, cin, ks):
        shapes[name + ".weight"] = (cout, cin, ks, ks)
        shapes[name + ".bias"] = (cout,)

    if is_edsr(cfg):
        Fn, N, k = cfg["nfeatures"], cfg["nlayers"], cfg["kernel_size"]
        conv("head.0", Fn, nchannels_in, k)
        for r in range(N):
            conv(f"body.{r}.body.0", Fn, Fn, k)
            conv(f"body.{r}.body.2", Fn, Fn, k)
        conv(f"body.{N}", Fn, Fn, k)
        for i, f in enumerate(upsampler_stages(scale_of(cfg))):
            conv(f"tail.0.{2 * i}", f * f * Fn, Fn, 3)
        conv("tail.1", nchannels_out, Fn, k)
        return shapes
    Fn, G, R, k = cfg["nfeatures"], cfg["nlayers"], cfg["nblocks"], cfg["kernel_size"]
    red = cfg["cbottleneck"]
    conv("head.0", Fn, nchannels_in, k)
    for g in range(G):
        for r in range(R):
            pre = f"body.{g}.body.{r}.body"
            conv(pre + ".0", Fn, Fn, k)
            conv(pre + ".2", Fn, Fn, k)
            conv(pre + ".3.conv_du.0", Fn // red, Fn, 1)
            conv(pre + ".3.conv_du.2", Fn, Fn // red, 1)
        conv(f"body.{g}.body.{R}", Fn, Fn, k)
    conv(f"body.{G}", Fn, Fn, k)
    for i, f in enumerate(upsampler_stages(scale_of(cfg))):
        conv(f"tail.0.{2 * i}", f * f * Fn, Fn, 3)
    conv("tail.1", nchannels_out, Fn, k)
    return shapes


def make_state_dict(cfg, nchannels_in: int, nchannels_out: int, seed: int = 4456, dtype=torch.float32):
    """Deterministic synthetic weights: every tensor is drawn from its own generator seeded with
    seed + crc32(key), U(-b, b) with b = 1/sqrt(fan_in) (the scale of nn.Conv2d's default init), so
    the reference, the oracle and the CUDA model can be given bit-identical weights by name."""
    sd = {}
    shapes = param_shapes(cfg, nchannels_in, nchannels_out)
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        wshape = shapes[name.rsplit(".", 1)[0] + ".weight"]
        fan_in = wshape[1] * wshape[2] * wshape[3]
        b = 1.0 / math.sqrt(fan_in)
        sd[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
    return sd


def _conv(x, sd, name, pad):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), padding=pad)


def ca_layer(x, sd, pre):
    """CALayer.forward, network.py:44-47."""
    y = F.adaptive_avg_pool2d(x, 1)
    y = F.relu(_conv(y, sd, pre + ".conv_du.0", 0))
    y = torch.sigmoid(_conv(y, sd, pre + ".conv_du.2", 0))
    return x * y


def rcab(x, sd, pre, pad):
    """RCAB.forward, network.py:61-64 (conv, ReLU, conv, CA, += x)."""
    res = F.relu(_conv(x, sd, pre + ".0", pad))
    res = _conv(res, sd, pre + ".2", pad)
    res = ca_layer(res, sd, pre + ".3")
    return res + x


def residual_group(x, sd, g, R, pad):
    """ResidualGroup.forward, network.py:74-77."""
    res = x
    for r in range(R):
        res = rcab(res, sd, f"body.{g}.body.{r}.body", pad)
    res = _conv(res, sd, f"body.{g}.body.{R}", pad)
    return res + x


def rcan_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    """RCAN.forward, network.py:22-27."""
    G, R, k = cfg["nlayers"], cfg["nblocks"], cfg["kernel_size"]
    pad = k // 2
    x = _conv(x, sd, "head.0", pad)
    res = x
    for g in range(G):
        res = residual_group(res, sd, g, R, pad)
    res = _conv(res, sd, f"body.{G}", pad)
    res = res + x
    for i, f in enumerate(upsampler_stages(scale_of(cfg))):
        res = F.pixel_shuffle(_conv(res, sd, f"tail.0.{2 * i}", 1), f)
    return _conv(res, sd, "tail.1", pad)


def edsr_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    """EDSR.forward, edsr/network.py:27-32; ResBlock.forward, common/residual.py:50-54."""
    N, k, rs = cfg["nlayers"], cfg["kernel_size"], cfg.get("res_scale", 1.0)
    pad = k // 2
    x = _conv(x, sd, "head.0", pad)
    res = x
    for r in range(N):
        t = F.relu(_conv(res, sd, f"body.{r}.body.0", pad))
        t = _conv(t, sd, f"body.{r}.body.2", pad).mul(rs)
        res = t + res
    res = _conv(res, sd, f"body.{N}", pad)
    res = res + x
    for i, f in enumerate(upsampler_stages(scale_of(cfg))):
        res = F.pixel_shuffle(_conv(res, sd, f"tail.0.{2 * i}", 1), f)
    return _conv(res, sd, "tail.1", pad)


def model_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg) -> torch.Tensor:
    return edsr_forward(x, sd, cfg) if is_edsr(cfg) else rcan_forward(x, sd, cfg)


# ---------------------------------------------------------------------------------------------
# interpolation and losses
# ---------------------------------------------------------------------------------------------
def downsample(t: torch.Tensor, scale: int) -> torch.Tensor:
    """array.py:72-76 with downsample_mode 'cubic' -> 'bicubic' (array.py:37-41)."""
    return F.interpolate(t, scale_factor=1.0 / scale, mode="bicubic")


def upsample(t: torch.Tensor, scale: int) -> torch.Tensor:
    """array.py:84-87."""
    return F.interpolate(t, scale_factor=scale, mode="bicubic")


def l2loss(prd, tar, squared=False):
    """stats.py:5-8."""
    loss = ((prd - tar) ** 2).mean()
    return loss if squared else torch.sqrt(loss)


def charbonnier(prd, tar, eps=1e-6):
    """dual_trainer.py:196-198."""
    return torch.mean(torch.sqrt(((prd - tar) ** 2) + eps))


def l1loss(prd, tar):
    return torch.mean(torch.abs(prd - tar))


def conform_to_product(prd, tar):
    """dual_trainer.py:200-203."""
    if prd.shape[2] < tar.shape[2] or prd.shape[3] < tar.shape[3]:
        tar = tar[:, :, :prd.shape[2], :prd.shape[3]]
    return tar


def loss_fn(name: str):
    return {"l2": l2loss, "charbonnier": charbonnier, "l1": l1loss}[name]


def single_product_loss(prd, tar, name="l2"):
    """dual_trainer.py:205-212."""
    return loss_fn(name)(prd, conform_to_product(prd, tar))


# ---------------------------------------------------------------------------------------------
# the training / inference step
# ---------------------------------------------------------------------------------------------
def apply_network(hr: torch.Tensor, sd, cfg):
    """dual_trainer.py:557-571: (input, product, target) from an HR batch."""
    lr = downsample(hr, scale_of(cfg))
    return lr, model_forward(lr, sd, cfg), hr


class AdamState:
    """torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, L2 weight decay), restated so the CUDA
    fused Adam can be checked tensor by tensor.  dual_trainer.py:126."""

    def __init__(self, sd, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.wd, self.betas, self.eps, self.t = lr, weight_decay, betas, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}

    def step(self, sd, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for k, p in sd.items():
            g = grads[k]
            if self.wd != 0.0:
                g = g + self.wd * p
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def loss_and_grads(hr: torch.Tensor, sd, cfg, loss_name: Optional[str] = None):
    """forward + loss + backward.  Returns (loss float, product, {key: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    _, prd, tar = apply_network(hr, params, cfg)
    loss = single_product_loss(prd, tar, loss_name or cfg.get("loss_fn", "l2"))
    loss.backward()
    return float(loss.item()), prd.detach(), {k: p.grad for k, p in params.items()}


def forward_backward(lr_in: torch.Tensor, dout: torch.Tensor, sd, cfg):
    """Model forward on an LR batch and the vector-Jacobian product with a given output gradient (autograd of
    network.py:22-27, what mloss.backward() runs below the loss).  Returns (product, {key: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    prd = model_forward(lr_in, params, cfg)
    prd.backward(dout)
    return prd.detach(), {k: p.grad for k, p in params.items()}


def train_step(hr, sd, cfg, adam: AdamState, loss_name: Optional[str] = None):
    """One optimizer step, in place on sd (dual_trainer.py:310-323)."""
    loss, prd, grads = loss_and_grads(hr, sd, cfg, loss_name)
    with torch.no_grad():
        adam.step(sd, grads)
    return loss, prd, grads


def flops_per_tile(cfg, cin, cout, S=48, train=False):
    """Algorithmic FLOPs per LR tile of S x S."""
    if is_edsr(cfg):
        Fn, G, R, red = cfg["nfeatures"], 0, 0, 1
        nbody = 2 * cfg["nlayers"] + 1
    else:
        Fn, G, R, red = cfg["nfeatures"], cfg["nlayers"], cfg["nblocks"], cfg["cbottleneck"]
        nbody = G * (2 * R + 1) + 1
    s = scale_of(cfg)
    ups, px = 0, 1
    for f in upsampler_stages(s):
        ups += f * f * Fn * Fn * px
        px *= f * f
    fwd = 2 * 9 * S * S * (cin * Fn + nbody * Fn * Fn + ups + Fn * cout * s * s) + G * R * 4 * Fn * Fn / red
    return fwd * (3 if train else 1)
