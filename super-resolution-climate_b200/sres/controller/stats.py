"""Loss function, mirror of sres/controller/stats.py:5-8 (RMSE over the whole batch tensor)."""
import torch

from sres_b200 import nn as _snn


def l2loss(prd: torch.Tensor, tar: torch.Tensor, squared=False) -> torch.Tensor:
    return _snn.l2loss(prd, tar, squared)
