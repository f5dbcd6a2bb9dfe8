"""EDSR plugin, mirror of sres/model/edsr/network.py:7-32 (ResBlock: sres/model/common/residual.py:30-54,
SPUpsample: sres/model/common/upsample.py:34-66) with the hyper-parameter resolution of
sres/model/common/common.py:9-28.  Runs on the same tensor-core convolution kernels as RCAN."""
from typing import Any, Dict

import torch.nn as nn

from sres.model.rcan.network import init_parms
from sres_b200.nn import EDSR as _CudaEDSR


class EDSR(_CudaEDSR):
    def __init__(self, **kwargs):
        parms: Dict[str, Any] = init_parms({}, kwargs)
        if parms.get("batch_norm", False):
            raise NotImplementedError("sres (B200 build): batch_norm=True is not supported (the reference never enables it)")
        super().__init__(nchannels_in=parms["nchannels_in"], nchannels_out=parms["nchannels_out"],
                         nfeatures=parms["nfeatures"], nlayers=parms["nlayers"], kernel_size=parms["kernel_size"],
                         bias=parms["bias"], scale=parms["scale"], res_scale=parms["res_scale"], device=parms.get("device"))
        self.parms.update({k: v for k, v in parms.items() if k not in self.parms})


def get_model(**config) -> nn.Module:
    return EDSR(**config)
