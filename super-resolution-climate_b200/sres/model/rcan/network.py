"""RCAN plugin, mirror of sres/model/rcan/network.py:5-27 and the hyper-parameter resolution of
sres/model/common/common.py:9-28 (cfg().model beats keyword arguments; scale = prod(downscale_factors))."""
import math
from typing import Any, Dict

import torch.nn as nn

from sres.base.util.config import cfg
from sres_b200.nn import RCAN as _CudaRCAN

common_parms = dict(nchannels_in=1, nchannels_out=1, nfeatures=64, kernel_size=3, nlayers=16,
                    downscale_factors=[2, 2], bias=True, batch_norm=False, res_scale=1.0, ups_mode="bicubic")


def init_parms(mparms: Dict[str, Any], custom_parms: Dict[str, Any]) -> Dict[str, Any]:
    parms = {pname: cfg().model.get(pname, dval) for pname, dval in common_parms.items()}
    parms["scale"] = math.prod(parms["downscale_factors"])
    for pdict in [mparms, custom_parms]:
        for pname, dval in pdict.items():
            parms[pname] = cfg().model.get(pname, dval)
    return parms


class RCAN(_CudaRCAN):
    def __init__(self, **kwargs):
        parms = init_parms(dict(cbottleneck=2, nblocks=20), kwargs)
        if parms.get("batch_norm", False):
            raise NotImplementedError("sres (B200 build): batch_norm=True is not supported (the reference never enables it)")
        super().__init__(nchannels_in=parms["nchannels_in"], nchannels_out=parms["nchannels_out"],
                         nfeatures=parms["nfeatures"], nlayers=parms["nlayers"], nblocks=parms["nblocks"],
                         cbottleneck=parms["cbottleneck"], kernel_size=parms["kernel_size"], bias=parms["bias"],
                         scale=parms["scale"], device=parms.get("device"))
        self.parms.update({k: v for k, v in parms.items() if k not in self.parms})


def get_model(**config) -> nn.Module:
    return RCAN(**config)
