"""Model factory, mirror of sres/model/manager.py:42-96 (SRModels): the plugin boundary.
`get_model()` imports sres.model.<cfg().model.name>.network and calls its get_model(**config)."""
import importlib
from typing import List

import torch
import torch.nn as nn

from sres.base.util.config import cfg


class SRModels:

    def __init__(self, device: torch.device):
        self.model_name = cfg().model.name
        self.device = device
        self.target_variables = list(cfg().task.target_variables)
        self._dataset = None
        self.model_config = dict(nchannels_in=len(cfg().task.input_variables),
                                 nchannels_out=len(cfg().task.target_variables), device=device)

    def get_dataset(self):
        if self._dataset is None:
            from sres.data.batch import BatchDataset
            self._dataset = BatchDataset(cfg().task)
        return self._dataset

    def get_channel_idxs(self, channels: List[str]) -> List[int]:
        names = list(cfg().task.input_variables.keys())
        return [names.index(c) for c in channels]

    def get_model(self) -> nn.Module:
        importpath = f"sres.model.{self.model_name}.network"
        model_package = importlib.import_module(importpath)
        return model_package.get_model(**self.model_config).to(self.device)
