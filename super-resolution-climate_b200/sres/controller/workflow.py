"""WorkflowController, mirror of sres/controller/workflow.py:17-86 (train / initialize / init_context /
inference); plotting views and NetCDF result files are out of scope."""
import time
from typing import Any, Dict, List

from sres.base.util.config import ConfigContext, cfg
from sres.controller.config import ResultStructure, TSet
from sres.controller.dual_trainer import ModelTrainer


class WorkflowController(object):

    def __init__(self, cname: str, configuration: Dict[str, Any], **kwargs):
        self.cname = cname
        self.seed = kwargs.get("seed", int(time.time() / 60))
        self.interp_loss = kwargs.get("interp_loss", False)
        self.refresh = kwargs.get("refresh", False)
        self.config: ConfigContext = None
        self.trainer: ModelTrainer = None
        self.model = None
        ConfigContext.set_defaults(**configuration)

    def train(self, models: List[str], **kwargs):
        for model in models:
            with ConfigContext(self.cname, model=model, **kwargs) as cc:
                self.config = cc
                self.trainer = ModelTrainer(cc)
                self.trainer.train(cfg().task.nepochs, self.refresh, seed=self.seed, interp_loss=self.interp_loss)

    def initialize(self, cname, model, **kwargs):
        self.model = model
        self.config = ConfigContext.activate_global(cname, model=model, **kwargs)
        self.trainer = ModelTrainer(self.config)

    def init_context(self, cc: ConfigContext, model: str):
        self.model = model
        self.config = cc
        self.trainer = ModelTrainer(self.config)

    def inference(self, timestep: int, data_structure: ResultStructure, **kwargs):
        varnames = self.trainer.target_variables
        if data_structure == ResultStructure.Image:
            return self.trainer.process_image(TSet.Validation, timestep, interp_loss=True, **kwargs)
        if data_structure == ResultStructure.Tiles:
            res, losses = self.trainer.evaluate(TSet.Validation, time_index=timestep, **kwargs)
            return {v: res for v in varnames}, {v: losses for v in varnames}
        raise Exception(f"Unknown result structure: {data_structure}")
