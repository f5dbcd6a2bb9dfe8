"""Checkpoint files of the trainer.

Same file scheme and dictionary layout as the reference's CheckpointManager (sres/controller/checkpoints.py:11-67):
`{platform.results}/checkpoints/{task.training_version}.{train|valid}.pt` holding
{epoch, itime, model_state_dict, optimizer_state_dict, loss}.  model_state_dict carries fp32 tensors under the
reference's parameter names and optimizer_state_dict has torch.optim.Adam's layout (sres_b200.nn.FusedAdam reads
and writes it), so checkpoints move between this build and the reference in both directions.

Unlike the reference, a checkpoint that exists but cannot be applied is an error, not a silent restart from
scratch: the optimizer state is validated BEFORE the model weights are touched.
"""
import os
import shutil
import time
from typing import Any, Dict, Optional

import torch

from sres.base.util.config import cfg
from sres.controller.config import TSet


class CheckpointError(RuntimeError):
    pass


class CheckpointManager:

    def __init__(self, model, optimizer):
        self.model, self.optimizer = model, optimizer

    @staticmethod
    def checkpoint_path(tset: TSet, backup: bool = False) -> str:
        which = TSet.Validation if tset == TSet.Test else tset     # test shares the validation file
        folder = os.path.join(str(cfg().platform.results), "checkpoints")
        os.makedirs(folder, mode=0o777, exist_ok=True)
        stem = f"{cfg().task.training_version}.{which.value}" + (".backup" if backup else "")
        return os.path.join(folder, stem + ".pt")

    def save_checkpoint(self, epoch: int, itime: int, tset: TSet, loss: float, interp_loss: float = float("nan")) -> str:
        started = time.time()
        path = self.checkpoint_path(tset)
        if os.path.isfile(path):
            shutil.copyfile(path, self.checkpoint_path(tset, backup=True))
        payload = {"epoch": epoch, "itime": itime, "model_state_dict": self.model.state_dict(),
                   "optimizer_state_dict": self.optimizer.state_dict(), "loss": loss}
        torch.save(payload, path)
        print(f" *** SAVE {tset.name} checkpoint, loss={loss:.5f} ({interp_loss:.5f}), to {path}, dt={time.time() - started:.4f} sec")
        return path

    def load_checkpoint(self, tset: TSet = TSet.Train, update_model: bool = False, quiet: bool = False, **_ignored) -> Optional[Dict[str, Any]]:
        """The training state {epoch, itime, loss} of the checkpoint, {} when there is no file yet.  With update_model the
        optimizer state and then the weights are restored; a file that does not fit raises CheckpointError."""
        path = self.checkpoint_path(tset)
        if not os.path.exists(path):
            if not quiet:
                print(f"No checkpoint file found at '{path}': starting from scratch.")
            return {}
        state = torch.load(path, map_location="cpu", weights_only=False)
        if update_model:
            try:
                self.optimizer.load_state_dict(state.pop("optimizer_state_dict"))   # validates before anything changes
                self.model.load_state_dict(state.pop("model_state_dict"))
            except Exception as err:
                raise CheckpointError(f"checkpoint {path} does not fit this model / optimizer: {err}") from err
        return state

    def clear_checkpoints(self) -> None:
        for tset in (TSet.Train, TSet.Validation):
            path = self.checkpoint_path(tset)
            if os.path.exists(path):
                os.remove(path)
