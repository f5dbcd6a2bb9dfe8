"""Device selection, mirror of sres/base/gpu.py:6-21.  Under torchrun the local rank picks the GPU
(one process per GPU); otherwise `pipeline.gpu` / FMOD_GPU as in the reference.  There is no CPU
branch: the RCAN hot path needs the CUDA library."""
import os

import torch

from sres.base.util.config import cfg


def set_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("sres (B200 build): no CUDA device available and there is no CPU fallback")
    index = int(os.environ["LOCAL_RANK"]) if "LOCAL_RANK" in os.environ else int(cfg().pipeline.gpu)
    torch.cuda.set_device(index)
    return torch.device("cuda", index)


def get_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("sres (B200 build): no CUDA device available and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def save_memory_snapshot():
    return None
