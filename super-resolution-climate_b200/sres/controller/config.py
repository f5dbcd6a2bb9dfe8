"""The two enumerations of the reference's controller API (sres/controller/config.py:6-16): which data split a call works on
and in which shape inference results come back.  Member names and values are the reference's -- they appear in its YAML
files (`task.ttsplit` keys), in checkpoint file names and in user scripts -- so they cannot differ; the lookup helpers are ours.
"""
from enum import Enum


class _ByValue(Enum):
    """Enum whose members can be looked up by their configuration string (`TSet.of("valid")`), case-insensitively, with an
    error that lists the legal spellings instead of the bare ValueError of Enum()."""

    @classmethod
    def of(cls, text):
        if isinstance(text, cls):
            return text
        key = str(text).strip().lower()
        for member in cls:
            if member.value == key or member.name.lower() == key:
                return member
        raise ValueError(f"{cls.__name__}: unknown value '{text}' (expected one of {[m.value for m in cls]})")


class ResultStructure(_ByValue):
    """Shape of `WorkflowController.inference` results (workflow.py:62-70): per-tile arrays or stitched images."""
    Tiles = "tiles"
    Image = "image"


class TSet(_ByValue):
    """Data split (dual_trainer.py:97-105, checkpoints.py:54-58): the value is the key of `task.ttsplit` and the checkpoint
    file suffix."""
    Train = "train"
    Validation = "valid"
    Test = "test"
    Upsample = "upsample"
