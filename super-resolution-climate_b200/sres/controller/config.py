"""Enums, mirror of sres/controller/config.py."""
from enum import Enum


class ResultStructure(Enum):
    Tiles = "tiles"
    Image = "image"


class TSet(Enum):
    Train = "train"
    Validation = "valid"
    Test = "test"
    Upsample = "upsample"
