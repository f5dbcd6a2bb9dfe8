"""sres_b200 -- B200-native RCAN hot path for super-resolution-climate (Python binding of libsres_b200.so)."""
from ._lib import SresError, lib  # noqa: F401
