"""b200fdtd — host side of the B200-native FDTD engine (ctypes over libb200fdtd.so)."""
from ._lib import B200FDTDError, lib, SO_PATH  # noqa: F401
