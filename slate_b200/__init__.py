"""slate_b200 -- B200-native (sm_100a) trailing-matrix-update engine behind SLATE's
device-BLAS / tile-kernel boundary.  See DESIGN.md and include/slate_b200.h."""
from ._lib import lib, check, SB200Error, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
