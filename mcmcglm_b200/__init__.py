"""mcmcglm_b200 -- B200-native CGGibbs engine behind the mcmcglm() API of mathiaslj/mcmcglm.

Only the slice-within-Gibbs coordinate-update path runs here (CUDA, sm_100a, via libcggibbs.so);
see DESIGN.md.  Importing the package does not load CUDA; constructing an Engine does.
"""
from ._lib import CggError  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["Engine", "CggError"]
