"""neptune_b200 -- B200-native (sm_100a) implementation of NEPTUNE's per-agent replan hot path.

Host-side modules (params, batch, scenes) are importable anywhere; every compute entry point goes
through the CUDA library behind the C-ABI of include/neptune_b200.h (neptune_b200.capi) and raises
if that library or a GPU is missing -- there is no CPU fallback.
"""
from .params import Params, config  # noqa: F401
from .batch import ReplanBatch, ReplanResult  # noqa: F401
