"""audio_key_estimation_b200 -- B200-native hot path of flo-stilz/Audio-Key-Estimation.

Constant-Q front-end (KeyDataset.py:485-509) + PitchClassNet forward (models.py:651-817) as
hand-written sm_100a CUDA kernels behind a C ABI (include/ake_b200.h, libake_b200.so), with this
thin Python host layer mirroring the reference's PyTorch-facing surface.
"""
from .models import PitchClassNet, PitchClassNet_Multi, decode, mirex_counters, mirex_from_counters  # noqa: F401
from .cqt import CQTPlan, cqt, cqt_logmag  # noqa: F401
from .pipeline import KeyEstimator  # noqa: F401
from .cache import cache_name, write_cqt_cache  # noqa: F401
from .options import default_opt  # noqa: F401
from .training import FusedAdam, TrainStep, criterion  # noqa: F401

__all__ = ["PitchClassNet", "PitchClassNet_Multi", "decode", "mirex_counters", "mirex_from_counters", "CQTPlan", "cqt", "cqt_logmag", "KeyEstimator", "cache_name", "write_cqt_cache", "default_opt", "TrainStep", "FusedAdam", "criterion"]
