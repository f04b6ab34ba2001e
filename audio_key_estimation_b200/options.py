"""The ``opt`` namespace of the reference's argparse front-ends, at its defaults."""
from __future__ import annotations

import argparse


def default_opt(**overrides) -> argparse.Namespace:
    """``opt`` as ``train_model.py:160-242`` builds it when no flag is given (eval.py:139-223 agrees
    on every field the model reads)."""
    d = dict(batch_size=8, lr=3e-4, drop=0.0, reg=0, gamma=0.96, acc_grad=8, epochs=100, window_size=592,
             local=False, gpu=0, octaves=8, conv_layers=3, n_filters=4, num_layers=2, kernel_size=7,
             key_weight=1.0, tonic_weight=1.0, genre_weight=0.1, resblock=False, denseblock=False, frames=5,
             genre=False, stay_sixth=False, p2pc_conv=False, head_layers=2, loc_window_size=10, time_pool_size=2,
             only_semitones=False, multi_scale=False, no_test=False, debug=False, linear_reg_multi=False,
             use_cos=False, pc2p_mem=False, no_ckpt=False, max_pool=False)
    d.update(overrides)
    return argparse.Namespace(**d)
