"""Deterministic stand-in weights for machines without the pretrained `ace_encoder_pretrained.pt` (22 MB, not
redistributed here): same names / shapes, numpy RandomState values at He scale. For tests and benchmarks only."""
import math

import numpy as np
import torch

from .encoder import ENCODER_KEYS, ENCODER_SHAPES


def random_encoder_state(seed):
    rs = np.random.RandomState(seed)
    sd = {}
    for k in ENCODER_KEYS:
        cout, cin, kh, kw = ENCODER_SHAPES[k]
        std = math.sqrt(2.0 / (cin * kh * kw))
        sd[k + ".weight"] = torch.from_numpy((rs.standard_normal((cout, cin, kh, kw)) * std).astype(np.float32))
        sd[k + ".bias"] = torch.from_numpy(rs.uniform(-0.1, 0.1, (cout,)).astype(np.float32))
    return sd
