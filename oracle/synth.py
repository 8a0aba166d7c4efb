"""ORACLE -- test infrastructure only.

Deterministic synthetic weights and inputs that do not depend on torch's RNG stream, so the very same
``state_dict`` can be regenerated here (next to the live reference) and on the GPU box.  BatchNorm
statistics are randomised ("BN stress"): the default init makes eval-mode BN an identity and would
hide folding bugs.
"""
import zlib

import numpy as np
import torch


def _rng(seed, key):
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def synth_state_dict(template, seed=0, seg_logit_gain=1.0):
    """template: mapping key -> tensor (shapes/dtypes are used, values ignored)."""
    out = {}
    for k, v in template.items():
        shape = tuple(v.shape)
        g = _rng(seed, k)
        if k.endswith("num_batches_tracked"):
            t = np.zeros(shape, dtype=np.int64)
        elif k.endswith("running_var"):
            t = g.uniform(0.5, 1.5, shape)
        elif k.endswith("running_mean"):
            t = g.normal(0.0, 0.2, shape)
        elif k.endswith(".weight") and (k.rsplit(".", 1)[0] + ".running_mean") in template:
            t = g.uniform(0.6, 1.4, shape)  # BatchNorm gamma
        elif "_w1" in k or "_w2" in k:
            t = g.uniform(-0.2, 1.5, shape)  # some fusion weights negative -> exercised ReLU
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = g.normal(0.0, 1.0 / np.sqrt(fan_in), shape)
            if k == "segheader.decoder.8.conv.weight" or (k.startswith("segheader.decoder.") and k.endswith(".conv.weight") and shape[0] <= 16):
                t = t * seg_logit_gain
        else:  # biases (conv / BN)
            t = g.normal(0.0, 0.1, shape)
        out[k] = torch.from_numpy(np.asarray(t)).to(v.dtype)
    return out


def synth_input(B, H, W, seed=0):
    g = np.random.default_rng([seed, 12345])
    return torch.from_numpy(g.normal(0.0, 1.0, (B, 3, H, W)).astype(np.float32))
