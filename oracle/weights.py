"""ORACLE (test infrastructure): deterministic synthetic weights keyed by state-dict name.

There are no StoryGen / SD-1.5 weights in this environment, so every parity test and
the benchmark use seeded random weights.  They are generated per KEY (not per construction order), so
the reference model, the oracle and the CUDA model — which build their modules in different orders —
receive bit-identical tensors for the same seed."""
from __future__ import annotations

import zlib
from typing import Dict

import torch


def seeded_state_dict(shapes: Dict[str, torch.Size], seed: int = 0, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator("cpu").manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if len(shape) >= 2:  # conv / linear weight: ~unit-gain variance preserving init
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            w = r * (gain / fan_in ** 0.5)
        elif key.endswith("weight"):  # norm scale
            w = 1.0 + 0.1 * r
        else:  # bias / norm shift
            w = 0.05 * r
        out[key] = w
    return out


def load_seeded(module: torch.nn.Module, seed: int = 0, gain: float = 1.0):
    sd = module.state_dict()
    module.load_state_dict(seeded_state_dict({k: v.shape for k, v in sd.items()}, seed, gain))
    return module
