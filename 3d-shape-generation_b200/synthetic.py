"""Synthetic, calibrated weights for benchmarks and demos (no checkpoints ship with the
reference, and a plain random-init denoiser makes the reverse loop diverge -- SURVEY H1)."""
from __future__ import annotations

import math
from typing import Dict

import torch


def synthetic_state_dict(model, seed: int = 24, alpha: float = 1.0 / 33.0, bn_seed: int = 7) -> Dict[str, torch.Tensor]:
    """Weights for `model` (a PointCloudDiffusion) drawn with the reference's initialisation law
    (Kaiming-normal fan_out, reference diffusion.py:40-54), `output.3` scaled by `alpha` so the
    loop stays finite, and randomised BatchNorm statistics / affine / biases."""
    g = torch.Generator().manual_seed(seed)
    gb = torch.Generator().manual_seed(bn_seed)
    sd = {}
    for key, ref in model.state_dict().items():
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(1, dtype=torch.int64)
        elif key.endswith("running_mean"):
            sd[key] = 0.1 * torch.randn(shape, generator=gb)
        elif key.endswith("running_var"):
            sd[key] = 0.5 + torch.rand(shape, generator=gb)
        elif _is_bn(model, key):
            sd[key] = (1.0 + 0.2 * torch.randn(shape, generator=gb)) if key.endswith("weight") else 0.1 * torch.randn(shape, generator=gb)
        elif ref.dim() >= 2:
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[0])
        else:
            sd[key] = torch.randn(shape, generator=gb) * 0.05
    sd["model.output.3.weight"] = sd["model.output.3.weight"] * alpha
    sd["model.output.3.bias"] = sd["model.output.3.bias"] * alpha
    return sd


def _is_bn(model, key: str) -> bool:
    mod = model
    for part in key.split(".")[:-1]:
        mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
    return isinstance(mod, torch.nn.BatchNorm1d)
