"""Deterministic synthetic model state shared by the golden generator (which feeds it to the
REFERENCE) and by the tests (which feed it to the product / the oracle).  Plain torch CPU ops only;
nothing here imports the reference, the oracle or the product."""
from __future__ import annotations

import math
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


@torch.no_grad()
def fill_state_dict(model: torch.nn.Module, seed: int = 1234) -> None:
    """Seeded random init keyed by parameter NAME (independent of construction order): conv / linear
    weights ~ N(0, 1/fan_in), biases ~ N(0, 0.05^2), norm scales 1 + N(0, 0.1^2).  The reference's
    zero-initialised `zero_module` convs get ordinary weights too, so every layer has a non-degenerate range."""
    sd = model.state_dict()
    for name in sorted(sd):
        p = sd[name]
        g = _gen(name, seed)
        if p.dim() >= 2:
            fan_in = p[0].numel()
            p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
        elif "norm" in name and name.endswith("weight") or name.endswith("in_layers.0.weight") \
                or name.endswith("out_layers.0.weight") or name.endswith("out.0.weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        else:
            p.copy_(0.05 * torch.randn(p.shape, generator=g))


def synth_alpha(name: str, w: torch.Tensor, delta: torch.Tensor, seed: int = 1234, noise: float = 0.5) -> torch.Tensor:
    """AdaRound alpha = its analytic initialisation (rounding == nearest) plus seeded noise, so a few
    percent of the hard rounding decisions differ from round-to-nearest, as after a real reconstruction."""
    rest = (w / delta) - torch.floor(w / delta)
    alpha = -torch.log(1.2 / (rest + 0.1) - 1)
    return alpha + noise * torch.randn(w.shape, generator=_gen(name + ".alpha", seed))


def latents(shape, seed: int) -> torch.Tensor:
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def ddim_betas(beta_start=0.0001, beta_end=0.02, n=1000) -> torch.Tensor:
    """linear schedule of ddim/runners/diffusion.py:49-52 (float64 linspace -> fp32 tensor)."""
    import numpy as np
    return torch.from_numpy(np.linspace(beta_start, beta_end, n, dtype=np.float64)).float()
