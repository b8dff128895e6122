"""AdaRound weight quantiser: the drop-in for the reference's `quant/adaptive_rounding.py:12-74` (same class / enum names,
constructor, attributes `alpha soft_tgt gamma zeta level symmetric delta zero_point rmode`, and the same arithmetic bit for bit).

This module is the *state* and the eager definition of the rounding rule.  The hot uses do not go through `forward`:
  * sampling: `engine._QL` hands (w, delta, zero_point, alpha) to `tfmq_pack_w4`, which takes the hard decision
    floor(w / delta) + (alpha >= 0) once and stores packed int4 codes;
  * reconstruction: `tfmq_adaround_soft` materialises the soft weights and `tfmq_adaround_step` applies the chain rule to
    alpha together with the regulariser gradient and the Adam update (quant/reconstruction.py here).
"""
from __future__ import annotations

from enum import Enum

import torch
from torch import nn

from .quant_layer import UniformAffineQuantizer, ste_round

RMODE = Enum("RMODE", ("LEARNED_ROUND_SIGMOID", "NEAREST", "NEAREST_STE", "STOCHASTIC", "LEARNED_HARD_SIGMOID"))

# rectified-sigmoid stretch of AdaRound: h(alpha) = clamp(sigmoid(alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)
GAMMA, ZETA = -0.1, 1.1


def _follow(value, device):
    """Quantiser parameters trail the weights across devices (the reference re-assigns them on every call)."""
    return value.to(device) if torch.is_tensor(value) else value


class AdaRoundQuantizer(nn.Module):
    """Takes over the grid (delta, zero_point, level) of an initialised UniformAffineQuantizer and learns, per weight,
    whether to round down or up: w_q = delta * (clamp(floor(w / delta) + h + z, lo, hi) - z)."""

    def __init__(self, uaqtizer: UniformAffineQuantizer, w: torch.Tensor,
                 rmode: RMODE = RMODE.LEARNED_ROUND_SIGMOID) -> None:
        super().__init__()
        for inherited in ("level", "symmetric", "delta", "zero_point"):
            setattr(self, inherited, getattr(uaqtizer, inherited))
        self.rmode, self.soft_tgt = rmode, False
        self.gamma, self.zeta = GAMMA, ZETA
        self.alpha = None
        self.init_alpha(x=w.clone())

    # ------------------------------------------------------------------ alpha
    def init_alpha(self, x: torch.Tensor) -> None:
        """The alpha whose soft target equals the fractional part of w / delta, i.e. soft rounding starts at the FP weight
        and hard rounding at round-to-nearest: alpha = -log((zeta - gamma) / (frac - gamma) - 1)."""
        if self.rmode is not RMODE.LEARNED_HARD_SIGMOID:
            raise NotImplementedError(f"AdaRound initialisation exists for LEARNED_HARD_SIGMOID only, not {self.rmode}")
        self.delta = _follow(self.delta, x.device)
        steps = x / self.delta
        frac = steps - torch.floor(steps)
        self.alpha = nn.Parameter(-torch.log((self.zeta - self.gamma) / (frac - self.gamma) - 1))

    def get_soft_tgt(self) -> torch.Tensor:
        stretched = torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma
        return torch.clamp(stretched, 0, 1)

    # ------------------------------------------------------------------ rounding rules
    def _integer_steps(self, steps: torch.Tensor) -> torch.Tensor:
        mode = self.rmode
        if mode is RMODE.NEAREST:
            return torch.round(steps)
        if mode is RMODE.NEAREST_STE:
            return ste_round(steps)
        down = torch.floor(steps)
        if mode is RMODE.STOCHASTIC:
            return down + torch.bernoulli(steps - down)
        if mode is RMODE.LEARNED_HARD_SIGMOID:
            if self.soft_tgt:
                return down + self.get_soft_tgt().to(steps.device)
            alpha = self.alpha if self.alpha.device == steps.device else self.alpha.to(steps.device)
            return down + (alpha >= 0).float()
        raise NotImplementedError(f"rounding mode {mode}")

    def code_range(self):
        half = self.level // 2
        return (-half, half - 1) if self.symmetric else (0, self.level - 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.delta, self.zero_point = _follow(self.delta, x.device), _follow(self.zero_point, x.device)
        lo, hi = self.code_range()
        codes = torch.clamp(self._integer_steps(x / self.delta) + self.zero_point, lo, hi)
        return self.delta * (codes - self.zero_point)

    def extra_repr(self) -> str:
        return f"level={self.level}, symmetric={self.symmetric}, rmode={self.rmode}"
