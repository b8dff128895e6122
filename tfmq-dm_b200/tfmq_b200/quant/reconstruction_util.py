"""Loss pieces of the AdaRound reconstruction -- mirror of the reference's
`quant/reconstruction_util.py` (RLOSS, LossFunc, LossFuncTimeEmbedding, LinearTempDecay).
The fused optimiser step (reconstruction.py) evaluates the same quantities inside the library's
kernels; these classes remain for API compatibility and for tests."""
from __future__ import annotations

import logging
from enum import Enum
from typing import Iterable, Union

import torch

from .adaptive_rounding import AdaRoundQuantizer
from .quant_block import BaseQuantBlock, QuantTemporalInformationBlock, QuantTemporalInformationBlockDDIM
from .quant_layer import QuantLayer, lp_loss

logger = logging.getLogger(__name__)

RLOSS = Enum("RLOSS", ("RELAXATION", "MSE", "FISHER_DIAG", "FISHER_FULL", "NONE"))
print_freq = 2000


class LinearTempDecay:
    """b = start_b until rel_start_decay * t_max, then linear to end_b (reference :176-198)."""

    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: int = 10, end_b: int = 2) -> None:
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b, self.end_b = start_b, end_b

    def __call__(self, t) -> float:
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, 1 - rel_t)


def unit_layers(o: Union[QuantLayer, BaseQuantBlock]) -> Iterable[QuantLayer]:
    """The QuantLayers whose rounding a reconstruction unit optimises (reference :66-80,141-161)."""
    if isinstance(o, QuantLayer):
        return [o]
    if isinstance(o, QuantTemporalInformationBlock):
        seen = [m for m in o.modules() if isinstance(m, QuantLayer)]
        for seq in o.emb_layers:
            seen += [m for m in seq.modules() if isinstance(m, QuantLayer)]
        return [m for m in seen if not m.ignore_recon]
    if isinstance(o, QuantTemporalInformationBlockDDIM):
        seen = [m for m in o.modules() if isinstance(m, QuantLayer)] + list(o.temb_projs)
        return [m for m in seen if not m.ignore_recon]
    return [m for m in o.modules() if isinstance(m, QuantLayer) and m.quant_emb is False and not m.ignore_recon]


def _round_term(layers, w: float, b: float):
    total = 0
    for m in layers:
        assert isinstance(m.wqtizer, AdaRoundQuantizer)
        h = m.wqtizer.get_soft_tgt()
        total = total + w * (1 - ((h - 0.5).abs() * 2).pow(b)).sum()
    return total


class LossFunc:
    def __init__(self, o, round_loss: RLOSS = RLOSS.RELAXATION, w: float = 1.0, rec_loss: RLOSS = RLOSS.MSE,
                 max_count: int = 2000, b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0,
                 p: float = 2.0) -> None:
        self.o, self.round_loss, self.w, self.rec_loss, self.p = o, round_loss, w, rec_loss, p
        self.loss_start = max_count * warmup
        self.temp_decay = LinearTempDecay(max_count, warmup + (1 - warmup) * decay_start, b_range[0], b_range[1])
        self.count = 0

    def __call__(self, pred, tgt, grad=None) -> torch.Tensor:
        self.count += 1
        if self.rec_loss != RLOSS.MSE:
            raise ValueError(f"Not supported reconstruction loss function: {self.rec_loss}")
        if isinstance(pred, (tuple, list)):
            rec = sum(lp_loss(p_, t_, p=self.p) for p_, t_ in zip(pred, tgt))
        else:
            rec = lp_loss(pred, tgt, p=self.p)
        b = self.temp_decay(self.count)
        if self.count < self.loss_start or self.round_loss == RLOSS.NONE:
            b = rl = 0
        else:
            rl = _round_term(unit_layers(self.o), self.w, b)
        total = rec + rl
        if self.count % print_freq == 0:
            logger.info("Total loss:\t{:.8f} (rec:{:.8f}, round:{:.8f})\tb={:.2f}\tcount={}".format(
                float(total), float(rec), float(rl), b, self.count))
        return total


class LossFuncTimeEmbedding(LossFunc):
    """Same loss over the tuple of TIB outputs (reference :94-173)."""
