"""Quantiser primitives and the QuantLayer wrapper -- host-side mirror of the reference's
`quant/quant_layer.py` (same names, arguments, attributes and state conventions).

Execution model.  These modules are the *calibration-time* graph: lazy quantiser initialisation,
running-stat range tracking and AdaRound reconstruction need per-module state and autograd, so the
module forward is the torch graph on the GPU with the range / scale-search / AdaRound arithmetic
delegated to the sm_100a kernels (ops.minmax_rows, ops.mse_scale_search, ops.act_range_update,
ops.adaround_*).  Sampling does not go through these forwards: QuantModel.forward dispatches to the
fused step engine (engine.py), where every op is one of the library's own kernels.
There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import logging
import os
from enum import Enum
from typing import List, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

logger = logging.getLogger(__name__)

# The three contractions of every reconstructed layer (forward, dgrad, wgrad) on the library's tensor-core kernels; False keeps
# torch's conv / linear kernels in the reconstruction loop (comparison runs).
TC_RECONSTRUCTION = os.environ.get("TFMQ_TC_RECON", "1") != "0"


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: tfmq_b200 runs on sm_100a GPUs only; got a {x.device} tensor (no CPU fallback)")


class StraightThrough(nn.Module):
    def forward(self, x):
        return x


REDUCTION = Enum("REDUCTION", ("NONE", "ALL"))
QMODE = Enum("QMODE", ("QDIFF", "NORMAL", "PTQD"))


def lp_loss(pred: torch.Tensor, tgt: torch.Tensor, p: float = 2.0, reduction: REDUCTION = REDUCTION.NONE):
    """reference quant/quant_layer.py:146-156."""
    err = (pred - tgt).abs().pow(p)
    if reduction == REDUCTION.NONE:
        return err.sum(1).mean()
    if reduction == REDUCTION.ALL:
        return err.mean()
    raise NotImplementedError


def ste_round(x: torch.Tensor) -> torch.Tensor:
    return (x.round() - x).detach() + x


def _minmax_from_range(x_min: float, x_max: float, like: torch.Tensor, symmetric: bool, level: int,
                       always_zero: bool):
    """(delta, zero_point) from a range, through the same double -> fp32 path as the reference
    (quant/quant_layer.py:24-35)."""
    x_min, x_max = min(x_min, 0), max(x_max, 0)
    delta = torch.tensor(float(x_max - x_min) / (level - 1))
    if symmetric:
        m = max(abs(x_min), x_max)
        x_min, x_max = -m, m
        delta = torch.tensor(float(x_max - x_min) / (level - 2))
    if always_zero:
        delta = torch.tensor(float(x_max) / (level - 1))
    if delta < 1e-8:
        delta = torch.tensor(1e-8)
    zero_point = torch.round(-x_min / delta) if not (symmetric or always_zero) else torch.tensor(0.0)
    return delta.to(like), zero_point.to(like)


def minmax(x: torch.Tensor, symmetric: bool = False, level: int = 256, always_zero: bool = False):
    """Scaler.MINMAX (quant/quant_layer.py:20-35); range reduction on the GPU."""
    _require_cuda(x, "minmax")
    from .. import ops
    mm = ops.minmax_rows(x.detach().reshape(1, -1).contiguous().float()).cpu()
    return _minmax_from_range(mm[0, 0].item(), mm[0, 1].item(), x, symmetric, level, always_zero)


def mse(x: torch.Tensor, symmetric: bool = False, level: int = 256, always_zero: bool = False):
    """Scaler.MSE (quant/quant_layer.py:38-64): 80-candidate L2.4 search, one kernel instead of
    80 x ~8 tensor ops and 80 host syncs."""
    _require_cuda(x, "mse")
    if symmetric or always_zero:
        raise NotImplementedError("Scaler.MSE: only the asymmetric form is used by the entry points")
    from .. import ops
    d, z = ops.mse_scale_search(x.detach().reshape(1, -1).contiguous().float(), level)
    return d[0].to(x), z[0].to(x)


def kl(x, symmetric=False, level=256, always_zero=False):
    raise NotImplementedError("Scaler.KL is not used by any entry point of the hot path")


def hist(x, symmetric=False, level=256, always_zero=False):
    raise NotImplementedError("Scaler.HIST is not used by any entry point of the hot path")


class Scaler(Enum):
    MINMAX = minmax
    MSE = mse
    KL = kl
    HIST = hist


class UniformAffineQuantizer(nn.Module):
    """reference quant/quant_layer.py:163-253 -- same constructor, attributes and lazy-init protocol."""

    def __init__(self, bits: int = 8, symmetric: bool = False, channel_wise: bool = False,
                 scaler: Scaler = Scaler.MINMAX, leaf_param: bool = False, always_zero: bool = False,
                 quant_emb: bool = False) -> None:
        super().__init__()
        self.level = 2 ** bits
        self.symmetric = symmetric
        self.channel_wise = channel_wise
        self.scaler = scaler
        self.leaf_param = leaf_param
        if self.leaf_param:
            self.x_min, self.x_max = None, None
        self.running_stat = False
        self.always_zero = always_zero
        self.delta = None
        self.zero_point = None
        self.init = False
        self.quant_emb = quant_emb
        self._range_state = None   # device (x_min, x_max, scratch) for the running-stat kernel

    # ------------------------------------------------------------------ init
    def _init_quantization_param(self, x: torch.Tensor, channel_wise: bool = False):
        _require_cuda(x, "UniformAffineQuantizer")
        from .. import ops
        if channel_wise:
            rows = x.detach().reshape(x.shape[0], -1).contiguous().float()
            if self.scaler is mse and not (self.symmetric or self.always_zero):
                delta, zero_point = ops.mse_scale_search(rows, self.level)
            else:
                mm = ops.minmax_rows(rows).cpu()
                ds, zs = [], []
                for c in range(rows.shape[0]):
                    d, z = _minmax_from_range(mm[c, 0].item(), mm[c, 1].item(), mm, self.symmetric, self.level,
                                              self.always_zero)
                    ds.append(d), zs.append(z)
                delta, zero_point = torch.stack(ds).to(x.device), torch.stack(zs).to(x.device)
            shape = (-1,) + (1,) * (x.dim() - 1)
            return delta.view(shape), zero_point.view(shape)
        if self.leaf_param:
            self.x_min, self.x_max = x.data.min(), x.data.max()
        return self.scaler(x, self.symmetric, self.level, self.always_zero)

    def bounds(self):
        if self.symmetric and not self.always_zero:
            return -self.level // 2, self.level // 2 - 1
        return 0, self.level - 1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _require_cuda(x, "UniformAffineQuantizer")
        if not self.init:
            self.delta, self.zero_point = self._init_quantization_param(x, self.channel_wise)
            if self.leaf_param:
                self.delta = nn.Parameter(self.delta)
            self.init = True
        if self.running_stat:
            self.act_momentum_update(x)
        nb, pb = self.bounds()
        x_q = torch.clamp(ste_round(x / self.delta) + self.zero_point, nb, pb)
        return self.delta * (x_q - self.zero_point)

    def act_momentum_update(self, x: torch.Tensor, act_range_momentum: float = 0.95) -> None:
        """reference :229-244, fused: batch min/max + EMA + MINMAX in two launches, no host sync."""
        assert self.init and self.leaf_param
        from .. import ops
        if self._range_state is None:
            st = torch.empty(4, device=x.device, dtype=torch.float32)
            st[0], st[1] = self.x_min, self.x_max
            bits = st.view(torch.int32)
            bits[2] = 0x7F800000            # ordered(+inf)
            bits[3] = -2139095041           # ordered(-inf) = 0xFF800000 ^ 0x7FFFFFFF as int32
            self._range_state = st
            self._aq = torch.empty(2, device=x.device, dtype=torch.float32)
        xc = x.detach().float().contiguous()
        ops.act_range_update(xc.reshape(-1, xc.shape[-1]), self._range_state, self._aq, act_range_momentum, self.level)
        self.x_min, self.x_max = self._range_state[0], self._range_state[1]
        self.zero_point = self._aq[1].clone()
        self.delta = nn.Parameter(self._aq[0].clone())

    def bitwidth_refactor(self, bits: int = 8) -> None:
        self.level = 2 ** bits

    def extra_repr(self) -> str:
        return (f"level={self.level}, symmetric={self.symmetric}, channel_wise={self.channel_wise}, "
                f"scaler={self.scaler.__name__}, leaf_param={self.leaf_param}")


class QuantLayer(nn.Module):
    """reference quant/quant_layer.py:259-354."""

    QMAP = {nn.Conv2d: F.conv2d, nn.Linear: F.linear}

    def __init__(self, layer: Union[nn.Conv2d, nn.Linear, nn.Conv1d], wq_params: dict = {}, aq_params: dict = {},
                 disable_aq: bool = False, aq_mode: List[int] = [QMODE.QDIFF.value], quant_emb: bool = False) -> None:
        super().__init__()
        self.wq_params, self.aq_params = wq_params, aq_params
        self.fwd_kwargs = {}
        if isinstance(layer, (nn.Conv2d, nn.Conv1d)):
            self.fwd_kwargs = dict(stride=layer.stride, padding=layer.padding, dilation=layer.dilation,
                                   groups=layer.groups)
        self.kwd_func = self.QMAP[type(layer)]
        self.w = layer.weight
        self.original_w = self.w.data.clone()
        self.b = self.original_b = None
        if layer.bias is not None:
            self.b = layer.bias
            self.original_b = self.b.data.clone()
        self.use_wq = self.use_aq = False
        self.disable_aq = disable_aq
        self.aq_mode = aq_mode
        self.quant_emb = quant_emb
        self.wq_params["quant_emb"] = quant_emb
        self.wqtizer = UniformAffineQuantizer(**self.wq_params)
        self.aqtizer = UniformAffineQuantizer(**self.aq_params)
        self.split = 0
        self.act_func = StraightThrough()
        self.ignore_recon = False
        self.extra_repr = layer.extra_repr
        self.w_override = None   # reconstruction injects the soft-rounded weight here (autograd leaf)

    def forward(self, x: torch.Tensor, split: int = 0) -> torch.Tensor:
        _require_cuda(x, "QuantLayer")
        if split != 0:
            # the Q-Diffusion shortcut split is unreachable from the entry points: shortcuts are never
            # QuantLayers (quant/quant_model.py:57-58)
            raise NotImplementedError("QuantLayer split path is dead in the reference entry points")
        if self.use_aq and not self.disable_aq:
            x = self.aqtizer(x)
        if self.use_wq:
            w = self.w_override if self.w_override is not None else self.wqtizer(self.w)
            b = self.b
        else:
            w, b = self.original_w, self.original_b
        w = w.to(x.device)
        if isinstance(b, torch.Tensor):
            b = b.to(x.device)
        if self.w_override is not None and self.use_wq and TC_RECONSTRUCTION:
            # reconstruction loop: forward / dgrad / wgrad of the layer on this library's tcgen05 kernels (tc_autograd.py)
            from .tc_autograd import tc_conv
            y = tc_conv(x, w, b if isinstance(b, torch.Tensor) else None, self.fwd_kwargs)
            if y is not None:
                return self.act_func(y)
        return self.act_func(self.kwd_func(x, w, b, **self.fwd_kwargs))

    def _apply(self, fn, *a, **kw):
        """`.to()` / `.cuda()` also move the FP copies.  The reference keeps `original_w` as a plain tensor attribute and
        re-copies it to the input's device on every forward (quant/quant_layer.py:333-335); a QuantModel built on the CPU and
        moved afterwards (cali_model_multi) would otherwise initialise AdaRound's alpha from a CPU tensor."""
        out = super()._apply(fn, *a, **kw)
        self.original_w = fn(self.original_w)
        if isinstance(self.original_b, torch.Tensor):
            self.original_b = fn(self.original_b)
        return out

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        self.use_wq = use_wq if not self.ignore_recon else False
        self.use_aq = use_aq if not self.ignore_recon else False

    def set_running_stat(self, running_stat: bool) -> None:
        self.aqtizer.running_stat = running_stat
