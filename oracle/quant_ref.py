"""ORACLE (test infrastructure, never shipped or benchmarked as the product).

CPU restatement of TFMQ-DM's quantiser arithmetic in plain torch fp32 ops, written from the
reference's behaviour; every function cites the reference lines it follows (paths relative to the
reference root).  Parity pin: tests/golden/*.pt were produced by importing the reference itself
(tests/golden/make_golden.py) and tests/test_oracle_golden.py checks these functions against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this package.
"""
from __future__ import annotations

import math

import torch


# --------------------------------------------------------------------------- scalers
def minmax_scale(x: torch.Tensor, level: int = 256, symmetric: bool = False, always_zero: bool = False):
    """quant/quant_layer.py:20-35 (Scaler.MINMAX): range always includes 0; delta is computed in
    Python double and stored as fp32; zero_point = round(-x_min / delta) in fp32."""
    x_min = min(x.min().item(), 0)
    x_max = max(x.max().item(), 0)
    delta = torch.tensor(float(x_max - x_min) / (level - 1))
    if symmetric:
        m = max(abs(x_min), x_max)
        x_min, x_max = -m, m
        delta = torch.tensor(float(x_max - x_min) / (level - 2))
    if always_zero:
        delta = torch.tensor(float(x_max) / (level - 1))
    if delta < 1e-8:
        delta = torch.tensor(1e-8)
    if symmetric or always_zero:
        zp = torch.tensor(0.0)
    else:
        zp = torch.round(-x_min / delta)
    return delta.to(torch.float32), zp.to(torch.float32)


def lp_loss(pred: torch.Tensor, tgt: torch.Tensor, p: float = 2.0, reduce_all: bool = False) -> torch.Tensor:
    """quant/quant_layer.py:146-156: REDUCTION.NONE -> sum over dim 1 then mean; ALL -> mean."""
    e = (pred - tgt).abs().pow(p)
    return e.mean() if reduce_all else e.sum(1).mean()


def mse_candidates(x_min: float, x_max: float, level: int = 256):
    """The 80 (delta, zero_point) candidates of Scaler.MSE (quant/quant_layer.py:45-55, asymmetric)."""
    out = []
    for i in range(80):
        new_min = x_min * (1.0 - (i * 0.01))
        new_max = x_max * (1.0 - (i * 0.01))
        new_delta = torch.tensor(float(new_max - new_min) / (level - 1))
        new_zp = torch.round(-new_min / new_delta)
        out.append((new_delta, new_zp))
    return out


def mse_scale(x: torch.Tensor, level: int = 256, return_index: bool = False):
    """quant/quant_layer.py:38-64 (Scaler.MSE, asymmetric): score = mean |dq(x)-x|^2.4, first strict minimum."""
    x_min, x_max = x.min().item(), x.max().item()
    best, s, best_i = (None, None), 1e10, -1
    for i, (d, z) in enumerate(mse_candidates(x_min, x_max, level)):
        x_q = torch.clamp(torch.round(x / d) + z, 0, level - 1)
        x_dq = d * (x_q - z)
        new_s = lp_loss(x_dq, x, p=2.4, reduce_all=True)
        if new_s < s:
            s, best, best_i = new_s, (d, z), i
    if return_index:
        return best[0], best[1], best_i
    return best


def channel_wise(scale_fn, w: torch.Tensor, level: int):
    """quant/quant_layer.py:193-204: one (delta, zp) per leading-dim slice, shaped [C,1,...]."""
    c = w.shape[0]
    delta, zp = torch.empty(c), torch.empty(c)
    for i in range(c):
        d, z = scale_fn(w[i], level)
        delta[i], zp[i] = d, z
    shape = (-1,) + (1,) * (w.dim() - 1)
    return delta.view(shape), zp.view(shape)


# --------------------------------------------------------------------------- uniform affine quantiser
def uaq_codes(x: torch.Tensor, delta, zp, level: int = 256) -> torch.Tensor:
    """quant/quant_layer.py:223-225 (asymmetric): clamp(round(x/delta) + zp, 0, level-1)."""
    return torch.clamp(torch.round(x / delta) + zp, 0, level - 1)


def uaq_fake_quant(x: torch.Tensor, delta, zp, level: int = 256) -> torch.Tensor:
    """quant/quant_layer.py:223-227: delta * (codes - zp)."""
    return delta * (uaq_codes(x, delta, zp, level) - zp)


def act_momentum_update(x: torch.Tensor, x_min, x_max, momentum: float = 0.95, level: int = 256):
    """quant/quant_layer.py:229-244: EMA of the batch min/max, then MINMAX on the clipped tensor
    (whose extremes are exactly the EMA values)."""
    x_min = x_min * momentum + x.min() * (1.0 - momentum)
    x_max = x_max * momentum + x.max() * (1.0 - momentum)
    xc = torch.where(x < x_min, x_min, x.clone())
    xc = torch.where(xc > x_max, x_max, xc)
    xc[..., 0] = x_min
    xc[..., 1] = x_max
    delta, zp = minmax_scale(xc, level)
    return x_min, x_max, delta, zp


# --------------------------------------------------------------------------- AdaRound
GAMMA, ZETA = -0.1, 1.1


def adaround_init_alpha(w: torch.Tensor, delta) -> torch.Tensor:
    """quant/adaptive_rounding.py:31-36."""
    rest = (w / delta) - torch.floor(w / delta)
    return -torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1)


def adaround_h(alpha: torch.Tensor) -> torch.Tensor:
    """quant/adaptive_rounding.py:40-41 (rectified sigmoid)."""
    return torch.clamp(torch.sigmoid(alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)


def adaround_codes(w, delta, zp, alpha, level: int = 16, soft: bool = False) -> torch.Tensor:
    """quant/adaptive_rounding.py:51-68: floor(w/delta) + (h(alpha) | alpha>=0) + zp, clamped."""
    x_floor = torch.floor(w / delta)
    x_int = x_floor + (adaround_h(alpha) if soft else (alpha >= 0).float())
    return torch.clamp(x_int + zp, 0, level - 1)


def adaround_fake_quant(w, delta, zp, alpha, level: int = 16, soft: bool = False) -> torch.Tensor:
    """quant/adaptive_rounding.py:67-70."""
    return delta * (adaround_codes(w, delta, zp, alpha, level, soft) - zp)


def round_loss(alpha: torch.Tensor, b: float, w: float) -> torch.Tensor:
    """quant/reconstruction_util.py:69-70: w * sum(1 - |2h-1|^b)."""
    return w * (1 - ((adaround_h(alpha) - 0.5).abs() * 2).pow(b)).sum()


def temp_decay(t: int, t_max: int, rel_start_decay: float, start_b: float, end_b: float) -> float:
    """quant/reconstruction_util.py:176-198 (LinearTempDecay)."""
    start = rel_start_decay * t_max
    if t < start:
        return start_b
    rel_t = (t - start) / (t_max - start)
    return end_b + (start_b - end_b) * max(0.0, 1 - rel_t)


def recon_loss(pred, tgt, alphas, count: int, iters: int, w: float = 0.01, warmup: float = 0.2,
               b_range=(20, 2)):
    """quant/reconstruction_util.py:36-91 (LossFunc, RLOSS.MSE + RELAXATION). `count` is 1-based.
    pred/tgt may be tuples (LossFuncTimeEmbedding, :119-165: rec summed over the outputs)."""
    if isinstance(pred, (tuple, list)):
        rec = sum(lp_loss(p_, t_, 2.0) for p_, t_ in zip(pred, tgt))
    else:
        rec = lp_loss(pred, tgt, 2.0)
    b = temp_decay(count, iters, warmup, b_range[0], b_range[1])
    if count < iters * warmup:
        return rec, rec, 0.0, 0.0
    rl = sum(round_loss(a, b, w) for a in alphas)
    return rec + rl, rec, rl, b


# --------------------------------------------------------------------------- QuantLayer
def quant_layer_forward(x, w, bias, wq=None, aq=None, conv: dict | None = None):
    """quant/quant_layer.py:306-340.  wq = (delta, zp) or (delta, zp, alpha) [hard AdaRound] or None
    (fp weights); aq = (delta, zp) or None; conv = F.conv2d kwargs, None for nn.Linear."""
    import torch.nn.functional as F

    if aq is not None:
        x = uaq_fake_quant(x, aq[0], aq[1], 256)
    if wq is not None:
        if len(wq) == 3 and wq[2] is not None:
            w = adaround_fake_quant(w, wq[0], wq[1], wq[2], 16, soft=False)
        else:
            w = uaq_fake_quant(w, wq[0], wq[1], 16)
    if conv is None:
        return F.linear(x, w, bias)
    return F.conv2d(x, w, bias, **conv)


def float64_accumulation():
    """Context manager: evaluate every QuantLayer's conv / linear with float64 accumulation (rounded once to fp32)
    instead of fp32 -- a strictly more accurate evaluation of the SAME fake-quant network.  Used to measure how far fp
    re-association alone moves the reference path (the activation-code flip cascade): it calibrates the per-layer
    flip-rate gates of the GPU parity tests and made the `alt_*` entries of the golden fixtures."""
    import contextlib
    import sys
    import torch.nn.functional as F
    mod = sys.modules[__name__]

    @contextlib.contextmanager
    def cm():
        orig = mod.quant_layer_forward

        def qlf64(x, w, bias, wq=None, aq=None, conv=None):
            if aq is not None:
                x = uaq_fake_quant(x, aq[0], aq[1], 256)
            if wq is not None:
                w = adaround_fake_quant(w, wq[0], wq[1], wq[2], 16) if (len(wq) == 3 and wq[2] is not None) \
                    else uaq_fake_quant(w, wq[0], wq[1], 16)
            b = bias.double() if bias is not None else None
            if conv is None:
                return F.linear(x.double(), w.double(), b).float()
            return F.conv2d(x.double(), w.double(), b, **conv).float()
        mod.quant_layer_forward = qlf64
        try:
            yield
        finally:
            mod.quant_layer_forward = orig
    return cm()


def silu(x):
    """ddim/models/diffusion.py:27-29 `nonlinearity` (x*sigmoid(x)); torch.nn.SiLU in the LDM UNet."""
    return x * torch.sigmoid(x)


def ddim_coefficients(a_t: float, a_prev: float, eta: float = 0.0):
    """ddim/functions/denoising.py:31-36: returns sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), c2, c1."""
    c1 = eta * math.sqrt((1 - a_t / a_prev) * (1 - a_prev) / (1 - a_t))
    c2 = math.sqrt((1 - a_prev) - c1 ** 2)
    return math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_prev), c2, c1
