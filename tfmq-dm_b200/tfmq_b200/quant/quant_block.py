"""Quantised blocks and the Temporal Information Block -- mirror of the reference's
`quant/quant_block.py` (class names, constructor arguments, attributes, `b2qb` map).

The blocks adopt the sub-modules of the FP block they replace (whose leaves QuantModel has already
turned into QuantLayers) and replay the FP block's dataflow.  They are matched by *class name*, so
they wrap both this package's host UNets and the reference's own model classes.
"""
from __future__ import annotations

import math
from types import MethodType
from typing import Dict, Tuple

import torch as th
import torch.nn as nn

from ..host.ddim_unet import get_timestep_embedding, nonlinearity
from ..host.ldm_unet import TimestepBlock, timestep_embedding
from .quant_layer import QuantLayer, StraightThrough, UniformAffineQuantizer


class BaseQuantBlock(nn.Module):
    def __init__(self, aq_params: dict = {}) -> None:
        super().__init__()
        self.use_wq = False
        self.use_aq = False        # never switched on by any caller: the attention-core quantisers are inert
        self.act_func = StraightThrough()
        self.ignore_recon = False

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        for m in self.modules():
            if isinstance(m, QuantLayer):
                m.set_quant_state(use_wq=use_wq, use_aq=use_aq)


def _softmax_quantizer(aq_params: dict, softmax_a_bit: int) -> UniformAffineQuantizer:
    p = aq_params.copy()
    p.update(bits=softmax_a_bit, symmetric=False, always_zero=True)
    return UniformAffineQuantizer(**p)


# ----------------------------------------------------------------------------- TIB
class QuantTemporalInformationBlockDDIM(BaseQuantBlock):
    """sinusoid(t) -> dense0 -> swish -> dense1 -> {swish -> temb_proj_i}: every time-embedding
    projection of the UNet as one reconstruction unit (reference :36-75).  Shares the QuantLayer
    objects that live inside the ResnetBlocks."""

    def __init__(self, temb: nn.Module, aq_params: dict = {}, ch: int = None) -> None:
        super().__init__(aq_params)
        self.temb = temb
        self.temb_projs = []
        self.ch = ch

    def add_temb_proj(self, temb_proj: nn.Linear) -> None:
        self.temb_projs.append(temb_proj)

    def forward(self, x: th.Tensor, t: th.Tensor) -> Tuple[th.Tensor]:
        assert t is not None
        temb = self.temb.dense[1](nonlinearity(self.temb.dense[0](get_timestep_embedding(t, self.ch))))
        return tuple(proj(nonlinearity(temb)) for proj in self.temb_projs)

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        super().set_quant_state(use_wq, use_aq)
        for proj in self.temb_projs:
            assert isinstance(proj, QuantLayer)
            proj.set_quant_state(use_wq=use_wq, use_aq=use_aq)


class QuantTemporalInformationBlock(BaseQuantBlock):
    """LDM / SD flavour: time_embed MLP + every ResBlock's emb_layers (reference :78-126)."""

    def __init__(self, t_emb: nn.Sequential, aq_params: dict = {}, model_channels: int = None,
                 num_classes: int = None) -> None:
        super().__init__(aq_params)
        self.t_emb = t_emb
        self.emb_layers = []
        self.label_emb_layer = None
        self.model_channels = model_channels
        self.num_classes = num_classes

    def add_emb_layer(self, layer: nn.Sequential) -> None:
        self.emb_layers.append(layer)

    def add_label_emb_layer(self, layer: nn.Sequential) -> None:
        self.label_emb = layer

    def forward(self, x: th.Tensor, t: th.Tensor, y: th.Tensor = None) -> Tuple[th.Tensor]:
        assert t is not None
        emb = self.t_emb(timestep_embedding(t, self.model_channels, repeat_only=False))
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
            emb = emb + self.label_emb(y)
        return tuple(layer(emb) for layer in self.emb_layers)

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        super().set_quant_state(use_wq, use_aq)
        for seq in self.emb_layers:
            for m in seq.modules():
                if isinstance(m, QuantLayer):
                    m.set_quant_state(use_wq=use_wq, use_aq=use_aq)


# ----------------------------------------------------------------------------- LDM / SD blocks
class QuantResBlock(BaseQuantBlock, TimestepBlock):
    """GN32 -> SiLU -> conv ; + SiLU -> Linear(emb) ; GN32 -> SiLU -> conv ; skip(x) + h (reference :130-209)."""

    def __init__(self, res, aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        for name in ("channels", "emb_channels", "dropout", "out_channels", "use_conv", "use_checkpoint",
                     "use_scale_shift_norm", "in_layers", "updown", "h_upd", "x_upd", "emb_layers", "out_layers",
                     "skip_connection"):
            setattr(self, name, getattr(res, name))

    def forward(self, x: th.Tensor, emb: th.Tensor = None, split: int = 0) -> th.Tensor:
        return self._forward(x, emb, split)

    def _forward(self, x: th.Tensor, emb: th.Tensor, split: int = 0) -> th.Tensor:
        if emb is None:
            assert len(x) == 2
            x, emb = x
        assert x.shape[2] == x.shape[3]
        if self.updown or self.use_scale_shift_norm:
            raise NotImplementedError("resblock up/down and scale-shift norm are unused by the four configs")
        h = self.in_layers(x)
        emb_out = self.emb_layers(emb).type(h.dtype)
        while emb_out.dim() < h.dim():
            emb_out = emb_out[..., None]
        h = self.out_layers(h + emb_out)
        return self.skip_connection(x) + h


def cross_attn_forward(self, x: th.Tensor, context: th.Tensor = None, mask: th.Tensor = None) -> th.Tensor:
    """CrossAttention.forward replacement (reference :212-245): quantised projections, fp attention core."""
    h = self.heads
    q = self.to_q(x)
    context = x if context is None else context
    k, v = self.to_k(context), self.to_v(context)
    b, n, _ = q.shape

    def heads_first(t):
        return t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], -1)

    q, k, v = heads_first(q), heads_first(k), heads_first(v)
    if self.use_aq:
        q, k = self.aqtizer_q(q), self.aqtizer_k(k)
    sim = th.einsum("b i d, b j d -> b i j", q, k) * self.scale
    if mask is not None:
        mask = mask.reshape(mask.shape[0], -1)
        sim.masked_fill_(~mask[:, None, :].repeat_interleave(h, 0), -th.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1)
    if self.use_aq:
        attn, v = self.aqtizer_w(attn), self.aqtizer_v(v)
    out = th.einsum("b i j, b j d -> b i d", attn, v)
    out = out.reshape(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1)
    return self.to_out(out)


class QuantBasicTransformerBlock(BaseQuantBlock):
    def __init__(self, tran, aq_params: dict = {}, softmax_a_bit: int = 8) -> None:
        super().__init__(aq_params)
        self.attn1, self.ff, self.attn2 = tran.attn1, tran.ff, tran.attn2
        self.norm1, self.norm2, self.norm3 = tran.norm1, tran.norm2, tran.norm3
        self.checkpoint = False
        for attn in (self.attn1, self.attn2):
            attn.aqtizer_q = UniformAffineQuantizer(**aq_params)
            attn.aqtizer_k = UniformAffineQuantizer(**aq_params)
            attn.aqtizer_v = UniformAffineQuantizer(**aq_params)
            attn.aqtizer_w = _softmax_quantizer(aq_params, softmax_a_bit)
            attn.forward = MethodType(cross_attn_forward, attn)
            attn.use_aq = False

    def forward(self, x: th.Tensor, context: th.Tensor = None) -> th.Tensor:
        assert context is not None
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x

    _forward = forward


class QuantQKMatMul(BaseQuantBlock):
    def __init__(self, aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        self.scale = None
        self.use_aq = False
        self.aqtizer_q = UniformAffineQuantizer(**aq_params)
        self.aqtizer_k = UniformAffineQuantizer(**aq_params)

    def forward(self, q: th.Tensor, k: th.Tensor) -> th.Tensor:
        q, k = q * self.scale, k * self.scale
        if self.use_aq:
            q, k = self.aqtizer_q(q), self.aqtizer_k(k)
        return th.einsum("bct,bcs->bts", q, k)


class QuantSMVMatMul(BaseQuantBlock):
    def __init__(self, aq_params: dict = {}, softmax_a_bit: int = 8) -> None:
        super().__init__(aq_params)
        self.use_aq = False
        self.aqtizer_v = UniformAffineQuantizer(**aq_params)
        self.aqtizer_w = _softmax_quantizer(aq_params, softmax_a_bit)

    def forward(self, weight: th.Tensor, v: th.Tensor) -> th.Tensor:
        if self.use_aq:
            weight, v = self.aqtizer_w(weight), self.aqtizer_v(v)
        return th.einsum("bts,bcs->bct", weight, v)


class QuantAttentionBlock(BaseQuantBlock):
    def __init__(self, attn, aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        self.channels, self.num_heads, self.use_checkpoint = attn.channels, attn.num_heads, attn.use_checkpoint
        self.norm, self.qkv, self.attention, self.proj_out = attn.norm, attn.qkv, attn.attention, attn.proj_out

    def forward(self, x: th.Tensor) -> th.Tensor:
        b, c, *spatial = x.shape
        x = x.reshape(b, c, -1)
        h = self.proj_out(self.attention(self.qkv(self.norm(x))))
        return (x + h).reshape(b, c, *spatial)

    _forward = forward


# ----------------------------------------------------------------------------- DDIM blocks
class QuantResnetBlock(BaseQuantBlock):
    """GN -> swish -> conv1 ; + temb_proj(swish(temb)) ; GN -> swish -> dropout -> conv2 ; (nin_shortcut) ; x + h
    (reference :392-444)."""

    def __init__(self, res, aq_params: dict = {}) -> None:
        super().__init__(aq_params)
        self.in_channels, self.out_channels = res.in_channels, res.out_channels
        self.use_conv_shortcut = res.use_conv_shortcut
        self.norm1, self.conv1, self.temb_proj = res.norm1, res.conv1, res.temb_proj
        self.norm2, self.dropout, self.conv2 = res.norm2, res.dropout, res.conv2
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                self.conv_shortcut = res.conv_shortcut
            else:
                self.nin_shortcut = res.nin_shortcut

    def forward(self, x: th.Tensor, temb: th.Tensor = None, split: int = 0):
        if temb is None:
            assert len(x) == 2
            x, temb = x
        h = self.conv1(nonlinearity(self.norm1(x)))
        h = h + self.temb_proj(nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x)
        return x + h


class QuantAttnBlock(BaseQuantBlock):
    def __init__(self, attn, aq_params: dict = {}, softmax_a_bit: int = 8) -> None:
        super().__init__(aq_params)
        self.in_channels = attn.in_channels
        self.norm, self.q, self.k, self.v, self.proj_out = attn.norm, attn.q, attn.k, attn.v, attn.proj_out
        self.aqtizer_q = UniformAffineQuantizer(**aq_params)
        self.aqtizer_k = UniformAffineQuantizer(**aq_params)
        self.aqtizer_v = UniformAffineQuantizer(**aq_params)
        self.aqtizer_w = _softmax_quantizer(aq_params, softmax_a_bit)

    def forward(self, x: th.Tensor) -> th.Tensor:
        hn = self.norm(x)
        q, k, v = self.q(hn), self.k(hn), self.v(hn)
        b, c, h, w = q.shape
        q = q.reshape(b, c, h * w).permute(0, 2, 1)
        k = k.reshape(b, c, h * w)
        if self.use_aq:
            q, k = self.aqtizer_q(q), self.aqtizer_k(k)
        att = nn.functional.softmax(th.bmm(q, k) * (int(c) ** (-0.5)), dim=2).permute(0, 2, 1)
        v = v.reshape(b, c, h * w)
        if self.use_aq:
            v, att = self.aqtizer_v(v), self.aqtizer_w(att)
        out = th.bmm(v, att).reshape(b, c, h, w)
        return x + self.proj_out(out)


def b2qb(use_aq: bool = False) -> Dict[str, type]:
    """FP block class name -> quantised block class (reference :508-520)."""
    table = {"ResBlock": QuantResBlock, "BasicTransformerBlock": QuantBasicTransformerBlock,
             "ResnetBlock": QuantResnetBlock, "AttnBlock": QuantAttnBlock}
    if use_aq:
        table["QKMatMul"] = QuantQKMatMul
        table["SMVMatMul"] = QuantSMVMatMul
    else:
        table["AttentionBlock"] = QuantAttentionBlock
    return table
