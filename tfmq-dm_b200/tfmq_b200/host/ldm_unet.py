"""Host FP UNet of the latent-diffusion pipeline (LDM-4 CelebA-HQ / LSUN configs).

Same module tree and state_dict keys as `ldm/modules/diffusionmodules/openaimodel.py:446-780`
(`time_embed.{0,2}`, `input_blocks.{i}.{j}`, `middle_block.{j}`, `output_blocks.{i}.{j}`, `out.{0,2}`;
ResBlock = `in_layers.{0,2}`, `emb_layers.1`, `out_layers.{0,3}`, `skip_connection`; AttentionBlock =
`norm`, `qkv`, `proj_out`; `Downsample.op`; `Upsample.conv`), so reference checkpoints load unchanged
and QuantModel's name rules (quant/quant_model.py:56-66: 'skip' / 'op' are left in fp; `emb_layers.1`
is a quant_emb layer; Conv1d is never wrapped) apply as they do to the reference graph.

SpatialTransformer UNets (SD v1.4, cin256; `use_spatial_transformer`, `context_dim`) use the transformer modules
of `ldm/modules/attention.py:37-261` re-expressed below with the same keys (`norm`, `proj_in`,
`transformer_blocks.{d}.{attn1,attn2}.{to_q,to_k,to_v,to_out.0}`, `ff.net.{0.proj,2}`, `norm{1,2,3}`, `proj_out`).

Only the options the four benchmark configs use are implemented (no scale-shift norm, no
resblock up/down, legacy attention order).
"""
from __future__ import annotations

import math
from abc import abstractmethod

import torch
import torch.nn as nn
import torch.nn.functional as F


_FREQ: dict = {}


def _frequencies(half: int, max_period: int, device) -> torch.Tensor:
    """The frequency row, evaluated on the host as the reference does and kept per device: no pageable upload per call (a
    captured reconstruction iteration, quant/reconstruction.py, could not contain one)."""
    key = (half, max_period, str(device))
    f = _FREQ.get(key)
    if f is None:
        f = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(device)
        _FREQ[key] = f
    return f


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000, repeat_only: bool = False):
    """[cos | sin] sinusoid, frequencies exp(-ln(max_period) i / half) (ldm/.../util.py:151-171)."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    args = timesteps[:, None].float() * _frequencies(half, max_period, timesteps.device)[None]
    emb = torch.cat([args.cos(), args.sin()], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def normalization(channels: int) -> nn.GroupNorm:
    return GroupNorm32(32, channels)


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.detach().zero_()
    return m


def checkpoint(func, inputs, params, flag):
    """Inference / PTQ never needs activation re-computation: run the function directly
    (the reference's `checkpoint(..., flag)` is numerically the identity wrapper, util.py:102-148)."""
    return func(*inputs)


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        ...


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None, split=0):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb, split=split)
            elif layer.__class__.__name__ == "SpatialTransformer":      # defined below (openaimodel.py:84-85)
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.conv(x) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding) if use_conv \
            else nn.AvgPool2d(2, 2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv:
            raise NotImplementedError("scale-shift norm / resblock up-down are not used by the supported configs")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, False
        self.updown = False
        self.h_upd = self.x_upd = nn.Identity()
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        self.skip_connection = nn.Identity() if self.out_channels == channels \
            else nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb, split=0):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb).type(h.dtype)[:, :, None, None]
        return self.skip_connection(x) + self.out_layers(h)


class QKMatMul(nn.Module):
    def __init__(self):
        super().__init__()
        self.scale = None

    def forward(self, q, k):
        return torch.einsum("bct,bcs->bts", q * self.scale, k * self.scale)


class SMVMatMul(nn.Module):
    def forward(self, weight, v):
        return torch.einsum("bts,bcs->bct", weight, v)


class QKVAttentionLegacy(nn.Module):
    """qkv: [N, H*3*C, T], heads split before q/k/v (openaimodel.py:372-409)."""

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads
        self.qkv_matmul, self.smv_matmul = QKMatMul(), SMVMatMul()

    def forward(self, qkv):
        bs, width, length = qkv.shape
        ch = width // (3 * self.n_heads)
        q, k, v = qkv.reshape(bs * self.n_heads, ch * 3, length).split(ch, dim=1)
        self.qkv_matmul.scale = 1 / math.sqrt(math.sqrt(ch))
        w = torch.softmax(self.qkv_matmul(q, k).float(), dim=-1).type(qkv.dtype)
        return self.smv_matmul(w, v).reshape(bs, -1, length)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False):
        super().__init__()
        if use_new_attention_order:
            raise NotImplementedError("only the legacy attention order is used by the supported configs")
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.attention = QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, *spatial = x.shape
        x = x.reshape(b, c, -1)
        h = self.proj_out(self.attention(self.qkv(self.norm(x))))
        return (x + h).reshape(b, c, *spatial)


# ------------------------------------------------------------------ transformer blocks (SD v1.4, cin256)
class GEGLU(nn.Module):
    """x, gate = proj(x).chunk(2); x * gelu(gate)   (ldm/modules/attention.py:37-44)."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class FeedForward(nn.Module):
    """net = [GEGLU(dim, 4 dim), Dropout, Linear(4 dim, dim)]   (attention.py:47-66; glu=True in the configs)."""

    def __init__(self, dim: int, dim_out=None, mult: int = 4, glu: bool = True, dropout: float = 0.0):
        super().__init__()
        if not glu:
            raise NotImplementedError("the benchmark configs use the gated feed-forward")
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim))

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):
    """Multi-head attention of tokens over `context` (self-attention when context is None), attention.py:152-194.
    QuantBasicTransformerBlock replaces `forward` by quant_block.cross_attn_forward."""

    def __init__(self, query_dim: int, context_dim=None, heads: int = 8, dim_head: int = 64, dropout: float = 0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None, mask=None):
        if mask is not None:
            raise NotImplementedError("attention masks are not used on the sampling path")
        b, n, _ = x.shape
        ctx = x if context is None else context
        h = self.heads

        def split(t):
            return t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], -1)

        q, k, v = split(self.to_q(x)), split(self.to_k(ctx)), split(self.to_v(ctx))
        attn = (torch.einsum("bid,bjd->bij", q, k) * self.scale).softmax(dim=-1)
        out = torch.einsum("bij,bjd->bid", attn, v)
        return self.to_out(out.reshape(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1))


class BasicTransformerBlock(nn.Module):
    """x += attn1(norm1 x); x += attn2(norm2 x, context); x += ff(norm3 x)   (attention.py:196-216)."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        return self.ff(self.norm3(x)) + x

    _forward = forward


class SpatialTransformer(nn.Module):
    """GroupNorm(32, eps 1e-6) -> 1x1 proj_in -> tokens [b, hw, c] -> transformer blocks -> 1x1 proj_out -> + x
    (attention.py:218-261)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None):
        super().__init__()
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim)
             for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner, in_channels, 1))

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        t = self.proj_in(self.norm(x))
        inner = t.shape[1]
        t = t.reshape(b, inner, h * w).permute(0, 2, 1)
        for blk in self.transformer_blocks:
            t = blk(t, context=context)
        t = t.permute(0, 2, 1).reshape(b, inner, h, w)
        return self.proj_out(t) + x


def sd_v14_config() -> dict:
    """unet_config.params of stable-diffusion/configs/stable-diffusion/v1-inference.yaml:29-44 (the YAML's
    `image_size: 32` is marked unused there; 64 is the latent extent of the 512x512 txt2img config)."""
    return dict(image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                transformer_depth=1, context_dim=768, legacy=False)


def cin256_config() -> dict:
    """unet_config.params of stable-diffusion/configs/latent-diffusion/cin256-v2.yaml:19-39."""
    return dict(image_size=64, in_channels=3, out_channels=3, model_channels=192, attention_resolutions=[8, 4, 2],
                num_res_blocks=2, channel_mult=[1, 2, 3, 5], num_heads=1, use_spatial_transformer=True,
                transformer_depth=1, context_dim=512)


def sd_mini_config() -> dict:
    """A small UNet with the structure of the SD v1.4 one (SpatialTransformer at every attention resolution, 2 heads,
    cross-attention over a 7-token context) that the CPU oracle runs in seconds: the parity-test stand-in for
    BASELINE configs[2] / [4]."""
    return dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
                num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True,
                transformer_depth=1, context_dim=96, legacy=False)


def celebahq_ldm4_config() -> dict:
    """unet_config.params of stable-diffusion/models/ldm/celeba256/config.yaml:17-34."""
    return dict(image_size=64, in_channels=3, out_channels=3, model_channels=224,
                attention_resolutions=[8, 4, 2], num_res_blocks=2, channel_mult=[1, 2, 3, 4],
                num_head_channels=32)


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        if use_spatial_transformer != (context_dim is not None):
            raise ValueError("use_spatial_transformer and context_dim go together (openaimodel.py:497-505)")
        if isinstance(context_dim, (list, tuple)):
            context_dim = list(context_dim)[0]
        if dims != 2 or num_classes is not None or resblock_updown or use_fp16 or n_embed is not None:
            raise NotImplementedError("option not used by the supported configs")
        assert num_heads != -1 or num_head_channels != -1
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.channel_mult = attention_resolutions, channel_mult
        self.num_classes, self.dtype, self.split = None, torch.float32, False
        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))

        def attn(ch):
            # head geometry as openaimodel.py:576-592 (legacy=True: dim_head = ch // num_heads for transformers)
            if num_head_channels == -1:
                nh, dh = num_heads, ch // num_heads
            else:
                nh, dh = ch // num_head_channels, num_head_channels
            if legacy:
                dh = ch // nh if use_spatial_transformer else num_head_channels
            if use_spatial_transformer:
                return SpatialTransformer(ch, nh, dh, depth=transformer_depth, context_dim=context_dim)
            return AttentionBlock(ch, num_heads=nh, num_head_channels=dh)

        def res(cin, cout):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        skip_chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        hs, h = [], x.type(self.dtype)
        for module in self.input_blocks:
            h = module(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for module in self.output_blocks:
            h = module(torch.cat([h, hs.pop()], dim=1), emb, context)
        return self.out(h.type(x.dtype))
