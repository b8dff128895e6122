"""Host FP UNet of the DDIM (CIFAR-10 / CelebA / LSUN) pipeline.

Same module tree and state_dict keys as the reference's `ddim/models/diffusion.py:192-354`
(`temb.dense.{0,1}`, `conv_in`, `down.{l}.block.{b}.{norm1,conv1,temb_proj,norm2,conv2,nin_shortcut}`,
`down.{l}.attn.{b}.{norm,q,k,v,proj_out}`, `down.{l}.downsample.conv`, `mid.{block_1,attn_1,block_2}`,
`up.{l}...`, `norm_out`, `conv_out`) so real DDIM checkpoints load unchanged and QuantModel's
name-based wrapping rules (quant/quant_model.py:56-66) see the same names.  This is glue between the
quantised leaves: at sampling time the step engine replaces this forward with one fused kernel
program; the torch forward here is the FP path used while calibrating.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


_FREQ: dict = {}


def _frequencies(half: int, device) -> torch.Tensor:
    """The frequency row, evaluated on the host as the reference does and kept per device: no pageable upload per call (a
    captured reconstruction iteration, quant/reconstruction.py, could not contain one)."""
    key = (half, str(device))
    f = _FREQ.get(key)
    if f is None:
        f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(device)
        _FREQ[key] = f
    return f


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int) -> torch.Tensor:
    """[sin | cos] sinusoid with frequencies exp(-ln(1e4) i / (half-1)) (ddim/models/diffusion.py:6-24)."""
    assert timesteps.dim() == 1
    half = embedding_dim // 2
    arg = timesteps.float()[:, None] * _frequencies(half, timesteps.device)[None, :]
    emb = torch.cat([arg.sin(), arg.cos()], dim=1)
    return F.pad(emb, (0, 1, 0, 0)) if embedding_dim % 2 == 1 else emb


def nonlinearity(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(x)


def Normalize(channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, channels: int, with_conv: bool):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(channels, channels, 3, 1, 1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, channels: int, with_conv: bool):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(channels, channels, 3, 2, 0)  # pads right/bottom by hand below

    def forward(self, x):
        if self.with_conv:
            return self.conv(F.pad(x, (0, 1, 0, 1)))
        return F.avg_pool2d(x, 2, 2)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels: int, out_channels: int | None = None, conv_shortcut: bool = False,
                 dropout: float = 0.0, temb_channels: int = 512):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, x, temb, split=0):
        h = self.conv1(nonlinearity(self.norm1(x)))
        h = h + self.temb_proj(nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)

    def forward(self, x):
        hn = self.norm(x)
        b, c, h, w = x.shape
        q = self.q(hn).reshape(b, c, h * w).permute(0, 2, 1)
        k = self.k(hn).reshape(b, c, h * w)
        v = self.v(hn).reshape(b, c, h * w)
        att = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
        out = torch.bmm(v, att.permute(0, 2, 1)).reshape(b, c, h, w)
        return x + self.proj_out(out)


def cifar10_config() -> SimpleNamespace:
    """Hyper-parameters of ddim/configs/cifar10.yml:12-24 (the fields Model reads)."""
    return SimpleNamespace(
        model=SimpleNamespace(type="simple", in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 2, 2],
                              num_res_blocks=2, attn_resolutions=[16], dropout=0.1, resamp_with_conv=True),
        data=SimpleNamespace(image_size=32, channels=3),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                                  num_diffusion_timesteps=1000),
        split_shortcut=False,
    )


class Model(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        m = config.model
        ch, ch_mult = m.ch, tuple(m.ch_mult)
        self.ch, self.temb_ch = ch, ch * 4
        self.num_resolutions, self.num_res_blocks = len(ch_mult), m.num_res_blocks
        self.resolution, self.in_channels = config.data.image_size, m.in_channels
        if m.type == "bayesian":
            self.logvar = nn.Parameter(torch.zeros(config.diffusion.num_diffusion_timesteps))

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        self.conv_in = nn.Conv2d(m.in_channels, ch, 3, 1, 1)

        res = self.resolution
        widths = [ch * k for k in (1,) + ch_mult]        # widths[l] feeds level l, widths[l+1] leaves it
        self.down = nn.ModuleList()
        cur = widths[0]
        for lvl in range(self.num_resolutions):
            stage = nn.Module()
            stage.block, stage.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(self.num_res_blocks):
                stage.block.append(ResnetBlock(in_channels=cur, out_channels=widths[lvl + 1],
                                               temb_channels=self.temb_ch, dropout=m.dropout))
                cur = widths[lvl + 1]
                if res in m.attn_resolutions:
                    stage.attn.append(AttnBlock(cur))
            if lvl != self.num_resolutions - 1:
                stage.downsample = Downsample(cur, m.resamp_with_conv)
                res //= 2
            self.down.append(stage)

        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=cur, out_channels=cur, temb_channels=self.temb_ch, dropout=m.dropout)
        self.mid.attn_1 = AttnBlock(cur)
        self.mid.block_2 = ResnetBlock(in_channels=cur, out_channels=cur, temb_channels=self.temb_ch, dropout=m.dropout)

        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            stage = nn.Module()
            stage.block, stage.attn = nn.ModuleList(), nn.ModuleList()
            out_w = widths[lvl + 1]
            for j in range(self.num_res_blocks + 1):
                skip_w = widths[lvl] if j == self.num_res_blocks else widths[lvl + 1]
                stage.block.append(ResnetBlock(in_channels=cur + skip_w, out_channels=out_w,
                                               temb_channels=self.temb_ch, dropout=m.dropout))
                cur = out_w
                if res in m.attn_resolutions:
                    stage.attn.append(AttnBlock(cur))
            if lvl != 0:
                stage.upsample = Upsample(cur, m.resamp_with_conv)
                res *= 2
            self.up.insert(0, stage)

        self.norm_out = Normalize(cur)
        self.conv_out = nn.Conv2d(cur, m.out_ch, 3, 1, 1)

    def forward(self, x, t):
        assert x.shape[2] == x.shape[3] == self.resolution
        temb = get_timestep_embedding(t, self.ch)
        temb = self.temb.dense[1](nonlinearity(self.temb.dense[0](temb)))
        hs = [self.conv_in(x)]
        for lvl, stage in enumerate(self.down):
            for j in range(self.num_res_blocks):
                h = stage.block[j](hs[-1], temb)
                if len(stage.attn) > 0:
                    h = stage.attn[j](h)
                hs.append(h)
            if lvl != self.num_resolutions - 1:
                hs.append(stage.downsample(hs[-1]))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], temb)), temb)
        for lvl in reversed(range(self.num_resolutions)):
            stage = self.up[lvl]
            for j in range(self.num_res_blocks + 1):
                h = stage.block[j](torch.cat([h, hs.pop()], dim=1), temb)
                if len(stage.attn) > 0:
                    h = stage.attn[j](h)
            if lvl != 0:
                h = stage.upsample(h)
        return self.conv_out(nonlinearity(self.norm_out(h)))
