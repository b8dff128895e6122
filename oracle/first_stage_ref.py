"""ORACLE (test infrastructure, never shipped or benchmarked as the product).

CPU restatement, in plain functional torch fp32, of the first-stage decode that follows the sampling path (SURVEY 8(f)
f3): `LatentDiffusion.decode_first_stage` (reference ldm/models/diffusion/ddpm.py:706-764, un-split branch),
`VQModelInterface.decode` (ldm/models/autoencoder.py:274-283), `AutoencoderKL.decode` (:330-333) and `Decoder.forward`
(ldm/modules/diffusionmodules/model.py:535-568) with its `ResnetBlock` (:121-141), `AttnBlock` (:178-202) and `Upsample`
(:53-57).  It works directly on a state_dict with the reference's key names.

The nearest-codebook lookup is NOT in the reference tree: `ldm/models/autoencoder.py:6` imports
`taming.modules.vqvae.quantize.VectorQuantizer2` from taming-transformers (CompVis, `-e git+...taming-transformers.git@master`
in stable-diffusion/environment.yaml, i.e. unpinned).  `vq_lookup` restates its published forward:
    d = sum(z^2, 1, keepdim) + sum(E^2, 1) - 2 einsum('bd,dn->bn', z, E^T);  idx = argmin(d, 1);  z_q = E[idx]
    z_q = z + (z_q - z).detach()                      (legacy / beta terms only touch the loss)
PARITY of that one function is UNPINNED (no golden from taming itself); everything downstream of it -- post_quant_conv and
the decoder -- is pinned by tests/test_oracle_golden.py against tests/golden/first_stage.pt, produced by running the
reference's own `VQModelInterface.decode` / `AutoencoderKL.decode` (tests/golden/make_golden.py::first_stage_golden).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, sd, name, pad):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=pad)


def resnet_block(x, sd, p):
    """model.py:121-141 with temb = None."""
    h = _conv(_swish(_gn(x, sd, p + ".norm1")), sd, p + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, p + ".norm2")), sd, p + ".conv2", 1)     # dropout: eval -> identity
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".conv_shortcut", 1)
    elif p + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".nin_shortcut", 0)
    return x + h


def attn_block(x, sd, p):
    """model.py:178-202."""
    h_ = _gn(x, sd, p + ".norm")
    q, k, v = (_conv(h_, sd, p + "." + n, 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, p + ".proj_out", 0)


def decoder_forward(z, sd: Dict[str, torch.Tensor], prefix: str = "decoder") -> torch.Tensor:
    """model.py:535-568; the level / block structure is read off the keys."""
    p = prefix
    h = _conv(z, sd, p + ".conv_in", 1)
    h = resnet_block(h, sd, p + ".mid.block_1")
    h = attn_block(h, sd, p + ".mid.attn_1")
    h = resnet_block(h, sd, p + ".mid.block_2")
    levels = 1 + max(int(k.split(".")[len(p.split(".")) + 1]) for k in sd if k.startswith(p + ".up."))
    for lvl in reversed(range(levels)):
        j = 0
        while f"{p}.up.{lvl}.block.{j}.conv1.weight" in sd:
            h = resnet_block(h, sd, f"{p}.up.{lvl}.block.{j}")
            if f"{p}.up.{lvl}.attn.{j}.q.weight" in sd:
                h = attn_block(h, sd, f"{p}.up.{lvl}.attn.{j}")
            j += 1
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            if f"{p}.up.{lvl}.upsample.conv.weight" in sd:
                h = _conv(h, sd, f"{p}.up.{lvl}.upsample.conv", 1)
    h = _swish(_gn(h, sd, p + ".norm_out"))
    return _conv(h, sd, p + ".conv_out", 1)


def vq_lookup(z: torch.Tensor, codebook: torch.Tensor):
    """taming VectorQuantizer2.forward (see the header).  z: [b, c, h, w]; returns (z_q [b, c, h, w], indices [b*h*w])."""
    zl = z.permute(0, 2, 3, 1).contiguous()
    zf = zl.view(-1, codebook.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) \
        - 2 * torch.einsum("bd,dn->bn", zf, codebook.t())
    idx = torch.argmin(d, dim=1)
    z_q = codebook[idx].view(zl.shape)
    z_q = zl + (z_q - zl)
    return z_q.permute(0, 3, 1, 2).contiguous(), idx


def first_stage_decode(z: torch.Tensor, sd: Dict[str, torch.Tensor], quantize: bool) -> torch.Tensor:
    """`first_stage_model.decode(z)`: [VQ lookup] -> post_quant_conv -> decoder."""
    if quantize:
        z, _ = vq_lookup(z, sd["quantize.embedding.weight"])
    z = _conv(z, sd, "post_quant_conv", 0)
    return decoder_forward(z, sd)


def decode_first_stage(z: torch.Tensor, sd: Dict[str, torch.Tensor], scale_factor: float, quantize: bool,
                       force_not_quantize: bool = False) -> torch.Tensor:
    """ddpm.py:706-764 (no split_input_params): `z = 1. / self.scale_factor * z`, then decode."""
    z = 1. / scale_factor * z
    return first_stage_decode(z, sd, quantize and not force_not_quantize)
