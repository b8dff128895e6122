"""ORACLE (test infrastructure, never shipped or benchmarked as the product).

CPU restatement, in plain functional torch fp32, of the reference's w4a8 fake-quantised UNet step
and DDIM sampler: the FP control flow of `ddim/models/diffusion.py:306-354` and
`ldm/modules/diffusionmodules/openaimodel.py:744-780` with every wrapped Conv2d / Linear evaluated
as `QuantLayer.forward` (quant/quant_layer.py:306-340).  It works directly on the FP model's
state_dict (reference key names) plus a `spec` that says, per wrapped layer, which quantisers are
live -- i.e. the result of QuantModel's surgery rules (quant/quant_model.py:56-66) and
`disable_out_quantization` (:103-120), restated in `wrapped_layer_names` / `build_spec`.

Pinned by tests/test_oracle_golden.py against tests/golden/{cifar,ldm4}_w4a8.pt, which were produced
by the reference itself (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import quant_ref as Q


# --------------------------------------------------------------------------- surgery rules
def wrapped_layer_names(sd: Dict[str, torch.Tensor]) -> List[str]:
    """Names of the Conv2d / Linear leaves QuantModel.quant_module wraps, in module order
    (quant/quant_model.py:56-66).  Conv1d weights (3-D) are never wrapped."""
    names = []
    for key, v in sd.items():
        if not key.endswith(".weight") or v.dim() not in (2, 4):
            continue
        path = key[: -len(".weight")].split(".")
        leaf, parent = path[-1], (path[-2] if len(path) > 1 else None)
        if "skip" in leaf or "op" in leaf or "shortcut" in leaf or (parent == "downsample" and leaf == "conv"):
            continue
        names.append(".".join(path))
    return names


def build_spec(sd, alpha_fn=None, level_w: int = 16):
    """Per wrapped layer: wq = (delta, zp, alpha|None) from channel-wise MINMAX (the scaler the
    sampling entry points use, ddim/runners/diffusion.py:248), and whether its input is quantised.
    Layers #0, #2 and the last stay fp; #1 and #3 keep weight quant only (quant_model.py:103-120)."""
    names = wrapped_layer_names(sd)
    fp = {names[0], names[2], names[-1]}
    no_aq = {names[1], names[3]}
    spec = {}
    for n in names:
        if n in fp:
            spec[n] = dict(wq=None, aq=False)
            continue
        w = sd[n + ".weight"]
        delta, zp = Q.channel_wise(Q.minmax_scale, w, level_w)
        alpha = alpha_fn(n, w, delta) if alpha_fn is not None else None
        spec[n] = dict(wq=(delta, zp, alpha), aq=n not in no_aq)
    return spec


class ActParams:
    """Finite-Set-Calibration table: activation (delta, zp) per layer for one sampling step."""

    def __init__(self, names: List[str], row: torch.Tensor):
        self.p = {n: (row[i, 0], row[i, 1]) for i, n in enumerate(names)}

    def get(self, name, x=None):
        return self.p.get(name)


class CalibratingActParams(ActParams):
    """Lazy MINMAX initialisation on first use, like UniformAffineQuantizer's first forward
    (quant/quant_layer.py:213-217); afterwards the parameters are frozen."""

    def __init__(self):
        self.p = {}

    def get(self, name, x=None):
        if name not in self.p:
            self.p[name] = Q.minmax_scale(x, 256)
        return self.p[name]


class _Net:
    def __init__(self, sd, spec, act: Optional[ActParams], record: Optional[dict] = None):
        self.sd, self.spec, self.act, self.record = sd, spec, act, record

    def layer(self, name, x, conv=None):
        w, b = self.sd[name + ".weight"], self.sd.get(name + ".bias")
        s = self.spec.get(name)
        if s is None:                      # a leaf the reference never wraps
            if w.dim() == 3:
                return F.conv1d(x, w, b)
            return F.linear(x, w, b) if conv is None else F.conv2d(x, w, b, **conv)
        aq = self.act.get(name, x) if (s["aq"] and self.act is not None) else None
        if s["aq"] and self.act is not None and aq is None:
            raise KeyError(f"no activation quant parameters for {name}")
        if self.record is not None and aq is not None and x.dim() in (3, 4):
            # conv inputs [b, c, h, w] and token / context inputs [b, n, c] of the transformer linears
            self.record[name] = Q.uaq_codes(x, aq[0], aq[1], 256).to(torch.uint8)
        y = Q.quant_layer_forward(x, w, b, wq=s["wq"], aq=aq, conv=conv)
        if self.record is not None and x.dim() == 2:
            self.record["out:" + name] = y          # time-embedding MLP outputs (teacher forcing)
        return y

    def mark(self, name, y):
        if self.record is not None:
            self.record["blk:" + name] = y
        return y

    def gn(self, name, x, eps):
        return F.group_norm(x, 32, self.sd[name + ".weight"], self.sd[name + ".bias"], eps)


P1 = dict(stride=1, padding=1)
P0 = dict(stride=1, padding=0)


# --------------------------------------------------------------------------- DDIM UNet
def ddim_timestep_embedding(t, dim):
    """ddim/models/diffusion.py:6-24."""
    half = dim // 2
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    emb = t.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def ddim_unet_forward(sd, cfg, x, t, spec, act: Optional[ActParams] = None, record=None):
    """ddim/models/diffusion.py:306-354 with quant/quant_block.py:415-444 (QuantResnetBlock) and
    :474-505 (QuantAttnBlock, attention core in fp32)."""
    net = _Net(sd, spec, act, record)
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    nres = len(mult)
    has = lambda k: (k + ".weight") in sd  # noqa: E731

    def resblock(p, x, temb):
        h = net.layer(p + ".conv1", Q.silu(net.gn(p + ".norm1", x, 1e-6)), P1)
        h = h + net.layer(p + ".temb_proj", Q.silu(temb))[:, :, None, None]
        h = net.layer(p + ".conv2", Q.silu(net.gn(p + ".norm2", h, 1e-6)), P1)
        if has(p + ".nin_shortcut"):
            x = net.layer(p + ".nin_shortcut", x, P0)
        return net.mark(p, x + h)

    def attn(p, x):
        hn = net.gn(p + ".norm", x, 1e-6)
        q, k, v = (net.layer(p + "." + n, hn, P0) for n in "qkv")
        b, c, h, w = q.shape
        q = q.reshape(b, c, h * w).permute(0, 2, 1)
        k = k.reshape(b, c, h * w)
        w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
        h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
        return net.mark(p, x + net.layer(p + ".proj_out", h_, P0))

    temb = net.layer("temb.dense.0", ddim_timestep_embedding(t, ch))
    temb = net.layer("temb.dense.1", Q.silu(temb))
    hs = [net.mark("conv_in", net.layer("conv_in", x, P1))]
    for l in range(nres):
        for j in range(nrb):
            h = resblock(f"down.{l}.block.{j}", hs[-1], temb)
            if has(f"down.{l}.attn.{j}.q"):
                h = attn(f"down.{l}.attn.{j}", h)
            hs.append(h)
        if l != nres - 1:
            hs.append(net.mark(f"down.{l}.downsample", net.layer(f"down.{l}.downsample.conv",
                                                                 F.pad(hs[-1], (0, 1, 0, 1)), dict(stride=2, padding=0))))
    h = resblock("mid.block_1", hs[-1], temb)
    h = attn("mid.attn_1", h)
    h = resblock("mid.block_2", h, temb)
    for l in reversed(range(nres)):
        for j in range(nrb + 1):
            h = resblock(f"up.{l}.block.{j}", torch.cat([h, hs.pop()], dim=1), temb)
            if has(f"up.{l}.attn.{j}.q"):
                h = attn(f"up.{l}.attn.{j}", h)
        if l != 0:
            h = net.mark(f"up.{l}.upsample", net.layer(f"up.{l}.upsample.conv",
                                                       F.interpolate(h, scale_factor=2.0, mode="nearest"), P1))
    return net.layer("conv_out", net.mark("final_act", Q.silu(net.gn("norm_out", h, 1e-6))), P1)


# --------------------------------------------------------------------------- LDM UNet
def ldm_timestep_embedding(t, dim):
    """ldm/modules/diffusionmodules/util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def ldm_unet_forward(sd, cfg, x, t, spec, act: Optional[ActParams] = None, record=None, context=None):
    """openaimodel.py:744-780 with quant/quant_block.py:178-209 (QuantResBlock) and the fp
    AttentionBlock / QKVAttentionLegacy (:320-326, :383-405); for SpatialTransformer UNets (SD v1.4, cin256) the
    transformer path of ldm/modules/attention.py:250-261 with quant/quant_block.py:212-299
    (QuantBasicTransformerBlock / cross_attn_forward: quantised projections, fp32 attention core -- the block's
    own q/k/v/softmax quantisers are inert).  Block structure is read off the state_dict keys."""
    net = _Net(sd, spec, act, record)
    mc = cfg["model_channels"]
    has = lambda k: (k + ".weight") in sd  # noqa: E731

    def resblock(p, x, emb):
        h = net.layer(p + ".in_layers.2", F.silu(net.gn(p + ".in_layers.0", x, 1e-5)), P1)
        h = h + net.layer(p + ".emb_layers.1", F.silu(emb))[:, :, None, None]
        h = net.layer(p + ".out_layers.3", F.silu(net.gn(p + ".out_layers.0", h, 1e-5)), P1)
        if has(p + ".skip_connection"):
            x = net.layer(p + ".skip_connection", x, P0)
        return net.mark(p, x + h)

    def attn(p, x, heads_ch):
        b, c, hh, ww = x.shape
        xf = x.reshape(b, c, -1)
        qkv = net.layer(p + ".qkv", net.gn(p + ".norm", xf, 1e-5))
        nh = c // heads_ch
        bs, width, length = qkv.shape
        chd = width // (3 * nh)
        q, k, v = qkv.reshape(bs * nh, chd * 3, length).split(chd, dim=1)
        scale = 1 / math.sqrt(math.sqrt(chd))
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), dim=-1)
        a = torch.einsum("bts,bcs->bct", w, v).reshape(bs, -1, length)
        return net.mark(p, (xf + net.layer(p + ".proj_out", a)).reshape(b, c, hh, ww))

    def cross_attn(p, xq, ctx, heads):
        """quant_block.py:212-245 with use_aq False on the block."""
        q = net.layer(p + ".to_q", xq)
        ctx = xq if ctx is None else ctx
        k, v = net.layer(p + ".to_k", ctx), net.layer(p + ".to_v", ctx)
        b, n, inner = q.shape
        dh = inner // heads

        def split(t_):
            return t_.reshape(b, t_.shape[1], heads, dh).permute(0, 2, 1, 3).reshape(b * heads, t_.shape[1], dh)

        q, k, v = split(q), split(k), split(v)
        attn = (torch.einsum("bid,bjd->bij", q, k) * (dh ** -0.5)).softmax(dim=-1)
        out = torch.einsum("bij,bjd->bid", attn, v)
        out = out.reshape(b, heads, n, dh).permute(0, 2, 1, 3).reshape(b, n, inner)
        return net.layer(p + ".to_out.0", out)

    def ln(p, t_):
        return F.layer_norm(t_, (t_.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)

    def spatial_transformer(p, x):
        """attention.py:250-261; GroupNorm eps 1e-6 (attention.py:76-77)."""
        b, c, hh, ww = x.shape
        h = net.layer(p + ".proj_in", net.gn(p + ".norm", x, 1e-6), P0)
        inner = h.shape[1]
        heads = cfg["num_heads"] if cfg.get("num_head_channels", -1) == -1 else c // cfg["num_head_channels"]
        tok = h.reshape(b, inner, hh * ww).permute(0, 2, 1)
        d = 0
        while has(f"{p}.transformer_blocks.{d}.attn1.to_q"):
            bp = f"{p}.transformer_blocks.{d}"
            tok = cross_attn(bp + ".attn1", ln(bp + ".norm1", tok), None, heads) + tok
            tok = cross_attn(bp + ".attn2", ln(bp + ".norm2", tok), context, heads) + tok
            y = net.layer(bp + ".ff.net.0.proj", ln(bp + ".norm3", tok))
            a, gate = y.chunk(2, dim=-1)
            tok = net.layer(bp + ".ff.net.2", a * F.gelu(gate)) + tok
            net.mark(bp, tok)
            d += 1
        h = tok.permute(0, 2, 1).reshape(b, inner, hh, ww)
        return net.mark(p, net.layer(p + ".proj_out", h, P0) + x)

    def run_block(p, h, emb):
        j = 0
        while True:
            q = f"{p}.{j}"
            if has(q + ".in_layers.2"):
                h = resblock(q, h, emb)
            elif has(q + ".transformer_blocks.0.attn1.to_q"):
                h = spatial_transformer(q, h)
            elif has(q + ".qkv"):
                h = attn(q, h, cfg["num_head_channels"])
            elif has(q + ".op"):
                h = net.mark(q, net.layer(q + ".op", h, dict(stride=2, padding=1)))
            elif has(q + ".conv"):
                h = net.mark(q, net.layer(q + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"), P1))
            elif has(q):
                h = net.mark(q, net.layer(q, h, P1))           # input_blocks.0.0
            else:
                return h
            j += 1

    emb = net.layer("time_embed.0", ldm_timestep_embedding(t, mc))
    emb = net.layer("time_embed.2", F.silu(emb))
    hs, h, i = [], x, 0
    while has(f"input_blocks.{i}.0") or has(f"input_blocks.{i}.0.in_layers.2") or has(f"input_blocks.{i}.0.op"):
        h = run_block(f"input_blocks.{i}", h, emb)
        hs.append(h)
        i += 1
    h = run_block("middle_block", h, emb)
    i = 0
    while has(f"output_blocks.{i}.0.in_layers.2"):
        h = run_block(f"output_blocks.{i}", torch.cat([h, hs.pop()], dim=1), emb)
        i += 1
    return net.layer("out.2", net.mark("final_act", F.silu(net.gn("out.0", h, 1e-5))), P1)


# --------------------------------------------------------------------------- DDIM sampler
def compute_alpha(betas, t):
    """ddim/functions/denoising.py:4-7."""
    beta = torch.cat([torch.zeros(1), betas], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def generalized_steps(x, seq, eps_fn, betas, eta: float = 0.0, noise_fn=None):
    """ddim/functions/denoising.py:10-41.  eps_fn(xt, t, step_index) -> predicted noise; the step
    index selects the FSC activation table (`act_{cnt}`)."""
    n = x.size(0)
    seq_next = [-1] + list(seq[:-1])
    xs, x0_preds = [x], []
    for cnt, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
        t = torch.ones(n) * i
        next_t = torch.ones(n) * j
        at = compute_alpha(betas, t.long())
        at_next = compute_alpha(betas, next_t.long())
        xt = xs[-1]
        et = eps_fn(xt, t, cnt)
        x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
        x0_preds.append(x0_t)
        c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
        c2 = ((1 - at_next) - c1 ** 2).sqrt()
        noise = noise_fn(x) if noise_fn is not None else torch.randn_like(x)
        xs.append(at_next.sqrt() * x0_t + c1 * noise + c2 * et)
    return xs, x0_preds


def ddim_coef_table(seq, betas, eta: float = 0.0):
    """Per-step (sqrt(a_t), sqrt(1-a_t), sqrt(a_next), c2, c1) in fp32, computed exactly as the
    tensors of generalized_steps are (so a device kernel fed these reproduces the update bit for bit)."""
    rows = []
    seq_next = [-1] + list(seq[:-1])
    for i, j in zip(reversed(seq), reversed(seq_next)):
        at = compute_alpha(betas, torch.tensor([i])).reshape(())
        an = compute_alpha(betas, torch.tensor([j])).reshape(())
        c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
        c2 = ((1 - an) - c1 ** 2).sqrt()
        rows.append([at.sqrt().item(), (1 - at).sqrt().item(), an.sqrt().item(), c2.item(), float(c1)])
    return rows
