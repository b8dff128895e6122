"""First-stage decode: the step after the sampling path -- latents -> images.

Replaces `LatentDiffusion.decode_first_stage` (reference ldm/models/diffusion/ddpm.py:706-764, the un-split branch),
`VQModelInterface.decode` (ldm/models/autoencoder.py:274-283) / `AutoencoderKL.decode` (:330-333) and the `Decoder` they
run (ldm/modules/diffusionmodules/model.py:462-568: conv_in, mid.block_1 / attn_1 / block_2, up levels of ResnetBlocks with
nearest x2 `Upsample` convs, norm_out, swish, conv_out).  The nearest-codebook lookup of the VQ models is
taming-transformers' `VectorQuantizer2.forward` (a dependency the reference does not vendor; restated in oracle/).

Two pieces:

* `FirstStageModel` -- a parameter container with the reference's state_dict keys (`decoder.*`, `post_quant_conv.*`,
  `quantize.embedding.weight`), so `load_state_dict(ckpt["state_dict"], strict=False)` of a real vq-f4 / kl-f8
  checkpoint fills it.  It has no torch forward: the floating-point decoder only runs as
* `DecoderEngine` -- a static program of this library's sm_100a kernels (the same fp16 hi/lo-split tensor-core convs,
  GroupNorm + SiLU producers and attention the step engine uses for the layers the reference keeps in floating point),
  replayed through a CUDA graph.  There is no CPU path.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .engine import StepEngine, T


# ------------------------------------------------------------------------------------------ configurations
def vq_f4_config() -> dict:
    """First stage of LDM-4 CelebA-HQ (BASELINE configs[1]): reference models/first_stage_models/vq-f4/config.yaml and
    models/ldm/celeba256/config.yaml (first_stage_config: VQModelInterface, embed_dim 3, n_embed 8192)."""
    return dict(embed_dim=3, n_embed=8192, scale_factor=1.0,
                ddconfig=dict(double_z=False, z_channels=3, resolution=256, in_channels=3, out_ch=3, ch=128,
                              ch_mult=(1, 2, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0))


def kl_f8_config() -> dict:
    """First stage of SD v1.4 (configs[2]): configs/stable-diffusion/v1-inference.yaml:17,46-67 (AutoencoderKL)."""
    return dict(embed_dim=4, n_embed=None, scale_factor=0.18215,
                ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                              ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0))


def first_stage_mini_config(kind: str = "vq") -> dict:
    """Small decoders with the structure of the two above (tests / golden fixtures): 64-channel mid attention, one
    upsampling level, a nin_shortcut; "vq-attn" also has AttnBlocks inside an up level."""
    dd = dict(double_z=kind == "kl", z_channels=4 if kind == "kl" else 3, resolution=32, in_channels=3, out_ch=3, ch=32,
              ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=[16] if kind == "vq-attn" else [], dropout=0.0)
    if kind == "kl":
        return dict(embed_dim=4, n_embed=None, scale_factor=0.18215, ddconfig=dd)
    return dict(embed_dim=3, n_embed=96, scale_factor=1.0, ddconfig=dd)


# ------------------------------------------------------------------------------------------ parameter containers
def Normalize(channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, channels, eps=1e-6, affine=True)        # model.py:38-39


class _Params(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("first-stage modules are parameter containers; decode through FirstStageModel.decode "
                           "(sm_100a DecoderEngine, no CPU path)")


class ResnetBlock(_Params):
    """model.py:82-141 with temb_channels = 0 (the autoencoders have no timestep)."""

    def __init__(self, in_channels: int, out_channels: int, conv_shortcut: bool = False):
        super().__init__()
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = Normalize(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(_Params):
    """model.py:150-202: one head over all channels."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class Upsample(_Params):
    def __init__(self, in_channels: int, with_conv: bool):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Decoder(_Params):
    """Module tree of model.py:462-533 (same attribute names => same state_dict keys)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("first stage: only the vanilla AttnBlock (the configs of the path use no other)")
        if give_pre_end or tanh_out:
            raise NotImplementedError("first stage: give_pre_end / tanh_out are not used by the configs of the path")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.z_channels, self.out_ch = resolution, in_channels, z_channels, out_ch
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class VectorQuantizer(_Params):
    """Codebook of the VQ first stages; key `quantize.embedding.weight`, init U(-1/n_e, 1/n_e) as taming's."""

    def __init__(self, n_e: int, e_dim: int):
        super().__init__()
        self.n_e, self.e_dim = n_e, e_dim
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)


class FirstStageModel(_Params):
    """Decode side of `VQModelInterface` (n_embed given) or `AutoencoderKL` (n_embed None) + the latent scaling of
    `LatentDiffusion.decode_first_stage`.  ddconfig / embed_dim / n_embed are the `params` of the reference's
    first_stage_config (e.g. models/first_stage_models/vq-f4/config.yaml, v1-inference.yaml:46-67)."""

    def __init__(self, ddconfig: dict, embed_dim: int, n_embed: Optional[int] = None, scale_factor: float = 1.0):
        super().__init__()
        self.embed_dim, self.n_embed, self.scale_factor = embed_dim, n_embed, float(scale_factor)
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        if n_embed is not None:
            self.quantize = VectorQuantizer(n_embed, embed_dim)
        self._engines: Dict[tuple, "DecoderEngine"] = {}

    def engine(self, batch: int, h: int, w: int, device=None) -> "DecoderEngine":
        dev = torch.device(device) if device is not None else self.post_quant_conv.weight.device
        key = (batch, h, w, str(dev))
        if key not in self._engines:
            self._engines[key] = DecoderEngine(self, batch, h, w, device=dev)
        return self._engines[key]

    @torch.no_grad()
    def decode(self, z: torch.Tensor, force_not_quantize: bool = False) -> torch.Tensor:
        """`first_stage_model.decode(z[, force_not_quantize])` on latents that are already un-scaled."""
        return self.engine(z.shape[0], z.shape[2], z.shape[3], z.device).decode(z, force_not_quantize, scale=False)

    @torch.no_grad()
    def decode_first_stage(self, z: torch.Tensor, predict_cids: bool = False, force_not_quantize: bool = False):
        """`LatentDiffusion.decode_first_stage` (ddpm.py:706-764): z / scale_factor, then decode."""
        if predict_cids:
            raise NotImplementedError("decode_first_stage(predict_cids=True) is not on the sampling path")
        return self.engine(z.shape[0], z.shape[2], z.shape[3], z.device).decode(z, force_not_quantize, scale=True)


# ------------------------------------------------------------------------------------------ the kernel program
class DecoderEngine(StepEngine):
    """Static kernel program of one decode at a fixed (batch, h, w) latent shape; shares the op-emission helpers of the
    step engine (`_gn`, `_fp_input`, `_plain_conv`, `_attention`), all layers floating point (fp16 hi/lo split x3)."""

    def __init__(self, fs: FirstStageModel, batch: int, h: int, w: int, device=None, use_graph: bool = True,
                 fuse_gn: bool = True):
        dev = torch.device(device) if device is not None else fs.post_quant_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("DecoderEngine needs the first stage on an sm_100a GPU (no CPU path)")
        self.dev, self.batch, self.use_graph, self.fuse_gn = dev, batch, use_graph, fuse_gn
        self.fp_mode, self.fp_passes = "h16", 3
        self.fs = fs
        self.ops: List = []
        self.tensors: List[T] = []
        self._gn_slots = 0
        self._h16: Dict = {}
        self._consts, self._plain, self._keep = {}, {}, []
        self.teacher = None
        self.block_out: Dict[str, T] = {}
        self._names = {id(m): n for n, m in fs.named_modules()}
        self.graph = None
        dec = fs.decoder
        zc = dec.z_channels
        with torch.cuda.device(dev):
            self.z_in = torch.zeros((batch, fs.embed_dim, h, w), dtype=torch.float32, device=dev)
            self.z_post = torch.zeros((batch, zc, h, w), dtype=torch.float32, device=dev)
            self.indices = torch.zeros((batch * h * w,), dtype=torch.int32, device=dev) if fs.n_embed else None
            self._trace(dec, batch, h, w)
            for t in self.tensors:
                r, _ = t.root()
                if r.buf is None:
                    r.buf = torch.zeros((r.n, r.h, r.w, r.c), dtype=torch.float32, device=dev)
            self.gn_ws = torch.zeros((max(self._gn_slots, 1), batch, 32, 2), dtype=torch.float64, device=dev)
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}

    # ---------------------------------------------------------------- program
    def _conv(self, conv: nn.Module, src, out: T, res: Optional[T] = None):
        self._plain_conv(conv, src, out, res, pad_lo=conv.kernel_size[0] // 2, stride=1)
        return out

    def _trace(self, dec: Decoder, N: int, h: int, w: int):
        fs = self.fs

        def resblock(blk: ResnetBlock, x: T) -> T:
            hh = self._new(x.n, x.h, x.w, blk.out_channels)
            self._conv(blk.conv1, self._fp_input(x, self._gn(x, blk.norm1), silu=True), hh)
            out = self._new(x.n, x.h, x.w, blk.out_channels)
            r = x
            if blk.in_channels != blk.out_channels:
                self._conv(blk.conv_shortcut if blk.use_conv_shortcut else blk.nin_shortcut, x, out)
                r = out
            self._conv(blk.conv2, self._fp_input(hh, self._gn(hh, blk.norm2), silu=True), out, res=r)
            self.block_out[self._names[id(blk)]] = out
            return out

        def attnblock(blk: AttnBlock, x: T) -> T:
            c = x.c
            # q, k, v: three 1x1 convs over the same normalised input -> ONE conv with the weights stacked
            qkv_conv = SimpleNamespace(
                weight=torch.cat([blk.q.weight, blk.k.weight, blk.v.weight], 0).detach(),
                bias=torch.cat([blk.q.bias, blk.k.bias, blk.v.bias], 0).detach())
            self._keep.append(qkv_conv)
            qkv = self._new(x.n, x.h, x.w, 3 * c)
            self._plain_conv(qkv_conv, self._fp_input(x, self._gn(x, blk.norm)), qkv, None, pad_lo=0, stride=1)
            planes = self._attn_out_planes(x.n, x.h, x.w, c, c)
            o = T(x.n, x.h, x.w, c) if planes is not None else self._new(x.n, x.h, x.w, c)

            def strides():
                s = (qkv.view.stride(0), 0, qkv.view.stride(2))
                return dict(q=s, k=s, v=s, o=(x.h * x.w * c, 0, c))

            def part(i):
                return lambda: qkv.view.reshape(-1)[i * c:]
            self._attention(part(0), part(1), part(2), o, 1, c, float(int(c) ** -0.5), strides, o_h16=planes)
            out = self._new(x.n, x.h, x.w, c)
            self._conv(blk.proj_out, planes if planes is not None else o, out, res=x)
            self.block_out[self._names[id(blk)]] = out
            return out

        x0 = self._new(N, h, w, dec.conv_in.out_channels)
        w_in = self._const(dec.conv_in.weight)
        b_in = self._const(dec.conv_in.bias)
        self.ops.append(lambda: ops.conv_in(self.z_post, w_in, b_in, x0.view))
        self.block_out["decoder.conv_in"] = x0
        x = resblock(dec.mid.block_1, x0)
        x = attnblock(dec.mid.attn_1, x)
        x = resblock(dec.mid.block_2, x)
        for lvl in reversed(range(dec.num_resolutions)):
            st = dec.up[lvl]
            for j in range(dec.num_res_blocks + 1):
                x = resblock(st.block[j], x)
                if len(st.attn) > 0:
                    x = attnblock(st.attn[j], x)
            if lvl != 0:
                if st.upsample.with_conv:
                    up = self._new(x.n, 2 * x.h, 2 * x.w, x.c)
                    self._conv(st.upsample.conv, self._fp_input(x, None, False, dict(upsample=True)), up)
                else:
                    up = self._new(x.n, 2 * x.h, 2 * x.w, x.c)
                    self.ops.append(lambda x=x, up=up: ops.act_prepare(x.view, dst_f32=up.view, upsample=True))
                self.block_out[self._names[id(st.upsample)]] = up
                x = up
        gn = self._gn(x, dec.norm_out)
        f = self._new(x.n, x.h, x.w, x.c)
        self.ops.append(lambda: ops.act_prepare(x.view, dst_f32=f.view, silu=True, **self._gn_args(gn)))
        w_out = self._const(dec.conv_out.weight)
        b_out = self._const(dec.conv_out.bias)
        self.image = torch.zeros((N, dec.out_ch, x.h, x.w), dtype=torch.float32, device=self.dev)
        self.ops.append(lambda: ops.conv_out(f.view, w_out, b_out, self.image))

    # ---------------------------------------------------------------- execution
    def _not_a_step_engine(self, *a, **k):
        raise RuntimeError("DecoderEngine only decodes latents (decode); it shares StepEngine's op-emission helpers, not its "
                           "sampling interface")
    forward = forward_teacher_forced = step = sample = set_schedule = select_step = set_guidance = _not_a_step_engine

    def _run(self, quantize: bool, scale: bool):
        fs = self.fs
        inv = float(torch.tensor(1.0 / fs.scale_factor, dtype=torch.float32)) if scale else 1.0
        cb = self._const(fs.quantize.embedding.weight) if quantize else None
        ops.first_stage_input(self.z_in, inv, self.z_post, codebook=cb, w=self._const(fs.post_quant_conv.weight),
                              bias=self._const(fs.post_quant_conv.bias), indices=self.indices if quantize else None)
        ops.fill_zero(self.gn_ws)
        for op in self.ops:
            op()

    @property
    def launches_per_decode(self) -> int:
        from . import _lib
        ctx = _lib.context(self.dev.index or 0)
        before = ctx.launches
        self._run(self.fs.n_embed is not None, True)
        torch.cuda.synchronize(self.dev)
        return ctx.launches - before

    @torch.no_grad()
    def decode(self, z: torch.Tensor, force_not_quantize: bool = False, scale: bool = True) -> torch.Tensor:
        """z: [batch, embed_dim, h, w] fp32 on the engine's device -> images [batch, out_ch, H, W] fp32."""
        if not z.is_cuda or tuple(z.shape) != tuple(self.z_in.shape):
            raise RuntimeError(f"DecoderEngine.decode: expected a CUDA latent of shape {tuple(self.z_in.shape)}, got "
                               f"{tuple(z.shape)} on {z.device} (no CPU path)")
        quantize = self.fs.n_embed is not None and not force_not_quantize
        self.z_in.copy_(z)
        if not self.use_graph:
            self._run(quantize, scale)
            return self.image.clone()
        # 1 / scale_factor is a kernel argument baked into the captured graph, and `decode_first_stage` callers may assign
        # `fs.scale_factor` between calls: its value is part of the key
        key = (quantize, scale, float(self.fs.scale_factor) if scale else None)
        g = self._graphs.get(key)
        if g is None:
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                self._run(quantize, scale)       # warm-up: module loading, shared-memory attributes
            torch.cuda.current_stream(self.dev).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(quantize, scale)
            self._graphs[key] = g
        g.replay()
        return self.image.clone()
