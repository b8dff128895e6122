"""Tensor-level wrappers over the C ABI (torch is only the owner of device memory here).

Every function launches on torch's current CUDA stream and is graph-capturable.
Activations are NHWC views: a 4-D tensor [N, H, W, C] whose last dim is contiguous and whose
pixel pitch (`stride(2)`) may exceed C (a window into a concat buffer).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ActDesc, AttnDesc, AttnH16Desc, ConvFpDesc, ConvH16Desc, ConvW4A8Desc, GnTarget, LinearDesc


def _ctx(t: torch.Tensor) -> _lib.Context:
    if not t.is_cuda:
        raise RuntimeError("tfmq_b200 kernels need CUDA tensors on an sm_100 device; there is no CPU path")
    return _lib.context(t.device.index or 0)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _nhwc(t: torch.Tensor):
    """(N, H, W, C, ld) of an NHWC view; checks the layout the kernels assume."""
    assert t.dim() == 4 and t.stride(3) == 1, "expected NHWC view with contiguous channels"
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "NHWC view must be pixel-contiguous"
    return n, h, w, c, ld


# --------------------------------------------------------------------------- weights
def pack_w4(w2d: torch.Tensor, delta: torch.Tensor, zp: torch.Tensor, alpha: torch.Tensor | None = None):
    """w2d: [cout, k] fp32 in (tap, cin) order. Returns (codes u8 [cout,k], packed u8 [cout,k/2], wsum i32 [cout])."""
    ctx = _ctx(w2d)
    cout, k = w2d.shape
    w2d = w2d.contiguous().float()
    delta = delta.reshape(-1).contiguous().float()
    zp = zp.reshape(-1).contiguous().float()
    assert delta.numel() == cout and zp.numel() == cout
    if alpha is not None:
        alpha = alpha.reshape(cout, k).contiguous().float()
    codes = torch.empty((cout, k), dtype=torch.uint8, device=w2d.device)
    packed = torch.empty((cout, k // 2), dtype=torch.uint8, device=w2d.device)
    wsum = torch.empty((cout,), dtype=torch.int32, device=w2d.device)
    ctx.call("tfmq_pack_w4", _p(w2d), _p(delta), _p(zp), _p(alpha), cout, k, _p(codes), _p(packed), _p(wsum), _stream())
    return codes, packed, wsum


# --------------------------------------------------------------------------- GroupNorm + producer
def gn_stats(x: torch.Tensor, groups: int, stats: torch.Tensor | None = None) -> torch.Tensor:
    ctx = _ctx(x)
    n, h, w, c, ld = _nhwc(x)
    if stats is None:
        stats = torch.zeros((n, groups, 2), dtype=torch.float64, device=x.device)
    ctx.call("tfmq_gn_stats", _p(x), ld, n, h * w, c, groups, _p(stats), _stream())
    return stats


def _set_stats(d, stats):
    """stats: list of (stats tensor [n, groups, 2] f64, cpg, ch_off) -- GroupNorm targets fed by a conv output."""
    if not stats:
        return
    assert len(stats) <= 2
    d.n_stat = len(stats)
    for i, (t, cpg, ch_off) in enumerate(stats):
        d.stat[i].stats, d.stat[i].cpg, d.stat[i].ch_off, d.stat[i].groups = t.data_ptr(), cpg, ch_off, t.shape[-2]


def gn_stats_part(x: torch.Tensor, stats: torch.Tensor, cpg: int, ch_off: int):
    ctx = _ctx(x)
    n, h, w, c, ld = _nhwc(x)
    t = GnTarget()
    t.stats, t.cpg, t.ch_off, t.groups = stats.data_ptr(), cpg, ch_off, stats.shape[-2]
    ctx.call("tfmq_gn_stats_part", _p(x), ld, n, h * w, c, C.byref(t), _stream())


def fill_zero(t: torch.Tensor):
    _ctx(t).call("tfmq_fill_zero", _p(t), t.numel() * t.element_size(), _stream())


def act_prepare(src: torch.Tensor, *, aq: torch.Tensor | None = None, dst_u8: torch.Tensor | None = None,
                halo: int = 0, dst_c_off: int = 0, dst_f32: torch.Tensor | None = None,
                gn_stats_t: torch.Tensor | None = None, gamma=None, beta=None, groups: int = 32,
                eps: float = 1e-5, silu: bool = False, upsample: bool = False, dst_h16=None,
                ln=None, geglu: bool = False):
    """[GN] -> [SiLU] -> u8 codes into a halo-padded NHWC buffer, or fp32 NHWC, or (dst_h16 = (hi, lo)) the fp16
    hi / lo planes `conv_h16` reads."""
    ctx = _ctx(src)
    n, h, w, c, ld = _nhwc(src)
    d = ActDesc()
    d.src, d.src_ld = src.data_ptr(), ld
    if geglu:                    # src rows are [value | gate]
        assert c % 2 == 0
        c //= 2
        d.geglu = 1
    if ln is not None:           # (gamma, beta, eps): LayerNorm over the channels of every token
        d.ln_gamma, d.ln_beta, d.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), float(ln[2])
    d.n, d.h, d.w, d.c = n, h, w, c
    d.upsample = int(upsample)
    d.gn_stats = gn_stats_t.data_ptr() if gn_stats_t is not None else None
    d.gamma = gamma.data_ptr() if gamma is not None else None
    d.beta = beta.data_ptr() if beta is not None else None
    d.groups, d.eps, d.silu = groups, eps, int(silu)
    oh, ow = (2 * h, 2 * w) if upsample else (h, w)
    if dst_u8 is not None:
        assert dst_u8.dtype == torch.uint8 and dst_u8.is_contiguous()
        assert tuple(dst_u8.shape[:3]) == (n, oh + 2 * halo, ow + 2 * halo), (dst_u8.shape, (n, oh, ow, halo))
        d.aq = aq.data_ptr()
        d.dst_u8 = dst_u8.data_ptr()
        d.halo, d.dst_c, d.dst_c_off = halo, dst_u8.shape[3], dst_c_off
    elif dst_h16 is not None:
        hi, lo = dst_h16
        assert hi.dtype == torch.float16 and lo.dtype == torch.float16 and hi.shape == lo.shape
        dn, dh, dw, dc, dld = _nhwc(hi)
        assert (dn, dh, dw, dc) == (n, oh, ow, c) and _nhwc(lo)[4] == dld
        d.dst_hi, d.dst_lo, d.dst_h_ld = hi.data_ptr(), lo.data_ptr(), dld
    else:
        dn, dh, dw, dc, dld = _nhwc(dst_f32)
        assert (dn, dh, dw, dc) == (n, oh, ow, c)
        d.dst_f32, d.dst_ld = dst_f32.data_ptr(), dld
    ctx.call("tfmq_act_prepare", C.byref(d), _stream())


# --------------------------------------------------------------------------- tensor-core convs
def conv_w4a8(act: torch.Tensor, ksize: int, packed, wzp_i32, wdelta, wsum, bias, aq, out: torch.Tensor,
              emb: torch.Tensor | None = None, res: torch.Tensor | None = None, stats=None):
    """act: u8 [N, H+2h, W+2h, Cin] (h = 1 for 3x3, 0 for 1x1); out: fp32 NHWC view [N,H,W,Cout]."""
    ctx = _ctx(out)
    n, h, w, cout, out_ld = _nhwc(out)
    halo = 1 if ksize == 3 else 0
    assert act.dtype == torch.uint8 and act.is_contiguous()
    assert tuple(act.shape[:3]) == (n, h + 2 * halo, w + 2 * halo)
    d = ConvW4A8Desc()
    d.act = act.data_ptr()
    d.n, d.h, d.w, d.cin, d.cout, d.ksize = n, h, w, act.shape[3], cout, ksize
    assert wzp_i32.dtype == torch.int32, "weight zero points are int32 (ABI 6)"
    d.packed, d.wzp, d.wdelta, d.wsum = packed.data_ptr(), wzp_i32.data_ptr(), wdelta.data_ptr(), wsum.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.aq = aq.data_ptr()
    if emb is not None:
        assert emb.stride(-1) == 1
        d.emb, d.emb_ld = emb.data_ptr(), (emb.stride(0) if emb.dim() == 2 and emb.shape[0] > 1 else 0)
    if res is not None:
        rn, rh, rw, rc, rld = _nhwc(res)
        assert (rn, rh, rw, rc) == (n, h, w, cout)
        d.res, d.res_ld = res.data_ptr(), rld
    d.out, d.out_ld = out.data_ptr(), out_ld
    _set_stats(d, stats)
    ctx.call("tfmq_conv_w4a8", C.byref(d), _stream())


def conv_fp(x: torch.Tensor, ksize: int, stride: int, pad_lo: int, w_hi, w_lo, out: torch.Tensor, bias=None,
            wscale=None, res=None, passes: int = 3, emb=None, stats=None):
    ctx = _ctx(out)
    n, h, w, cin, x_ld = _nhwc(x)
    on, oh, ow, cout, out_ld = _nhwc(out)
    assert on == n
    d = ConvFpDesc()
    d.x, d.x_ld = x.data_ptr(), x_ld
    d.n, d.h, d.w, d.cin, d.cout = n, h, w, cin, cout
    d.ksize, d.stride, d.pad_lo, d.out_h, d.out_w = ksize, stride, pad_lo, oh, ow
    d.w_hi = w_hi.data_ptr()
    d.w_lo = w_lo.data_ptr() if w_lo is not None else None
    d.wscale = wscale.data_ptr() if wscale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    if res is not None:
        rn, rh, rw, rc, rld = _nhwc(res)
        assert (rn, rh, rw, rc) == (n, oh, ow, cout)
        d.res, d.res_ld = res.data_ptr(), rld
    d.out, d.out_ld = out.data_ptr(), out_ld
    d.passes = passes
    if emb is not None:
        assert emb.stride(-1) == 1
        d.emb, d.emb_ld = emb.data_ptr(), (emb.stride(0) if emb.dim() == 2 and emb.shape[0] > 1 else 0)
    _set_stats(d, stats)
    ctx.call("tfmq_conv_fp", C.byref(d), _stream())


def conv_h16(x_hi: torch.Tensor, x_lo: torch.Tensor, ksize: int, stride: int, pad_lo: int, w_hi, w_lo, out,
             bias=None, wscale=None, res=None, emb=None, stats=None, out_h16=None, ksplit: int = 0):
    """fp32-accurate conv on kind::f16 from pre-split fp16 hi / lo planes (see include/tfmq_b200.h).
    out_h16 = (hi, lo): the result is written as fp16 hi / lo planes instead of fp32 `out` (which may be None).
    ksplit: split-K for pixel-axis reductions (weight gradients): > 1 K ranges per output tile, < 0 chosen by the library."""
    ctx = _ctx(x_hi)
    assert x_hi.dtype == torch.float16 and x_lo.dtype == torch.float16 and w_hi.dtype == torch.float16
    n, h, w, cin, x_ld = _nhwc(x_hi)
    assert _nhwc(x_lo) == (n, h, w, cin, x_ld)
    d = ConvH16Desc()
    if out_h16 is not None:
        assert res is None and emb is None and not stats
        ph, pl = out_h16
        assert ph.dtype == torch.float16 and pl.dtype == torch.float16
        on, oh, ow, cout, h_ld = _nhwc(ph)
        assert _nhwc(pl) == (on, oh, ow, cout, h_ld)
        d.out_hi, d.out_lo, d.out_h_ld = ph.data_ptr(), pl.data_ptr(), h_ld
        out_ld = 0
    else:
        on, oh, ow, cout, out_ld = _nhwc(out)
    assert on == n
    d.x_hi, d.x_lo, d.x_ld = x_hi.data_ptr(), x_lo.data_ptr(), x_ld
    d.n, d.h, d.w, d.cin, d.cout = n, h, w, cin, cout
    d.ksize, d.stride, d.pad_lo, d.out_h, d.out_w = ksize, stride, pad_lo, oh, ow
    d.w_hi = w_hi.data_ptr()
    d.w_lo = w_lo.data_ptr() if w_lo is not None else None
    d.wscale = wscale.data_ptr() if wscale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    if res is not None:
        rn, rh, rw, rc, rld = _nhwc(res)
        assert (rn, rh, rw, rc) == (n, oh, ow, cout)
        d.res, d.res_ld = res.data_ptr(), rld
    d.out, d.out_ld = (out.data_ptr() if out_h16 is None else None), out_ld
    if emb is not None:
        assert emb.stride(-1) == 1
        d.emb, d.emb_ld = emb.data_ptr(), (emb.stride(0) if emb.dim() == 2 and emb.shape[0] > 1 else 0)
    _set_stats(d, stats)
    d.ksplit = ksplit
    ctx.call("tfmq_conv_h16", C.byref(d), _stream())


def split_h16(w2d: torch.Tensor, wscale: torch.Tensor | None = None, keep_lo: bool = False):
    """Load-time split of fp32 weights [cout][k] into fp16 planes: w * 2^e = hi + lo per output channel, with e chosen so
    that max|w| lands in [2^12, 2^13) (lo stays a normal fp16).  Returns (hi, lo or None, scale) where scale[c] = 2^-e
    (times `wscale` if given) is what the conv epilogue multiplies by; lo is None when every weight is exact in fp16."""
    w2d = w2d.contiguous().float()
    amax = w2d.abs().amax(dim=1).clamp_min(1e-30)
    e = 12 - torch.floor(torch.log2(amax))
    e = torch.where(amax * torch.exp2(e) >= 8192.0, e - 1, e)      # guard the log2 rounding at powers of two
    sc = torch.exp2(e)
    ws = w2d * sc[:, None]
    hi = ws.half()
    lo = (ws - hi.float()).half()
    inv = torch.exp2(-e)
    if wscale is not None:
        inv = inv * wscale.float()
    # keep_lo: always return the lo plane (no host sync on `any()`): the reconstruction loop re-splits its weights every iteration
    return hi.contiguous(), (lo.contiguous() if (keep_lo or bool((lo != 0).any())) else None), inv.contiguous()


def split_tf32(w2d: torch.Tensor):
    """w = hi + lo with hi exactly representable in tf32 (low 13 mantissa bits zero)."""
    w2d = w2d.contiguous().float()
    hi = (w2d.view(torch.int32) & -8192).view(torch.float32)
    lo = w2d - hi
    return hi.contiguous(), lo.contiguous()


def gemm_i8_peak(a_u8: torch.Tensor, b_s8: torch.Tensor, out_i32: torch.Tensor):
    ctx = _ctx(out_i32)
    m, k = a_u8.shape
    n = b_s8.shape[0]
    ctx.call("tfmq_gemm_i8_peak", _p(a_u8), _p(b_s8), m, n, k, _p(out_i32), _stream())


# --------------------------------------------------------------------------- small kernels
def conv_in(x_nchw: torch.Tensor, w, bias, out: torch.Tensor):
    ctx = _ctx(out)
    n, cin, h, wd = x_nchw.shape
    on, oh, ow, cout, ld = _nhwc(out)
    assert x_nchw.is_contiguous() and (on, oh, ow) == (n, h, wd)
    ctx.call("tfmq_conv_in", _p(x_nchw), _p(w), _p(bias), n, h, wd, cin, cout, _p(out), ld, _stream())


def conv_out(x: torch.Tensor, w, bias, out_nchw: torch.Tensor):
    ctx = _ctx(x)
    n, h, wd, cin, ld = _nhwc(x)
    cout = out_nchw.shape[1]
    assert out_nchw.is_contiguous()
    ctx.call("tfmq_conv_out", _p(x), ld, _p(w), _p(bias), n, h, wd, cin, cout, _p(out_nchw), _stream())


def _linear_desc(x, out, w_f32, codes, wzp_f, wdelta, bias, aq, silu_in) -> LinearDesc:
    m, in_f = x.shape
    d = LinearDesc()
    d.x, d.x_ld = x.data_ptr(), x.stride(0)
    d.m, d.in_f, d.out_f = m, in_f, out.shape[1]
    d.silu_in = int(silu_in)
    d.aq = aq.data_ptr() if aq is not None else None
    d.w_f32 = w_f32.data_ptr() if w_f32 is not None else None
    d.codes = codes.data_ptr() if codes is not None else None
    d.wzp_f = wzp_f.data_ptr() if wzp_f is not None else None
    d.wdelta = wdelta.data_ptr() if wdelta is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.out, d.out_ld = out.data_ptr(), out.stride(0)
    return d


def linear_small(x: torch.Tensor, out: torch.Tensor, *, w_f32=None, codes=None, wzp_f=None, wdelta=None, bias=None,
                 aq=None, silu_in: bool = False):
    ctx = _ctx(out)
    d = _linear_desc(x, out, w_f32, codes, wzp_f, wdelta, bias, aq, silu_in)
    ctx.call("tfmq_linear_small", C.byref(d), _stream())


class LinearGroup:
    """Several small-M linears replayed as one launch (tfmq_linear_grouped).  `layers` is a list of keyword
    dicts with the arguments of `linear_small` (x, out, ...); every tensor must stay alive with this object."""

    def __init__(self, layers):
        assert len(layers) >= 1
        self.layers = layers
        out0 = layers[0]["out"]
        self.ctx = _ctx(out0)
        n = len(layers)
        descs = (LinearDesc * n)()
        for i, kw in enumerate(layers):
            descs[i] = _linear_desc(kw["x"], kw["out"], kw.get("w_f32"), kw.get("codes"), kw.get("wzp_f"),
                                    kw.get("wdelta"), kw.get("bias"), kw.get("aq"), kw.get("silu_in", False))
        starts = (C.c_int * (n + 1))()
        self.ctx.call("tfmq_linear_grouped_plan", descs, n, starts)
        raw = bytes(descs)
        self.descs_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(out0.device)
        self.starts_dev = torch.tensor(list(starts), dtype=torch.int32, device=out0.device)
        self.n, self.total = n, int(starts[n])
        self.m = layers[0]["x"].shape[0]
        self.max_in = max(kw["x"].shape[1] for kw in layers)

    def run(self):
        self.ctx.call("tfmq_linear_grouped", _p(self.descs_dev), _p(self.starts_dev), self.n, self.total, self.m,
                      self.max_in, _stream())


def timestep_embedding(t: torch.Tensor, dim: int, style: int, out: torch.Tensor):
    _ctx(out).call("tfmq_timestep_embedding", _p(t), t.numel(), dim, style, _p(out), _stream())


ATTN_PLANE_DIMS = (32, 40, 64, 80, 160, 256, 384, 512, 576, 960)   # head dims whose kernels can write fp16 hi / lo planes


def attention(q, k, v, o, b: int, heads: int, tq: int, tk: int, d: int, scale: float, strides, o_h16=None):
    """strides: dict name -> (sb, sh, st) in elements for q, k, v, o.  o_h16 = (hi, lo): the output goes to these fp16
    planes (o's strides) instead of fp32 `o`, which may then be None."""
    ctx = _ctx(q)
    a = AttnDesc()
    if o_h16 is not None:
        assert o_h16[0].dtype == torch.float16 and o_h16[1].dtype == torch.float16
        a.o_hi, a.o_lo = o_h16[0].data_ptr(), o_h16[1].data_ptr()
    if o is None and o_h16 is None:
        raise ValueError("attention: give the fp32 output `o` or the fp16 planes `o_h16`")
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o)):
        sb, sh, st = strides[name]
        setattr(a, name, t.data_ptr() if t is not None else None)
        setattr(a, name + "_sb", sb)
        setattr(a, name + "_sh", sh)
        setattr(a, name + "_st", st)
    a.b, a.heads, a.tq, a.tk, a.d, a.scale = b, heads, tq, tk, d, scale
    ctx.call("tfmq_attention", C.byref(a), _stream())


ATTN_TC_DIMS = tuple(range(16, 65, 8))    # head dims of the tcgen05 kernel (tfmq_attention_h16)


def attention_h16(q, k, v, o, b: int, heads: int, tq: int, tk: int, d: int, scale: float, strides, o_h16=None):
    """The attention core on tcgen05 / TMEM from pre-split operands: q, k, v = (hi, lo) fp16 plane pairs; strides as in
    `attention` (in halves for q / k / v; in elements of the output for o).  Output: fp32 `o` or the planes `o_h16`."""
    ctx = _ctx(q[0])
    a = AttnH16Desc()
    for name, t in (("q", q), ("k", k), ("v", v)):
        assert t[0].dtype == torch.float16 and t[1].dtype == torch.float16
        sb, sh, st = strides[name]
        setattr(a, name + "_hi", t[0].data_ptr())
        setattr(a, name + "_lo", t[1].data_ptr())
        setattr(a, name + "_sb", sb)
        setattr(a, name + "_sh", sh)
        setattr(a, name + "_st", st)
    if o_h16 is not None:
        assert o_h16[0].dtype == torch.float16 and o_h16[1].dtype == torch.float16
        a.o_hi, a.o_lo = o_h16[0].data_ptr(), o_h16[1].data_ptr()
    elif o is None:
        raise ValueError("attention_h16: give the fp32 output `o` or the fp16 planes `o_h16`")
    else:
        a.o = o.data_ptr()
    a.o_sb, a.o_sh, a.o_st = strides["o"]
    a.b, a.heads, a.tq, a.tk, a.d, a.scale = b, heads, tq, tk, d, scale
    ctx.call("tfmq_attention_h16", C.byref(a), _stream())


def ddim_update(x, e, coef, x_prev, x0_out=None, noise=None):
    _ctx(x).call("tfmq_ddim_update", _p(x), _p(e), _p(noise), _p(coef), x.numel(), _p(x_prev), _p(x0_out), _stream())


def cfg_combine(e_u, e_c, s: float, out):
    _ctx(out).call("tfmq_cfg_combine", _p(e_u), _p(e_c), C.c_float(s), out.numel(), _p(out), _stream())


def plms_eps(e0, olds, out):
    """Adams-Bashforth combination of e0 with the stored predictions `olds` = [old_eps[-1], old_eps[-2], ...] (up to 3);
    `olds` = ["euler", e_next] selects the (e0 + e_next) / 2 of the first step."""
    ctx = _ctx(out)
    if len(olds) == 2 and isinstance(olds[0], str):
        order, es = 1, [olds[1], None, None]
    else:
        order = 0 if not olds else min(len(olds), 3) + 1
        es = (list(olds) + [None, None, None])[:3]
    ctx.call("tfmq_plms_eps", _p(e0), _p(es[0]), _p(es[1]), _p(es[2]), order, out.numel(), _p(out), _stream())


def first_stage_input(z: torch.Tensor, inv_scale: float, out: torch.Tensor, codebook=None, w=None, bias=None,
                      indices=None):
    """Latent -> decoder input: z / scale_factor, [nearest-codebook lookup], [post_quant_conv 1x1]; NCHW fp32 in and out."""
    ctx = _ctx(out)
    n, c, h, wd = z.shape
    assert z.is_contiguous() and out.is_contiguous() and out.shape[0] == n and tuple(out.shape[2:]) == (h, wd)
    if codebook is not None:
        assert codebook.is_contiguous() and codebook.shape[1] == c
    if w is not None:
        w = w.reshape(out.shape[1], c)
        assert w.is_contiguous()
    ctx.call("tfmq_first_stage_input", _p(z), C.c_float(inv_scale), _p(codebook),
             codebook.shape[0] if codebook is not None else 0, _p(w), _p(bias), n, h * wd, c, out.shape[1], _p(out),
             _p(indices), _stream())


# --------------------------------------------------------------------------- calibration
def minmax_rows(x2d: torch.Tensor) -> torch.Tensor:
    rows, cols = x2d.shape
    mm = torch.empty((rows, 2), dtype=torch.float32, device=x2d.device)
    _ctx(x2d).call("tfmq_minmax_rows", _p(x2d), rows, cols, _p(mm), _stream())
    return mm


def mse_scale_search(x2d: torch.Tensor, level: int):
    rows, cols = x2d.shape
    delta = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    zp = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    _ctx(x2d).call("tfmq_mse_scale_search", _p(x2d), rows, cols, level, _p(delta), _p(zp), _stream())
    return delta, zp


def act_range_update(x: torch.Tensor, state: torch.Tensor, aq: torch.Tensor, momentum: float = 0.95,
                     level: int = 256):
    x2 = x.reshape(-1, x.shape[-1]) if x.stride(-1) == 1 else x.contiguous().reshape(-1, x.shape[-1])
    assert x2.stride(1) == 1
    _ctx(x).call("tfmq_act_range_update", _p(x2), x2.stride(0), x2.shape[0], x2.shape[1], C.c_float(momentum), level,
                 _p(state), _p(aq), _stream())


def adaround_soft(w2d, delta, zp, alpha, level: int, out):
    cout, k = w2d.shape
    _ctx(w2d).call("tfmq_adaround_soft", _p(w2d), _p(delta), _p(zp), _p(alpha), cout, k, level, _p(out), _stream())


def adaround_step(w2d, delta, zp, alpha, grad_w, adam_m, adam_v, level: int, step: int, lr: float, b: float,
                  lam: float, round_loss):
    cout, k = w2d.shape
    _ctx(w2d).call("tfmq_adaround_step", _p(w2d), _p(delta), _p(zp), _p(alpha), _p(grad_w), _p(adam_m), _p(adam_v),
                   cout, k, level, step, C.c_float(lr), C.c_float(b), C.c_float(lam), _p(round_loss), _stream())


def rec_loss(pred, tgt, batch: int, loss, grad=None):
    _ctx(pred).call("tfmq_rec_loss", _p(pred), _p(tgt), pred.numel(), batch, _p(loss), _p(grad), _stream())
