"""Forward / backward of a QuantLayer's convolution or linear map on this library's tcgen05 kernels, for the AdaRound /
TIAR reconstruction loop (reference quant/reconstruction.py:182-198: `out_quant = block(*cur_inputs)`, `err.backward()`).

The unit's graph is still walked by torch autograd (GroupNorm, SiLU, softmax and the adds are its elementwise glue), but the
three contractions of every reconstructed layer run here, fp32-accurate (fp16 hi / lo split, three products, fp32
accumulation in tensor memory -- the same `tfmq_conv_h16` kernel the sampling path uses for its floating-point layers):

  forward   y  = conv(x, W_soft) + b                        implicit GEMM over NHWC planes
  dgrad     dx = conv(dy, flip(W_soft)^T)                   the same kernel on the flipped / transposed weights
  wgrad     dW[co, ci, ky, kx] = sum_p dy[p, co] * x[p + (ky-1, kx-1), ci]
            a pixel-reduction GEMM: the kernel runs it as a 1x1 "convolution" whose pixel axis is the (kx, ci) row index and
            whose channel axis is the pixel index (K = batch * (H+2) * Wp), with dy^T as the weight operand.  No im2col: x
            and dy are laid out channel-major over the SAME zero-haloed pixel grid (row pitch Wp = a multiple of 8), so a
            tap is a constant offset along K -- the row shift (ky) is a 16-byte aligned pointer offset on the x planes, the
            column shift (kx) one of three pre-shifted copies of them.  The halo pixels of dy are zero, so the extra K
            positions add nothing.

Layout changes between torch's NCHW and the kernels' NHWC / channel-major planes are torch copies (fp16 for the planes).
Shapes the kernel does not take (stride 2, groups, channels not a multiple of 16, non power-of-two maps) return None from
`tc_conv` and the caller keeps torch's op for that layer.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops


def _pow2(v: int) -> bool:
    return v > 0 and (v & (v - 1)) == 0


def supported(x: torch.Tensor, w: torch.Tensor, fwd_kwargs: dict) -> bool:
    if not (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32):
        return False
    if w.dim() == 2:
        return x.dim() >= 2 and w.shape[1] % 16 == 0 and w.shape[0] % 16 == 0
    if w.dim() != 4 or x.dim() != 4:
        return False
    k = w.shape[2]
    st, pad, dil, grp = (fwd_kwargs.get(n, d) for n, d in (("stride", (1, 1)), ("padding", (0, 0)), ("dilation", (1, 1)),
                                                             ("groups", 1)))
    return (k in (1, 3) and w.shape[3] == k and tuple(st) == (1, 1) and tuple(pad) == (k // 2, k // 2) and tuple(dil) == (1, 1)
            and grp == 1 and w.shape[1] % 16 == 0 and w.shape[0] % 16 == 0 and _pow2(x.shape[2]) and _pow2(x.shape[3]))


def _absmax(t: torch.Tensor) -> torch.Tensor:
    """max |t| as a device scalar in ONE pass over t (`t.abs().amax()` writes |t| out and reads it back: three passes)."""
    mn, mx = torch.aminmax(t)
    return torch.maximum(-mn, mx)


def _planes(t_nhwc: torch.Tensor, autoscale: bool = False):
    """fp16 hi / lo planes of an NHWC tensor.  autoscale (gradients): the tensor is first multiplied by the power of two that
    brings its largest magnitude to [2^10, 2^11) -- back-propagated errors of 1e-6 would sit in fp16's subnormal range --
    and the inverse is returned as a device scalar for the caller to fold into the kernel's per-channel output scale."""
    inv = None
    if autoscale:
        amax = _absmax(t_nhwc).clamp_min(1e-30)
        s = torch.exp2(10.0 - torch.floor(torch.log2(amax)))
        t_nhwc = t_nhwc * s
        inv = 1.0 / s
    hi = torch.empty(t_nhwc.shape, dtype=torch.float16, device=t_nhwc.device)
    lo = torch.empty_like(hi)
    ops.act_prepare(t_nhwc, dst_h16=(hi, lo))
    return (hi, lo, inv) if autoscale else (hi, lo)


def _gemm_nt(a_rows: torch.Tensor, b_rows: torch.Tensor, ksplit: int = 0) -> torch.Tensor:
    """out[i, j] = sum_k a_rows[i, k] * b_rows[j, k]  (both fp32, K contiguous), on tfmq_conv_h16 as a 1x1 convolution with
    i as the pixel axis and j as the output-channel axis.  K is padded to a multiple of 16, j to a multiple of 16."""
    m, k = a_rows.shape
    n = b_rows.shape[0]
    kp, np_ = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    if kp != k:
        a_rows, b_rows = F.pad(a_rows, (0, kp - k)), F.pad(b_rows, (0, kp - k))
    if np_ != n:
        b_rows = F.pad(b_rows, (0, 0, 0, np_ - n))
    a_hi, a_lo, a_inv = _planes(a_rows.contiguous().view(m, 1, 1, kp), autoscale=True)
    w_hi, w_lo, wscale = ops.split_h16(b_rows.contiguous(), keep_lo=True)
    out = torch.empty((m, 1, 1, np_), dtype=torch.float32, device=a_rows.device)
    ops.conv_h16(a_hi, a_lo, 1, 1, 0, w_hi, w_lo, out, wscale=(wscale * a_inv).contiguous(), ksplit=ksplit)
    return out.view(m, np_)[:, :n]


_WG_BUF: dict = {}


def _wgrad_buffers(dev, ci: int, co: int, n: int, h: int, w: int):
    """Zero-haloed channel-major plane buffers for the 3x3 weight gradient, cached per shape: only the interior is ever
    written, so the halo and the guard bands stay zero between iterations."""
    key = (dev, ci, co, n, h, w)
    b = _WG_BUF.get(key)
    if b is None:
        hp, wp = h + 2, (w + 2 + 7) // 8 * 8
        q = (n * hp * wp + 15) // 16 * 16
        guard = (wp + 8 + 15) // 16 * 16
        b = dict(hp=hp, wp=wp, q=q, guard=guard,
                 x=torch.zeros((2, 3, ci, guard + q + guard), dtype=torch.float16, device=dev),
                 g=torch.zeros((2, co, q), dtype=torch.float16, device=dev))
        _WG_BUF[key] = b
    return b


def trim_buffers(keep: int = 8) -> None:
    """Drop the cached weight-gradient plane buffers once more than `keep` shapes have accumulated.  Called between units
    (reconstruction.py), never inside one: a captured iteration graph holds the addresses of the buffers it was captured with."""
    if len(_WG_BUF) > keep:
        _WG_BUF.clear()


def _flat_planes(t: torch.Tensor):
    """fp16 hi / lo planes of a contiguous fp32 tensor in ITS layout (the split is elementwise)."""
    flat = t.view(1, 1, -1, 64)
    hi = torch.empty(flat.shape, dtype=torch.float16, device=t.device)
    lo = torch.empty_like(hi)
    ops.act_prepare(flat, dst_h16=(hi, lo))
    return hi.view(t.shape), lo.view(t.shape)


def _wgrad3x3(x, gs, g_inv, ci: int, co: int):
    """dW [co, ci, 3, 3] from the layer input x [n, ci, h, w] and the scaled output gradient gs [n, co, h, w] (g_inv undoes
    the scale).  Both go channel-major onto the zero-haloed pixel grid: from NCHW that is a copy of whole image rows; the
    two column-shifted copies of x are then contiguous copies of the first."""
    n, _, h, w = x.shape
    b = _wgrad_buffers(x.device, ci, co, n, h, w)
    hp, wp, q, guard = b["hp"], b["wp"], b["q"], b["guard"]
    nq = n * hp * wp
    for pl, (xs, gp) in enumerate(zip(_flat_planes(x.contiguous()), _flat_planes(gs.contiguous()))):
        # plane s holds x shifted by kx - 1 = s - 1 columns: x[.., col] sits at padded column col + 2 - s
        b["x"][pl, 1, :, guard:guard + nq].view(ci, n, hp, wp)[:, :, 1:h + 1, 1:w + 1].copy_(xs.permute(1, 0, 2, 3))
        b["x"][pl, 0, :, guard:guard + q].copy_(b["x"][pl, 1, :, guard - 1:guard - 1 + q])
        b["x"][pl, 2, :, guard:guard + q].copy_(b["x"][pl, 1, :, guard + 1:guard + 1 + q])
        b["g"][pl, :, :nq].view(co, n, hp, wp)[:, :, 1:h + 1, 1:w + 1].copy_(gp.permute(1, 0, 2, 3))
    pitch = guard + q + guard
    rows = 3 * ci
    scale = g_inv.expand(co).contiguous()
    outs = []
    for r in range(3):                                      # ky - 1 = r - 1 rows = a pointer offset of (r - 1) * wp
        off = guard + (r - 1) * wp
        a_hi = b["x"][0].view(rows, pitch)[:, off:off + q].as_strided((rows, 1, 1, q), (pitch, pitch, pitch, 1))
        a_lo = b["x"][1].view(rows, pitch)[:, off:off + q].as_strided((rows, 1, 1, q), (pitch, pitch, pitch, 1))
        o = torch.empty((rows, 1, 1, co), dtype=torch.float32, device=x.device)
        ops.conv_h16(a_hi, a_lo, 1, 1, 0, b["g"][0], b["g"][1], o, wscale=scale, ksplit=-1)   # few tiles, K = all pixels
        outs.append(o.view(3, ci, co))                      # [kx, ci, co]
    return torch.stack(outs, 0).permute(3, 2, 0, 1).contiguous()    # [ky, kx, ci, co] -> [co, ci, ky, kx]


class _TcConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        co, k = w.shape[0], (w.shape[2] if w.dim() == 4 else 1)
        is_conv = w.dim() == 4
        if is_conv:
            x_nhwc = x.permute(0, 2, 3, 1).contiguous()
            w2d = w.permute(0, 2, 3, 1).reshape(co, -1)
        else:
            x_nhwc = x.reshape(-1, 1, 1, x.shape[-1]).contiguous()
            w2d = w
        xh, xl = _planes(x_nhwc)
        w_hi, w_lo, wscale = ops.split_h16(w2d.contiguous(), keep_lo=True)
        out = torch.empty(x_nhwc.shape[:3] + (co,), dtype=torch.float32, device=x.device)
        ops.conv_h16(xh, xl, k, 1, k // 2, w_hi, w_lo, out, bias=b.contiguous() if b is not None else None, wscale=wscale)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        ctx.is_conv = is_conv
        return out.permute(0, 3, 1, 2) if is_conv else out.view(x.shape[:-1] + (co,))

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors[:2]
        co = w.shape[0]
        gx = gw = gb = None
        if ctx.is_conv:
            k, ci = w.shape[2], w.shape[1]
            # back-propagated errors of 1e-6 would sit in fp16's subnormal range: scale by the power of two that brings the
            # largest magnitude to [2^10, 2^11); the inverse goes into the kernels' per-channel output scale
            g_s = torch.exp2(10.0 - torch.floor(torch.log2(_absmax(gy).clamp_min(1e-30))))
            g_inv = 1.0 / g_s
            gs = gy * g_s
            if ctx.needs_input_grad[0]:
                gy_nhwc = gs.permute(0, 2, 3, 1).contiguous()
                gh, gl = _planes(gy_nhwc)
                wt = w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, -1).contiguous()          # [ci][(ky', kx', co)]
                t_hi, t_lo, tscale = ops.split_h16(wt, keep_lo=True)
                gxn = torch.empty(gy_nhwc.shape[:3] + (ci,), dtype=torch.float32, device=gy.device)
                ops.conv_h16(gh, gl, k, 1, k // 2, t_hi, t_lo, gxn, wscale=(tscale * g_inv).contiguous())
                gx = gxn.permute(0, 3, 1, 2)
            if ctx.needs_input_grad[1] and k == 3:
                gw = _wgrad3x3(x, gs, g_inv, ci, co)
            elif ctx.needs_input_grad[1]:
                n, _, h, wd = x.shape
                xt = x.reshape(n, ci, h * wd).permute(1, 0, 2).reshape(ci, n * h * wd)          # rows ci, K = pixels
                gt = gy.permute(1, 0, 2, 3).reshape(co, n * h * wd)                             # rows co, K = pixels
                gw = _gemm_nt(xt, gt, ksplit=-1).t().reshape(w.shape)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = gy.sum((0, 2, 3))
        else:
            g2 = gy.reshape(-1, co)
            x2 = x.reshape(-1, x.shape[-1])
            if ctx.needs_input_grad[0]:
                gx = _gemm_nt(g2, w.t().contiguous()).view(x.shape)                            # dx[m, i] = sum_o dy[m, o] w[o, i]
            if ctx.needs_input_grad[1]:
                gw = _gemm_nt(g2.t().contiguous(), x2.t().contiguous(), ksplit=-1)              # dW[o, i] = sum_m dy[m, o] x[m, i]
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = g2.sum(0)
        return gx, gw, gb


def tc_conv(x: torch.Tensor, w: torch.Tensor, b, fwd_kwargs: dict):
    """conv2d / linear of a QuantLayer on the library's tensor-core kernels with a matching backward; None if the shape is
    one the kernels do not take."""
    if not supported(x, w, fwd_kwargs):
        return None
    return _TcConv.apply(x, w, b)
