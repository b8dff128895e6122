"""TEST INFRASTRUCTURE (CPU): torch stand-ins for the tensor-level kernel wrappers of tfmq_b200/ops.py that the first-stage
DecoderEngine program uses, with the kernels' documented semantics (include/tfmq_b200.h).  They let the HOST logic of the
program -- op order, shapes, strides, residual aliasing, GroupNorm statistics routed through conv epilogues, fp16 hi/lo planes --
be checked against the oracle without a GPU (tools/dry_trace_first_stage.py --emulate, tests/test_host_logic_cpu.py).  Never
imported by the product; nothing here is a fallback."""
import torch
import torch.nn.functional as F


def _split(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


def _add_stats(out_nhwc, stats):
    for t, cpg, ch_off in stats or []:
        c = out_nhwc.shape[-1]
        gidx = (ch_off + torch.arange(c)) // cpg
        x = out_nhwc.double()
        t[:, :, 0].index_add_(1, gidx, x.sum((1, 2)))
        t[:, :, 1].index_add_(1, gidx, (x * x).sum((1, 2)))


def fill_zero(t):
    t.zero_()


def first_stage_input(z, inv_scale, out, codebook=None, w=None, bias=None, indices=None):
    from oracle import first_stage_ref as FS
    v = z * torch.tensor(inv_scale, dtype=torch.float32)
    if codebook is not None:
        v, idx = FS.vq_lookup(v, codebook)
        if indices is not None:
            indices.copy_(idx.to(indices.dtype))
    out.copy_(F.conv2d(v, w.reshape(out.shape[1], z.shape[1], 1, 1), bias) if w is not None else v)


def conv_in(x_nchw, w, bias, out):
    out.copy_(F.conv2d(x_nchw, w, bias, padding=1).permute(0, 2, 3, 1))


def conv_out(x, w, bias, out_nchw):
    out_nchw.copy_(F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=1))


def gn_stats_part(x, stats, cpg, ch_off):
    _add_stats(x, [(stats, cpg, ch_off)])


def act_prepare(src, *, aq=None, dst_u8=None, halo=0, dst_c_off=0, dst_f32=None, gn_stats_t=None, gamma=None, beta=None,
                groups=32, eps=1e-5, silu=False, upsample=False, dst_h16=None, ln=None, geglu=False):
    assert dst_u8 is None and ln is None and not geglu, "the decoder program has no quantised / token producers"
    x = src.float()
    n, h, w, c = x.shape
    if gn_stats_t is not None:
        cnt = (c // groups) * h * w
        mean = gn_stats_t[..., 0] / cnt
        var = (gn_stats_t[..., 1] / cnt - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var.float() + eps)
        g = torch.arange(c) // (c // groups)
        a = rstd[:, g] * gamma[None, :]
        b = beta[None, :] - a * mean.float()[:, g]
        x = x * a[:, None, None, :] + b[:, None, None, :]
    if silu:
        x = x * torch.sigmoid(x)
    if upsample:
        x = x.repeat_interleave(2, 1).repeat_interleave(2, 2)
    if dst_h16 is not None:
        hi, lo = _split(x)
        dst_h16[0].copy_(hi)
        dst_h16[1].copy_(lo)
    else:
        dst_f32.copy_(x)


def conv_h16(x_hi, x_lo, ksize, stride, pad_lo, w_hi, w_lo, out, bias=None, wscale=None, res=None, emb=None, stats=None):
    assert stride == 1
    x = (x_hi.double() + x_lo.double()).permute(0, 3, 1, 2)
    w2 = w_hi.double() + (w_lo.double() if w_lo is not None else 0.0)
    cout, cin = w2.shape[0], x.shape[1]
    w4 = w2.reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    y = F.conv2d(x, w4, None, padding=pad_lo).permute(0, 2, 3, 1)
    if wscale is not None:
        y = y * wscale.double()
    if bias is not None:
        y = y + bias.double()
    if emb is not None:
        y = y + emb.double().reshape(-1, 1, 1, cout)
    if res is not None:
        y = y + res.double()                 # may alias `out`: read before the write below
    out.copy_(y.float())
    _add_stats(out, stats)


def attention(q, k, v, o, b, heads, tq, tk, d, scale, strides, o_h16=None):
    def view(t, name, tt):
        sb, sh, st = strides[name]
        return torch.as_strided(t, (b, heads, tt, d), (sb, sh, st, 1))
    qq, kk, vv = view(q, "q", tq).double(), view(k, "k", tk).double(), view(v, "v", tk).double()
    r = (torch.softmax(qq @ kk.transpose(-1, -2) * scale, -1) @ vv).float()
    if o_h16 is not None:
        hi, lo = _split(r)
        view(o_h16[0], "o", tq).copy_(hi)
        view(o_h16[1], "o", tq).copy_(lo)
    else:
        view(o, "o", tq).copy_(r)


def install(ops):
    for name in ("fill_zero", "first_stage_input", "conv_in", "conv_out", "gn_stats_part", "act_prepare", "conv_h16",
                 "attention"):
        setattr(ops, name, globals()[name])
