"""GPU parity tests of the individual kernels, called through the C ABI (tfmq_b200.ops -> ctypes).
The checker is the CPU oracle (oracle/quant_ref.py) plus exact integer / float64 torch math."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from tfmq_b200 import ops
    return ops


def _qref():
    from oracle import quant_ref
    return quant_ref


def nhwc(t_nchw):
    return t_nchw.permute(0, 2, 3, 1).contiguous()


def to_ohwi(w):  # [O,I,kh,kw] -> [O, kh*kw*I]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


# ------------------------------------------------------------------ pack
@pytest.mark.parametrize("cout,k,ada", [(32, 64, False), (224, 9 * 224, True), (16, 32, True)])
def test_pack_w4_bit_exact(dev, cout, k, ada):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(cout * 7 + k)
    w = torch.randn(cout, k, generator=g) * 0.05
    delta, zp = q.channel_wise(q.minmax_scale, w, 16)
    alpha = None
    if ada:
        alpha = q.adaround_init_alpha(w, delta) + torch.randn(cout, k, generator=g)
        ref = q.adaround_codes(w, delta, zp, alpha, 16)
    else:
        ref = q.uaq_codes(w, delta, zp, 16)
    codes, packed, wsum = ops.pack_w4(w.to(dev), delta.to(dev), zp.to(dev), alpha.to(dev) if ada else None)
    assert torch.equal(codes.cpu().float(), ref)
    # packed layout: byte i of group g = code[g*32+i] | code[g*32+16+i] << 4
    c = ref.to(torch.uint8).view(cout, k // 32, 2, 16)
    exp = (c[:, :, 0] | (c[:, :, 1] << 4)).reshape(cout, k // 2)
    assert torch.equal(packed.cpu(), exp)
    assert torch.equal(wsum.cpu().long(), (ref - zp).sum(1).long())


# ------------------------------------------------------------------ raw int8 GEMM (UMMA descriptors)
@pytest.mark.parametrize("m,n,k", [(128, 16, 128), (256, 224, 256), (384, 256, 1152), (128, 128, 4096)])
def test_gemm_i8_exact(dev, m, n, k):
    ops = _ops()
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randint(0, 256, (m, k), generator=g, dtype=torch.int32)
    b = torch.randint(-15, 16, (n, k), generator=g, dtype=torch.int32)
    out = torch.empty((m, n), dtype=torch.int32, device=dev)
    ops.gemm_i8_peak(a.to(torch.uint8).to(dev), b.to(torch.int8).to(dev), out)
    ref = a.double() @ b.double().t()
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().double(), ref)


# ------------------------------------------------------------------ w4a8 conv
def _w4a8_case(dev, n, h, w, cin, cout, ksize, use_emb, use_res, seed, ld_extra=0):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(seed)
    wt = torch.randn(cout, cin, ksize, ksize, generator=g) * 0.05
    bias = torch.randn(cout, generator=g) * 0.1
    delta_w, zp_w = q.channel_wise(q.minmax_scale, wt, 16)
    x = torch.randn(n, cin, h, w, generator=g)
    delta_a, zp_a = q.minmax_scale(x, 256)
    codes_a = q.uaq_codes(x, delta_a, zp_a, 256)                    # [n,cin,h,w] float ints
    # ---- exact reference in float64 on de-quantised integers
    wq = q.uaq_codes(wt, delta_w, zp_w, 16)
    xi = (codes_a - zp_a).double()
    wi = (wq - zp_w).double()
    acc = F.conv2d(xi, wi, None, padding=ksize // 2)
    ref = acc * (delta_a.double() * delta_w.double().view(1, -1, 1, 1)) + bias.double().view(1, -1, 1, 1)
    emb = res = None
    if use_emb:
        emb = torch.randn(n, cout, generator=g)
        ref = ref + emb.double()[:, :, None, None]
    if use_res:
        res = torch.randn(n, cout, h, w, generator=g)
        ref = ref + res.double()
    # ---- device
    halo = ksize // 2
    act = torch.full((n, h + 2 * halo, w + 2 * halo, cin), int(zp_a.item()), dtype=torch.uint8)
    act[:, halo:halo + h, halo:halo + w, :] = nhwc(codes_a).to(torch.uint8)
    _, packed, wsum = ops.pack_w4(to_ohwi(wt).to(dev), delta_w.to(dev), zp_w.to(dev))
    buf = torch.zeros((n, h, w, cout + ld_extra), dtype=torch.float32, device=dev)
    out = buf[..., :cout]
    res_d = None
    if use_res:
        out.copy_(nhwc(res).to(dev))   # residual aliases the output, as in the engine
        res_d = out
    aq = torch.tensor([delta_a.item(), zp_a.item()], dtype=torch.float32, device=dev)
    ops.conv_w4a8(act.to(dev), ksize, packed, zp_w.reshape(-1).to(torch.int32).to(dev),
                  delta_w.reshape(-1).contiguous().to(dev), wsum, bias.to(dev), aq, out,
                  emb=emb.to(dev) if use_emb else None, res=res_d)
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 3, 1, 2).double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-6 * scale + 1e-6, f"max err {err} (scale {scale})"
    # and the fp32 reference path (quant_layer_forward) agrees within fp32 accumulation error
    ref32 = q.quant_layer_forward(x, wt, bias, wq=(delta_w, zp_w), aq=(delta_a, zp_a),
                                  conv=dict(padding=ksize // 2))
    if use_emb:
        ref32 = ref32 + emb[:, :, None, None]
    if use_res:
        ref32 = ref32 + res
    assert (got.float() - ref32).abs().max().item() <= 1e-4 * max(scale, 1.0)


@pytest.mark.parametrize("n,h,w,cin,cout,ks,emb,res", [
    (2, 16, 16, 128, 32, 3, False, False),     # one k-block per tap
    (1, 64, 64, 224, 224, 3, True, True),      # LDM-4 top level: ragged 224 = 128 + 96 channels
    (3, 8, 8, 160, 48, 3, True, False),        # tile spans two images; batch tail
    (2, 32, 32, 256, 256, 1, False, True),     # CIFAR attention projection (1x1)
    (5, 1, 1, 512, 256, 1, False, False),      # linear, M = 5 rows
    (16, 4, 4, 512, 256, 3, True, True),       # CIFAR lowest resolution
    (2, 32, 32, 448, 448, 3, True, True),      # two N tiles
    (8, 64, 64, 32, 96, 3, True, True),        # several units per persistent CTA pair, 3 residual chunks each
    (6, 64, 64, 64, 224, 1, False, True),      # 7 residual chunks per unit, residual loads two chunks ahead across units
    (6, 64, 64, 64, 224, 1, False, False),
])
def test_conv_w4a8_exact(dev, n, h, w, cin, cout, ks, emb, res):
    _w4a8_case(dev, n, h, w, cin, cout, ks, emb, res, seed=n * 1000 + cin + cout + ks)


def test_conv_w4a8_zero_points_outside_the_code_range(dev):
    """Scaler.MSE does not force the weight range to contain 0 (quant/quant_layer.py:38-64): a channel whose weights share one
    sign gets a zero point below 0 or above 15.  The reference handles it through clamp(round(w / d) + zp, 0, 15) and
    d * (q - zp); here the codes stay in [0, 15] and the epilogue folds zp in as int32.  Exact against float64."""
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    n, h, w, cin, cout = 2, 16, 16, 64, 32
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    wt[0] = wt[0].abs() + 0.5            # all positive, far from 0: zp << 0
    wt[1] = -wt[1].abs() - 0.2           # all negative: zp >> 15
    wt[2] = wt[2].abs() + 0.01
    w2 = wt.reshape(cout, -1)
    lo, hi = w2.min(1).values, w2.max(1).values
    delta = (hi - lo) / 15
    zp = torch.round(-lo / delta)
    assert zp[0] < 0 and zp[1] > 15
    codes = torch.clamp(torch.round(wt / delta.view(-1, 1, 1, 1)) + zp.view(-1, 1, 1, 1), 0, 15)
    x = torch.randn(n, cin, h, w, generator=g)
    da, za = (x.max() - x.min()) / 255, torch.round(-x.min() / ((x.max() - x.min()) / 255))
    qa = torch.clamp(torch.round(x / da) + za, 0, 255)
    ref = F.conv2d((qa - za).double(), (codes - zp.view(-1, 1, 1, 1)).double(), padding=1) * (
        da.double() * delta.double().view(1, -1, 1, 1))
    act = torch.full((n, h + 2, w + 2, cin), int(za.item()), dtype=torch.uint8)
    act[:, 1:-1, 1:-1, :] = nhwc(qa).to(torch.uint8)
    got_codes, packed, wsum = ops.pack_w4(to_ohwi(wt).to(dev), delta.to(dev), zp.to(dev))
    assert torch.equal(got_codes.cpu().float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), codes)
    assert torch.equal(wsum.cpu().double(), (codes - zp.view(-1, 1, 1, 1)).double().sum((1, 2, 3)))
    out = torch.zeros((n, h, w, cout), device=dev)
    aq = torch.tensor([da.item(), za.item()], device=dev)
    ops.conv_w4a8(act.to(dev), 3, packed, zp.to(torch.int32).to(dev), delta.contiguous().to(dev), wsum,
                  torch.zeros(cout, device=dev), aq, out)
    torch.cuda.synchronize()
    err = (out.cpu().permute(0, 3, 1, 2).double() - ref).abs().max().item()
    assert err <= 2e-6 * ref.abs().max().item() + 1e-6, err


def test_conv_w4a8_strided_output(dev):
    _w4a8_case(dev, 2, 16, 16, 64, 64, 3, True, True, seed=5, ld_extra=32)


# ------------------------------------------------------------------ fp (tf32x3) conv
@pytest.mark.parametrize("n,h,w,cin,cout,ks,stride,pad_lo", [
    (2, 32, 32, 224, 448, 1, 1, 0),     # skip_connection 1x1
    (1, 64, 64, 224, 224, 3, 1, 1),     # weight-only-quantised first conv
    (2, 32, 32, 224, 224, 3, 2, 1),     # LDM Downsample.op
    (2, 32, 32, 128, 128, 3, 2, 0),     # DDIM Downsample (pad right/bottom only)
    (4, 8, 8, 896, 2688, 1, 1, 0),      # LDM qkv Conv1d
])
def test_conv_fp_tf32x3(dev, n, h, w, cin, cout, ks, stride, pad_lo):
    ops = _ops()
    g = torch.Generator().manual_seed(h * 31 + cin + cout + stride)
    wt = torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)
    bias = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(n, cin, h, w, generator=g)
    if ks == 3 and pad_lo == 0:
        xin = F.pad(x, (0, 1, 0, 1))
        ref = F.conv2d(xin.double(), wt.double(), bias.double(), stride=stride)
    else:
        ref = F.conv2d(x.double(), wt.double(), bias.double(), stride=stride, padding=ks // 2)
    oh, ow = ref.shape[2], ref.shape[3]
    res = torch.randn(n, cout, oh, ow, generator=g)
    ref = ref + res.double()
    hi, lo = ops.split_tf32(to_ohwi(wt).to(dev))
    out = nhwc(res).to(dev)
    ops.conv_fp(nhwc(x).to(dev), ks, stride, pad_lo, hi, lo, out, bias=bias.to(dev), res=out, passes=3)
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 3, 1, 2).double()
    err = (got - ref).abs().max().item()
    # the tensor-core fp32 accumulator truncates: error grows with the number of accumulation steps (K/8)
    steps = cin * ks * ks / 8
    tol = max(3e-6, 4 * 2.0 ** -26 * steps) * max(1.0, ref.abs().max().item())
    assert err <= tol, f"tf32x3 max err {err} (tol {tol})"
    # single pass is plain tf32: coarse agreement only
    out1 = torch.zeros_like(out)
    ops.conv_fp(nhwc(x).to(dev), ks, stride, pad_lo, hi, lo, out1, bias=bias.to(dev), passes=1)
    torch.cuda.synchronize()
    err1 = (out1.cpu().permute(0, 3, 1, 2).double() - (ref - res.double())).abs().max().item()
    assert err1 <= 2e-2, f"tf32 single-pass err {err1}"


# ------------------------------------------------------------------ fp (fp16 hi/lo split x3) conv
@pytest.mark.parametrize("n,h,w,cin,cout,ks,stride,pad_lo", [
    (2, 32, 32, 224, 448, 1, 1, 0),     # skip_connection 1x1 (K tail: 224 = 3 * 64 + 32)
    (1, 64, 64, 224, 224, 3, 1, 1),     # weight-only-quantised first conv
    (2, 32, 32, 224, 224, 3, 2, 1),     # LDM Downsample.op
    (2, 32, 32, 128, 128, 3, 2, 0),     # DDIM Downsample (pad right/bottom only)
    (4, 8, 8, 896, 2688, 1, 1, 0),      # LDM qkv Conv1d
    (16, 16, 16, 1568, 672, 1, 1, 0),   # skip over a concat, many tiles per CTA
])
def test_conv_h16x3(dev, n, h, w, cin, cout, ks, stride, pad_lo):
    """kind::f16 on fp16 hi/lo planes: the activation planes come from act_prepare's split output, the weight planes
    from split_h16; the result must be as accurate as the tf32x3 path (accumulation-order level)."""
    ops = _ops()
    g = torch.Generator().manual_seed(h * 31 + cin + cout + stride)
    wt = torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)
    wt[0] *= 300.0        # per-channel weight scales far apart
    wt[1] *= 1e-4
    bias = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(n, cin, h, w, generator=g) * 3.0
    x[0, 0, 0, 0] = 1234.5
    x[0, 1, 0, 0] = 3e-6
    if ks == 3 and pad_lo == 0:
        xin = F.pad(x, (0, 1, 0, 1))
        ref = F.conv2d(xin.double(), wt.double(), bias.double(), stride=stride)
    else:
        ref = F.conv2d(x.double(), wt.double(), bias.double(), stride=stride, padding=ks // 2)
    oh, ow = ref.shape[2], ref.shape[3]
    res = torch.randn(n, cout, oh, ow, generator=g)
    ref = ref + res.double()
    w_hi, w_lo, wscale = ops.split_h16(to_ohwi(wt).to(dev))
    assert w_lo is not None
    xd = nhwc(x).to(dev)
    x_hi = torch.empty(xd.shape, dtype=torch.float16, device=dev)
    x_lo = torch.empty_like(x_hi)
    ops.act_prepare(xd, dst_h16=(x_hi, x_lo))
    torch.cuda.synchronize()
    # the split itself: hi + lo reproduces x to 2^-21 relative (22 significand bits), or 2^-25 absolute for tiny values
    back = x_hi.double() + x_lo.double()
    assert ((back - xd.double()).abs() <= 2.0 ** -21 * xd.double().abs() + 2.0 ** -24).all()
    out = nhwc(res).to(dev)
    ops.conv_h16(x_hi, x_lo, ks, stride, pad_lo, w_hi, w_lo, out, bias=bias.to(dev), wscale=wscale, res=out)
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 3, 1, 2).double()
    # same bound as the tf32x3 path (the tensor-core fp32 accumulator truncates: error grows with the number of
    # accumulation steps, K/16 here), per output channel because the channels' weight scales are far apart
    steps = cin * ks * ks / 16
    tol = max(3e-6, 8 * 2.0 ** -26 * steps)     # 2x the tf32 test's factor: the planted 1234.5 sits early in the sum
    allowed = tol * ref.abs().amax(dim=(0, 2, 3)).clamp_min(1.0)
    err = ((got - ref).abs().amax(dim=(0, 2, 3)) / allowed).max().item()
    assert err <= 1.0, f"h16x3 err / allowed = {err}"


def test_conv_h16_integer_weights_exact(dev):
    """Weight-only-quantised layer: integer weights (code - zp) are exact in fp16, no lo plane, per-channel delta in wscale."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    n, h, w, cin, cout = 2, 16, 16, 64, 64
    wi = torch.randint(-8, 8, (cout, 3 * 3 * cin), generator=g).float()
    delta = torch.rand(cout, generator=g) * 0.01 + 0.001
    x = torch.randn(n, h, w, cin, generator=g)
    w_hi, w_lo, wscale = ops.split_h16(wi.to(dev), delta.to(dev))
    assert w_lo is None
    xd = x.to(dev)
    x_hi = torch.empty(xd.shape, dtype=torch.float16, device=dev)
    x_lo = torch.empty_like(x_hi)
    ops.act_prepare(xd, dst_h16=(x_hi, x_lo))
    out = torch.zeros((n, h, w, cout), device=dev)
    ops.conv_h16(x_hi, x_lo, 3, 1, 1, w_hi, None, out, wscale=wscale)
    torch.cuda.synchronize()
    wt = (wi * delta[:, None]).reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), padding=1)
    err = (out.cpu().permute(0, 3, 1, 2).double() - ref).abs().max().item()
    assert err <= 2e-6 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("m,n,k,ksplit,conv", [(384, 224, 16384, -1, False), (1344, 224, 9472, 7, False),
                                               (130, 48, 4096, 3, False), (256, 64, 131072, -1, False),
                                               (2, 96, 96, 5, True)])
def test_conv_h16_split_k(dev, m, n, k, ksplit, conv):
    """Split-K of the fp16-split kernel (tfmq_conv_h16_desc.ksplit): the k-blocks of an output tile are shared by several
    CTAs and the partial tiles are added into the cleared output by TMA reduce -- the weight-gradient GEMMs of the
    reconstruction loop (few output tiles, K = every pixel of the batch).  As a GEMM (1x1, one pixel per row) with ragged M,
    a bias (must be added once, not per K range), an output that holds garbage before the call; and as a 3x3 conv whose
    K ranges cut through the taps.  Against float64."""
    ops = _ops()
    g = torch.Generator().manual_seed(m + n)
    if conv:
        x = torch.randn(m, 16, 16, k, generator=g)
        wt = torch.randn(n, 3 * 3 * k, generator=g) / math.sqrt(9 * k)
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.view(n, 3, 3, k).permute(0, 3, 1, 2).double(), padding=1)
        ref = ref.permute(0, 2, 3, 1)
        ks = 3
    else:
        x = torch.randn(m, 1, 1, k, generator=g)
        wt = torch.randn(n, k, generator=g) / math.sqrt(k)
        ref = (x.view(m, k).double() @ wt.double().t()).view(m, 1, 1, n)
        ks = 1
    bias = torch.randn(n, generator=g)
    ref = ref + bias.double()
    w_hi, w_lo, wscale = ops.split_h16(wt.to(dev), keep_lo=True)
    xd = x.to(dev)
    x_hi = torch.empty(xd.shape, dtype=torch.float16, device=dev)
    x_lo = torch.empty_like(x_hi)
    ops.act_prepare(xd, dst_h16=(x_hi, x_lo))
    out = torch.full(ref.shape, 7.0, device=dev)
    ops.conv_h16(x_hi, x_lo, ks, 1, ks // 2, w_hi, w_lo, out, bias=bias.to(dev), wscale=wscale, ksplit=ksplit)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= max(3e-6, 8 * 2.0 ** -26 * (k * ks * ks / 16)) * max(1.0, ref.abs().max().item()), err
    with pytest.raises(RuntimeError):        # no residual with split-K
        ops.conv_h16(x_hi, x_lo, ks, 1, ks // 2, w_hi, w_lo, out, wscale=wscale, res=out, ksplit=2)


# ------------------------------------------------------------------ GroupNorm + producer
@pytest.mark.parametrize("n,h,w,c,eps", [(2, 16, 16, 128, 1e-6), (3, 32, 32, 224, 1e-5), (1, 8, 8, 1792, 1e-5)])
def test_gn_silu_quant(dev, n, h, w, c, eps):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(c + h)
    x = torch.randn(n, c, h, w, generator=g) * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    y = q.silu(F.group_norm(x, 32, gamma, beta, eps))
    delta, zp = q.minmax_scale(y, 256)
    ref_codes = q.uaq_codes(y, delta, zp, 256)
    xd = nhwc(x).to(dev)
    stats = ops.gn_stats(xd, 32)
    xs = x.double().view(n, 32, -1)
    assert torch.allclose(stats[..., 0].cpu(), xs.sum(-1), rtol=1e-6, atol=1e-4)
    assert torch.allclose(stats[..., 1].cpu(), (xs * xs).sum(-1), rtol=1e-6, atol=1e-4)
    aq = torch.tensor([delta.item(), zp.item()], device=dev)
    dst = torch.zeros((n, h + 2, w + 2, c), dtype=torch.uint8, device=dev)
    ops.act_prepare(xd, aq=aq, dst_u8=dst, halo=1, gn_stats_t=stats, gamma=gamma.to(dev), beta=beta.to(dev),
                    groups=32, eps=eps, silu=True)
    f32 = torch.empty((n, h, w, c), device=dev)
    ops.act_prepare(xd, dst_f32=f32, gn_stats_t=stats, gamma=gamma.to(dev), beta=beta.to(dev), groups=32, eps=eps,
                    silu=True)
    torch.cuda.synchronize()
    got = dst.cpu()
    zpi = int(zp.item())
    assert (got[:, 0] == zpi).all() and (got[:, -1] == zpi).all() and (got[:, :, 0] == zpi).all() \
        and (got[:, :, -1] == zpi).all()
    inner = got[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
    diff = (inner - ref_codes).abs()
    assert diff.max().item() <= 1, "codes differ by more than one step"
    flip = (diff > 0).float().mean().item()
    assert flip < 2e-3, f"flip rate {flip}"
    assert (f32.cpu().permute(0, 3, 1, 2) - y).abs().max().item() < 2e-5
    # teacher-forced quantiser: same fp32 input bits -> identical codes
    dst2 = torch.zeros((n, h, w, c), dtype=torch.uint8, device=dev)
    ops.act_prepare(f32, aq=aq, dst_u8=dst2, halo=0)
    torch.cuda.synchronize()
    tf = q.uaq_codes(f32.cpu(), delta, zp, 256)
    assert torch.equal(dst2.cpu().float(), tf)


def test_act_prepare_upsample_and_offset(dev):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 8, 8, 64, generator=g)
    delta, zp = q.minmax_scale(x, 256)
    aq = torch.tensor([delta.item(), zp.item()], device=dev)
    dst = torch.zeros((2, 18, 18, 96), dtype=torch.uint8, device=dev)
    ops.act_prepare(x.to(dev), aq=aq, dst_u8=dst, halo=1, dst_c_off=32, upsample=True)
    torch.cuda.synchronize()
    ref = q.uaq_codes(x, delta, zp, 256).repeat_interleave(2, 1).repeat_interleave(2, 2)
    got = dst.cpu()
    assert torch.equal(got[:, 1:-1, 1:-1, 32:].float(), ref)
    assert (got[..., :32] == 0).all()
    assert (got[:, 0, :, 32:] == int(zp.item())).all()


def test_token_producers_layernorm_and_geglu(dev):
    """LayerNorm -> quantise and GEGLU -> quantise on token rows (BasicTransformerBlock inputs): fp32 output within
    fp32 rounding of torch's, codes equal except where the pre-rounding value sits on a rounding boundary."""
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(21)
    n, t, c = 2, 64, 320
    x = torch.randn(n, t, c, generator=g) * 1.7 + 0.2
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    y = F.layer_norm(x, (c,), gamma, beta, 1e-5)
    xd = x.reshape(n, 1, t, c).to(dev)
    out = torch.empty_like(xd)
    ops.act_prepare(xd, dst_f32=out, ln=(gamma.to(dev), beta.to(dev), 1e-5))
    torch.cuda.synchronize()
    assert (out.cpu().reshape(n, t, c) - y).abs().max().item() < 2e-6
    da, za = q.minmax_scale(y, 256)
    aq = torch.tensor([da.item(), za.item()], device=dev)
    u8 = torch.empty((n, 1, t, c), dtype=torch.uint8, device=dev)
    ops.act_prepare(xd, aq=aq, dst_u8=u8, ln=(gamma.to(dev), beta.to(dev), 1e-5))
    torch.cuda.synchronize()
    ref = q.uaq_codes(y, da, za, 256)
    diff = (u8.cpu().reshape(n, t, c).float() - ref).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 2e-3
    # GEGLU: rows are [value | gate]
    h = torch.randn(n, t, 2 * c, generator=g) * 1.5
    a, gate = h.chunk(2, dim=-1)
    z = a * F.gelu(gate)
    hd = h.reshape(n, 1, t, 2 * c).to(dev)
    outg = torch.empty((n, 1, t, c), device=dev)
    ops.act_prepare(hd, dst_f32=outg, geglu=True)
    torch.cuda.synchronize()
    assert (outg.cpu().reshape(n, t, c) - z).abs().max().item() < 2e-6
    dz, zz = q.minmax_scale(z, 256)
    aqz = torch.tensor([dz.item(), zz.item()], device=dev)
    ops.act_prepare(hd, aq=aqz, dst_u8=u8, geglu=True)
    torch.cuda.synchronize()
    refz = q.uaq_codes(z, dz, zz, 256)
    diffz = (u8.cpu().reshape(n, t, c).float() - refz).abs()
    assert diffz.max().item() <= 1 and (diffz > 0).float().mean().item() < 2e-3


# ------------------------------------------------------------------ small kernels
def test_linear_small_variants(dev):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(11)
    m, i, o = 5, 512, 256
    x = torch.randn(m, i, generator=g)
    w = torch.randn(o, i, generator=g) / math.sqrt(i)
    b = torch.randn(o, generator=g) * 0.1
    out = torch.empty((m, o), device=dev)
    ops.linear_small(x.to(dev), out, w_f32=w.to(dev), bias=b.to(dev))
    assert (out.cpu() - F.linear(x, w, b)).abs().max().item() < 1e-5
    # weight-only quantised (disable_aq), fp input
    dw, zw = q.channel_wise(q.minmax_scale, w, 16)
    codes = q.uaq_codes(w, dw, zw, 16).to(torch.uint8)
    ops.linear_small(x.to(dev), out, codes=codes.to(dev), wzp_f=zw.reshape(-1).to(dev),
                     wdelta=dw.reshape(-1).contiguous().to(dev), bias=b.to(dev))
    ref = q.quant_layer_forward(x, w, b, wq=(dw, zw))
    assert (out.cpu() - ref).abs().max().item() < 1e-5
    # SiLU -> act quant -> w4 (temb_proj / emb_layers path)
    xs = q.silu(x)
    da, za = q.minmax_scale(xs, 256)
    aq = torch.tensor([da.item(), za.item()], device=dev)
    ops.linear_small(x.to(dev), out, codes=codes.to(dev), wzp_f=zw.reshape(-1).to(dev),
                     wdelta=dw.reshape(-1).contiguous().to(dev), bias=b.to(dev), aq=aq, silu_in=True)
    ref = q.quant_layer_forward(xs, w, b, wq=(dw, zw), aq=(da, za))
    assert (out.cpu() - ref).abs().max().item() < 2e-2 * da.item() * 16 + 1e-5


def test_linear_grouped_equals_separate_launches(dev):
    """The Temporal Information Block's per-block projections as one launch: bit-identical to one launch each."""
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(12)
    m, i = 11, 896
    x = torch.randn(m, i, generator=g).to(dev)
    layers, sep = [], []
    for o in (896, 224, 448, 672, 8, 100):
        w = torch.randn(o, i, generator=g) / math.sqrt(i)
        b = (torch.randn(o, generator=g) * 0.1).to(dev)
        dw, zw = q.channel_wise(q.minmax_scale, w, 16)
        codes = q.uaq_codes(w, dw, zw, 16).to(torch.uint8).to(dev)
        da, za = q.minmax_scale(q.silu(x.cpu()), 256)
        aq = torch.tensor([da.item() * (1 + 0.01 * o), za.item()], device=dev)
        kw = dict(x=x, codes=codes, wzp_f=zw.reshape(-1).to(dev), wdelta=dw.reshape(-1).contiguous().to(dev), bias=b,
                  aq=aq, silu_in=True)
        layers.append(dict(kw, out=torch.zeros((m, o), device=dev)))
        sep.append(dict(kw, out=torch.zeros((m, o), device=dev)))
    wf = (torch.randn(64, i, generator=g) / math.sqrt(i)).to(dev)    # an fp member
    layers.append(dict(x=x, w_f32=wf, out=torch.zeros((m, 64), device=dev)))
    sep.append(dict(x=x, w_f32=wf, out=torch.zeros((m, 64), device=dev)))
    grp = ops.LinearGroup(layers)
    grp.run()
    for kw in sep:
        kw = dict(kw)
        xx, out = kw.pop("x"), kw.pop("out")
        ops.linear_small(xx, out, **kw)
    torch.cuda.synchronize()
    for a, b_ in zip(layers, sep):
        assert torch.equal(a["out"], b_["out"])
        assert a["out"].abs().max().item() > 0


@pytest.mark.parametrize("hh,ww", [(16, 16), (12, 20), (40, 8)])     # widths that are / are not multiples of 8
def test_conv_in_out(dev, hh, ww):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, hh, ww, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(64, generator=g)
    out = torch.empty((2, hh, ww, 64), device=dev)
    ops.conv_in(x.to(dev), w.to(dev), b.to(dev), out)
    assert (out.cpu().permute(0, 3, 1, 2) - F.conv2d(x, w, b, padding=1)).abs().max().item() < 1e-5
    h = torch.randn(2, 64, hh, ww, generator=g)
    w2 = torch.randn(3, 64, 3, 3, generator=g) * 0.05
    b2 = torch.randn(3, generator=g)
    o2 = torch.empty((2, 3, hh, ww), device=dev)
    ops.conv_out(nhwc(h).to(dev), w2.to(dev), b2.to(dev), o2)
    assert (o2.cpu() - F.conv2d(h, w2, b2, padding=1)).abs().max().item() < 1e-5


def test_conv_out_model_shapes(dev):
    ops = _ops()
    for n, c, hw in ((1, 128, 32), (2, 224, 64)):
        g = torch.Generator().manual_seed(c)
        h = torch.randn(n, c, hw, hw, generator=g)
        w2 = torch.randn(3, c, 3, 3, generator=g) * 0.05
        b2 = torch.randn(3, generator=g)
        o2 = torch.empty((n, 3, hw, hw), device=dev)
        ops.conv_out(nhwc(h).to(dev), w2.to(dev), b2.to(dev), o2)
        err = (o2.cpu() - F.conv2d(h, w2, b2, padding=1)).abs().max().item()
        assert err < 2e-5, (n, c, hw, err)


def test_ddim_update_and_embedding(dev):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 32, 32, generator=g)
    e = torch.randn(2, 3, 32, 32, generator=g)
    at, an = torch.tensor(0.3), torch.tensor(0.45)
    x0 = (x - e * (1 - at).sqrt()) / at.sqrt()
    c2 = ((1 - an) - 0.0 ** 2).sqrt()
    ref = an.sqrt() * x0 + 0.0 * torch.randn_like(x) * 0 + c2 * e
    coef = torch.stack([at.sqrt(), (1 - at).sqrt(), an.sqrt(), c2, torch.tensor(0.0)]).to(dev)
    xp = torch.empty_like(x, device=dev)
    x0d = torch.empty_like(x, device=dev)
    ops.ddim_update(x.to(dev), e.to(dev), coef, xp, x0d)
    assert (x0d.cpu() - x0).abs().max().item() <= 1e-6
    assert (xp.cpu() - ref).abs().max().item() <= 1e-6
    # eta > 0: c1 * noise, in the DDIM runner's order (coef[5] = 0) and the LDM sampler's (coef[5] = 1), bit for bit
    nz = torch.randn(2, 3, 32, 32, generator=g)
    c1 = 0.8 * ((1 - at / an) * (1 - an) / (1 - at)).abs().sqrt()
    c2n = ((1 - an) - c1 ** 2).sqrt()
    x0f = (x - e * (1 - at).sqrt()) / at.sqrt()
    for order, want in ((0.0, an.sqrt() * x0f + c1 * nz + c2n * e), (1.0, an.sqrt() * x0f + c2n * e + c1 * nz)):
        coef6 = torch.stack([at.sqrt(), (1 - at).sqrt(), an.sqrt(), c2n, c1, torch.tensor(order)]).to(dev)
        ops.ddim_update(x.to(dev), e.to(dev), coef6, xp, x0d, noise=nz.to(dev))
        assert torch.equal(xp.cpu(), want), order
    t = torch.tensor([999.0, 500.0, 1.0])
    for style, dim in ((0, 128), (1, 224)):
        out = torch.empty((3, dim), device=dev)
        ops.timestep_embedding(t.to(dev), dim, style, out)
        half = dim // 2
        if style == 0:
            fr = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
            ref = torch.cat([torch.sin(t[:, None] * fr), torch.cos(t[:, None] * fr)], 1)
        else:
            fr = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
            ref = torch.cat([torch.cos(t[:, None] * fr), torch.sin(t[:, None] * fr)], 1)
        assert (out.cpu() - ref).abs().max().item() < 5e-4


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("b,heads,tq,tk,d,layout", [
    (2, 14, 1024, 1024, 32, "ldm"),    # LDM-4 top attention level (tensor-core path)
    (3, 28, 64, 64, 32, "ldm"),
    (2, 1, 256, 256, 256, "ddim"),     # CIFAR single-head (generic path)
    (2, 8, 100, 77, 40, "tokens"),     # SD cross-attention shape, ragged tq / tk
    (1, 8, 64, 64, 160, "tokens"),
    (2, 8, 200, 77, 160, "tokens"),    # SD deepest cross-attention, ragged (fp16-split kernel at d = 160)
    (1, 8, 300, 300, 40, "tokens"),    # d = 40 zero-padded to three k16 steps, ragged tails
    (2, 4, 130, 70, 80, "tokens"),
    (2, 1, 200, 200, 384, "tokens"),   # cin256: one wide head, warps split the head dimension
    (2, 1, 150, 150, 512, "tokens"),   # first-stage decoder mid.attn_1 (one 512-channel head)
    (2, 1, 70, 70, 576, "tokens"),
    (1, 1, 64, 64, 960, "tokens"),
    (3, 1, 100, 1, 384, "tokens"),     # cin256 cross-attention over a single class token
])
def test_attention_fp32(dev, b, heads, tq, tk, d, layout):
    ops = _ops()
    g = torch.Generator().manual_seed(tq + d)
    q = torch.randn(b, heads, tq, d, generator=g)
    k = torch.randn(b, heads, tk, d, generator=g)
    v = torch.randn(b, heads, tk, d, generator=g)
    scale = d ** -0.5
    att = torch.softmax((q.double() @ k.double().transpose(-1, -2)) * scale, -1)
    ref = att @ v.double()
    if layout == "ldm":
        # token-major [b, t, heads*3*d] with per-head (q|k|v) interleave, as the fp qkv 1x1 conv writes it
        qkv = torch.empty(b, tq, heads, 3, d)
        qkv[:, :, :, 0] = q.transpose(1, 2)
        qkv[:, :, :, 1] = k.transpose(1, 2)
        qkv[:, :, :, 2] = v.transpose(1, 2)
        buf = qkv.reshape(b, tq, heads * 3 * d).to(dev)
        o = torch.empty((b, tq, heads * d), device=dev)
        st = (tq * heads * 3 * d, 3 * d, heads * 3 * d)
        flat = buf.view(-1)
        ops.attention(flat, flat[d:], flat[2 * d:], o, b, heads, tq, tk, d, scale,
                      dict(q=st, k=st, v=st, o=(tq * heads * d, d, heads * d)))
        got = o.cpu().view(b, tq, heads, d).transpose(1, 2)
    else:
        # separate [b, t, heads*d] token-major tensors
        def tok(x):
            return x.transpose(1, 2).reshape(b, x.shape[2], heads * d).contiguous().to(dev)
        qd, kd, vd = tok(q), tok(k), tok(v)
        o = torch.empty((b, tq, heads * d), device=dev)
        ops.attention(qd, kd, vd, o, b, heads, tq, tk, d, scale,
                      dict(q=(tq * heads * d, d, heads * d), k=(tk * heads * d, d, heads * d),
                           v=(tk * heads * d, d, heads * d), o=(tq * heads * d, d, heads * d)))
        got = o.cpu().view(b, tq, heads, d).transpose(1, 2)
    torch.cuda.synchronize()
    err = (got.double() - ref).abs().max().item()
    assert err < 2e-5, f"attention max err {err}"


@pytest.mark.parametrize("b,heads,tq,d", [(2, 14, 256, 32), (1, 8, 100, 40), (2, 1, 90, 512), (1, 2, 70, 160)])
def test_attention_fp16_plane_output(dev, b, heads, tq, d):
    """o_hi / o_lo output == the fp32 output pushed through tfmq_act_prepare's split, bit for bit."""
    ops = _ops()
    g = torch.Generator().manual_seed(d)
    c = heads * d
    q, k, v = (torch.randn(b, tq, c, generator=g).to(dev) for _ in range(3))
    st = dict(q=(tq * c, d, c), k=(tq * c, d, c), v=(tq * c, d, c), o=(tq * c, d, c))
    o = torch.empty((b, tq, c), device=dev)
    ops.attention(q, k, v, o, b, heads, tq, tq, d, d ** -0.5, st)
    hi = torch.zeros((b, tq, c), dtype=torch.float16, device=dev)
    lo = torch.zeros_like(hi)
    ops.attention(q, k, v, None, b, heads, tq, tq, d, d ** -0.5, st, o_h16=(hi, lo))
    rhi = torch.empty((b, tq, 1, c), dtype=torch.float16, device=dev)
    rlo = torch.empty_like(rhi)
    ops.act_prepare(o.view(b, tq, 1, c), dst_h16=(rhi, rlo))
    assert torch.equal(hi.view(-1), rhi.view(-1)) and torch.equal(lo.view(-1), rlo.view(-1))
    assert (hi.float() + lo.float() - o).abs().max().item() <= 2.0 ** -21 * o.abs().max().item()
    with pytest.raises(RuntimeError):      # FFMA kernel (head dim 24): no plane output
        ops.attention(q, k, v, None, b, 1, tq, tq, 24, 0.2, st, o_h16=(hi, lo))


# ------------------------------------------------------------------ calibration kernels
def test_minmax_and_mse_search(dev):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(9)
    w = torch.randn(48, 576, generator=g) * 0.1
    mm = ops.minmax_rows(w.to(dev)).cpu()
    assert torch.equal(mm[:, 0], w.min(1).values) and torch.equal(mm[:, 1], w.max(1).values)
    delta, zp = ops.mse_scale_search(w.to(dev), 16)
    delta, zp = delta.cpu(), zp.cpu()
    agree = 0
    for r in range(w.shape[0]):
        d_ref, z_ref, i_ref = q.mse_scale(w[r], 16, return_index=True)
        cands = q.mse_candidates(w[r].min().item(), w[r].max().item(), 16)
        # the device result must be exactly one of the reference's 80 candidates ...
        match = [i for i, (d, z) in enumerate(cands) if d.item() == delta[r].item() and z.item() == zp[r].item()]
        assert match, f"row {r}: ({delta[r].item()}, {zp[r].item()}) is not a reference candidate"
        agree += int(i_ref in match)
        # ... and score within fp32 noise of the reference's pick
        s_dev = q.lp_loss(q.uaq_fake_quant(w[r], delta[r], zp[r], 16), w[r], 2.4, True)
        s_ref = q.lp_loss(q.uaq_fake_quant(w[r], d_ref, z_ref, 16), w[r], 2.4, True)
        assert s_dev <= s_ref * (1 + 1e-4)
    assert agree >= w.shape[0] - 2
    # per-tensor activations, 8 bit
    x = torch.randn(1, 50000, generator=g)
    d1, z1 = ops.mse_scale_search(x.to(dev), 256)
    d_ref, z_ref = q.mse_scale(x[0], 256)
    s_dev = q.lp_loss(q.uaq_fake_quant(x[0], d1.cpu()[0], z1.cpu()[0], 256), x[0], 2.4, True)
    s_ref = q.lp_loss(q.uaq_fake_quant(x[0], d_ref, z_ref, 256), x[0], 2.4, True)
    assert s_dev <= s_ref * (1 + 1e-4)


def test_act_range_update(dev):
    ops, q = _ops(), _qref()
    from tfmq_b200 import ops as O
    g = torch.Generator().manual_seed(10)
    x0 = torch.randn(4, 8, 8, 64, generator=g)
    xmin, xmax = x0.min(), x0.max()
    state = torch.empty(4, device=dev)
    state[0], state[1] = xmin, xmax
    st_i = state.view(torch.int32)
    st_i[2] = torch.tensor(float("inf")).view(torch.int32)
    st_i[3] = (torch.tensor(float("-inf")).view(torch.int32) ^ 0x7FFFFFFF)
    aq = torch.zeros(2, device=dev)
    for it in range(3):
        x = torch.randn(4, 8, 8, 64, generator=g) * (1 + it)
        xmin, xmax, d_ref, z_ref = q.act_momentum_update(x, xmin, xmax)
        O.act_range_update(x.to(dev), state, aq)
        torch.cuda.synchronize()
        assert state[0].item() == xmin.item() and state[1].item() == xmax.item()
        assert aq[0].item() == d_ref.item() and aq[1].item() == z_ref.item()


def test_adaround_kernels_vs_autograd(dev):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(12)
    cout, k = 32, 288
    w = torch.randn(cout, k, generator=g) * 0.05
    delta, zp = q.channel_wise(q.minmax_scale, w, 16)
    alpha0 = q.adaround_init_alpha(w, delta)
    out = torch.empty((cout, k), device=dev)
    dl, zl = delta.reshape(-1).contiguous().to(dev), zp.reshape(-1).contiguous().to(dev)
    ops.adaround_soft(w.to(dev), dl, zl, alpha0.to(dev), 16, out)
    ref = q.adaround_fake_quant(w, delta, zp, alpha0, 16, soft=True)
    assert (out.cpu() - ref).abs().max().item() < 1e-6
    # three optimiser steps against torch autograd + torch.optim.Adam
    alpha_t = alpha0.clone().requires_grad_(True)
    opt = torch.optim.Adam([alpha_t])
    tgt = torch.randn(cout, k, generator=g) * 0.05
    alpha_d = alpha0.clone().to(dev)
    m = torch.zeros_like(alpha_d)
    v = torch.zeros_like(alpha_d)
    for step in range(1, 4):
        b, lam = 20.0 - step, 0.01
        opt.zero_grad()
        ws = q.adaround_fake_quant(w, delta, zp, alpha_t, 16, soft=True)
        rec = ((ws - tgt) ** 2).sum()
        rl = q.round_loss(alpha_t, b, lam)
        (rec + rl).backward()
        # device: same dL/dw_soft, fused chain rule + regulariser + Adam
        gw = (2 * (ws.detach() - tgt)).to(dev)
        rl_d = torch.zeros(1, device=dev)
        ops.adaround_step(w.to(dev), dl, zl, alpha_d, gw, m, v, 16, step, 1e-3, b, lam, rl_d)
        opt.step()
        torch.cuda.synchronize()
        assert abs(rl_d.item() - rl.item()) <= 1e-4 * abs(rl.item()) + 1e-6
        assert (alpha_d.cpu() - alpha_t.detach()).abs().max().item() < 2e-5
    # reconstruction loss + gradient
    pred = torch.randn(4, 3, 8, 8, generator=g)
    tg = torch.randn(4, 3, 8, 8, generator=g)
    loss = torch.zeros(1, device=dev)
    grad = torch.empty_like(pred, device=dev)
    denom = pred.numel() // pred.shape[1]          # lp_loss: sum over dim 1, mean over the rest
    ops.rec_loss(pred.to(dev), tg.to(dev), denom, loss, grad)
    assert abs(loss.item() - q.lp_loss(pred, tg, 2.0).item()) < 1e-4
    assert (grad.cpu() - 2 * (pred - tg) / denom).abs().max().item() < 1e-6


# ------------------------------------------------------------------ GN statistics fused into the conv epilogue
@pytest.mark.parametrize("n,h,w,cin,cout,cpg,ch_off,groups", [
    (2, 32, 32, 64, 224, 7, 0, 32),        # plain: 32 groups of 7 channels
    (2, 16, 16, 64, 448, 35, 672, 32),     # second part of a 672+448 concat: groups of 35 straddle the boundary
    (3, 8, 8, 64, 64, 2, 0, 32),           # tile spans two images (per-image sums in the epilogue), odd batch tail
    (16, 8, 8, 96, 896, 28, 0, 32),        # LDM-4 lowest resolution: CTA pairs, tiles of two images
    (5, 4, 4, 64, 256, 8, 0, 32),          # CIFAR lowest resolution: eight images per tile, batch tail
    (4, 8, 8, 64, 1792, 56, 0, 32),        # wide layer: 2 x tile_n x 8 B may exceed the smem budget -> separate pass
])
def test_conv_epilogue_gn_stats(dev, n, h, w, cin, cout, cpg, ch_off, groups):
    ops, q = _ops(), _qref()
    g = torch.Generator().manual_seed(cout + ch_off)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    bias = torch.randn(cout, generator=g) * 0.1
    delta_w, zp_w = q.channel_wise(q.minmax_scale, wt, 16)
    x = torch.randn(n, cin, h, w, generator=g)
    delta_a, zp_a = q.minmax_scale(x, 256)
    codes_a = q.uaq_codes(x, delta_a, zp_a, 256)
    act = torch.full((n, h + 2, w + 2, cin), int(zp_a.item()), dtype=torch.uint8)
    act[:, 1:-1, 1:-1, :] = nhwc(codes_a).to(torch.uint8)
    _, packed, wsum = ops.pack_w4(to_ohwi(wt).to(dev), delta_w.to(dev), zp_w.to(dev))
    out = torch.zeros((n, h, w, cout), device=dev)
    aq = torch.tensor([delta_a.item(), zp_a.item()], device=dev)
    stats = torch.zeros((n, groups, 2), dtype=torch.float64, device=dev)
    ops.conv_w4a8(act.to(dev), 3, packed, zp_w.reshape(-1).to(torch.int32).to(dev),
                  delta_w.reshape(-1).contiguous().to(dev), wsum, bias.to(dev), aq, out, stats=[(stats, cpg, ch_off)])
    torch.cuda.synchronize()
    o = out.cpu().double()                                   # [n,h,w,cout]
    ref = torch.zeros(n, groups, 2, dtype=torch.float64)
    for ch in range(cout):
        gi = (ch_off + ch) // cpg
        ref[:, gi, 0] += o[..., ch].sum((1, 2))
        ref[:, gi, 1] += (o[..., ch] ** 2).sum((1, 2))
    assert torch.allclose(stats.cpu(), ref, rtol=1e-5, atol=1e-3)
    # same through the fp (tf32) conv
    hi, lo = ops.split_tf32(to_ohwi(wt).to(dev))
    out2 = torch.zeros_like(out)
    stats2 = torch.zeros_like(stats)
    ops.conv_fp(nhwc(x).to(dev), 3, 1, 1, hi, lo, out2, bias=bias.to(dev), stats=[(stats2, cpg, ch_off)])
    torch.cuda.synchronize()
    o2 = out2.cpu().double()
    ref2 = torch.zeros_like(ref)
    for ch in range(cout):
        gi = (ch_off + ch) // cpg
        ref2[:, gi, 0] += o2[..., ch].sum((1, 2))
        ref2[:, gi, 1] += (o2[..., ch] ** 2).sum((1, 2))
    assert torch.allclose(stats2.cpu(), ref2, rtol=1e-5, atol=1e-3)
