"""GPU parity tests of the tcgen05 attention core (tfmq_attention_h16) and of the fp16-plane output of tfmq_conv_h16 that
feeds it, through the C ABI.  Checker: float64 torch math (the reference evaluates the attention core in fp32:
ldm/modules/diffusionmodules/openaimodel.py:383-405, quant/quant_block.py:212-245,474-505)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ATTN_TOL = 2e-5      # max-abs error of the attention output against float64 (values O(1)); same bar as tfmq_attention


def _ops():
    from tfmq_b200 import ops
    return ops


def split_planes(ops, x):
    """fp32 [b, t, c] device tensor -> (hi, lo) fp16 planes by the library's own producer (tfmq_act_prepare)."""
    b, t, c = x.shape
    hi = torch.empty((b, t, 1, c), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    ops.act_prepare(x.reshape(b, t, 1, c), dst_h16=(hi, lo))
    return hi.reshape(b, t, c), lo.reshape(b, t, c)


def run_case(dev, b, heads, tq, tk, d, layout, qscale=1.0, planes_out=False):
    ops = _ops()
    g = torch.Generator().manual_seed(tq * 7 + tk + d)
    q = torch.randn(b, heads, tq, d, generator=g) * qscale
    k = torch.randn(b, heads, tk, d, generator=g)
    v = torch.randn(b, heads, tk, d, generator=g)
    scale = d ** -0.5
    att = torch.softmax((q.double() @ k.double().transpose(-1, -2)) * scale, -1)
    ref = att @ v.double()
    if layout == "ldm":
        assert tq == tk
        qkv = torch.empty(b, tq, heads, 3, d)
        qkv[:, :, :, 0] = q.transpose(1, 2)
        qkv[:, :, :, 1] = k.transpose(1, 2)
        qkv[:, :, :, 2] = v.transpose(1, 2)
        hi, lo = split_planes(ops, qkv.reshape(b, tq, heads * 3 * d).to(dev))
        fh, fl = hi.view(-1), lo.view(-1)
        st = (tq * heads * 3 * d, 3 * d, heads * 3 * d)
        qp, kp, vp = (fh, fl), (fh[d:], fl[d:]), (fh[2 * d:], fl[2 * d:])
        strides = dict(q=st, k=st, v=st, o=(tq * heads * d, d, heads * d))
    else:
        def tok(x):
            return x.transpose(1, 2).reshape(b, x.shape[2], heads * d).contiguous().to(dev)
        qp, kp, vp = split_planes(ops, tok(q)), split_planes(ops, tok(k)), split_planes(ops, tok(v))
        strides = dict(q=(tq * heads * d, d, heads * d), k=(tk * heads * d, d, heads * d),
                       v=(tk * heads * d, d, heads * d), o=(tq * heads * d, d, heads * d))
    o = torch.full((b, tq, heads * d), float("nan"), device=dev)
    ops.attention_h16(qp, kp, vp, o, b, heads, tq, tk, d, scale, strides)
    torch.cuda.synchronize()
    got = o.cpu().view(b, tq, heads, d).transpose(1, 2)
    err = (got.double() - ref).abs().max().item()
    if planes_out:
        oh = torch.zeros((b, tq, heads * d), dtype=torch.float16, device=dev)
        ol = torch.zeros_like(oh)
        ops.attention_h16(qp, kp, vp, None, b, heads, tq, tk, d, scale, strides, o_h16=(oh, ol))
        wh, wl = split_planes(ops, o)
        torch.cuda.synchronize()
        assert torch.equal(oh, wh) and torch.equal(ol, wl), "plane output differs from the split of the fp32 output"
    return err


@pytest.mark.parametrize("b,heads,tq,tk,d,layout", [
    (2, 14, 1024, 1024, 32, "ldm"),    # LDM-4 attention at 32x32 (8 query blocks x 8 key tiles per head)
    (2, 21, 256, 256, 32, "ldm"),      # 16x16
    (3, 28, 64, 64, 32, "ldm"),        # 8x8: half-filled query block, one partial key tile
    (1, 8, 300, 300, 40, "tokens"),    # SD v1.4 head dim 40 (runs as 48 with TMA zero fill), ragged tails
    (2, 8, 100, 77, 40, "tokens"),     # SD cross-attention: 77 context tokens
    (2, 4, 130, 70, 64, "tokens"),
    (2, 3, 200, 130, 16, "tokens"),
    (1, 2, 129, 1, 24, "tokens"),      # a single key (cin256's class token)
])
def test_attention_h16_tc(dev, b, heads, tq, tk, d, layout):
    err = run_case(dev, b, heads, tq, tk, d, layout, planes_out=True)
    assert err < ATTN_TOL, f"attention max err {err}"


def test_attention_h16_tc_rescale(dev):
    """Sharp score distributions: the running maximum grows by far more than the lazy-rescale threshold between key
    tiles, so the accumulators in tensor memory are rescaled in place."""
    err = run_case(dev, 2, 4, 384, 1024, 32, "tokens", qscale=12.0)
    assert err < ATTN_TOL, f"attention (rescale path) max err {err}"


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 448, 1344), (4, 8, 8, 896, 2688), (1, 16, 16, 672, 2016)])
def test_conv_h16_plane_output(dev, n, h, w, cin, cout):
    """out_hi / out_lo of tfmq_conv_h16 == the fp32 output pushed through tfmq_act_prepare's split, bit for bit
    (the qkv projection of an LDM AttentionBlock writes the planes the attention kernel reads)."""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + cout)
    wt = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    bias = torch.randn(cout, generator=g) * 0.1
    x = (torch.randn(n, h, w, cin, generator=g) * 2.0).to(dev)
    w_hi, w_lo, wscale = ops.split_h16(wt.to(dev))
    x_hi = torch.empty(x.shape, dtype=torch.float16, device=dev)
    x_lo = torch.empty_like(x_hi)
    ops.act_prepare(x, dst_h16=(x_hi, x_lo))
    out = torch.empty((n, h, w, cout), device=dev)
    ops.conv_h16(x_hi, x_lo, 1, 1, 0, w_hi, w_lo, out, bias=bias.to(dev), wscale=wscale)
    ph = torch.zeros((n, h, w, cout), dtype=torch.float16, device=dev)
    pl = torch.zeros_like(ph)
    ops.conv_h16(x_hi, x_lo, 1, 1, 0, w_hi, w_lo, None, bias=bias.to(dev), wscale=wscale, out_h16=(ph, pl))
    wh = torch.empty_like(ph)
    wl = torch.empty_like(ph)
    ops.act_prepare(out, dst_h16=(wh, wl))
    torch.cuda.synchronize()
    ref = F.linear(x.cpu().double(), wt.double(), bias.double())
    assert (out.cpu().double() - ref).abs().max().item() < 1e-4
    assert torch.equal(ph, wh) and torch.equal(pl, wl)


@pytest.mark.parametrize("dbg", [32, 128, 256, 512, 768])
def test_attention_h16_tc_skewed_warps(dbg):
    """Race hunting: the kernel's debug switches put one softmax warp (or the rescale path) to sleep for 20 us at chosen
    points, so the other warps, the UMMA warps and the TMA producer run as far ahead as the barriers let them.  Every
    result must be unchanged (a parity wait that can alias, or a TMEM buffer handed over too early, shows up here).
    The switch is read once per process: the cases run in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, TFMQ_ATTN_DBG=str(dbg))
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.join(here, "test_gpu_attention_tc.py"), "-k",
                        "test_attention_h16_tc and not skewed"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
