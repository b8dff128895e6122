"""Accuracy of quant/tc_autograd.py (forward, dgrad, wgrad on tfmq_conv_h16) against float64, next to torch's own fp32 kernels
(TF32 off).  Prints max |error| / max |value| per tensor.  Run on the GPU box: python tools/precision_tc_autograd.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tfmq-dm_b200"))
import torch
import torch.nn.functional as F
from tfmq_b200.quant.tc_autograd import tc_conv

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()


def case(name, xs, ws, kw, gscale):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(xs, generator=g).to(dev)
    w = (torch.randn(ws, generator=g) * 0.05).to(dev)
    b = torch.randn(ws[0], generator=g).to(dev)
    gy = None
    res = {}
    for tag in ("f64", "torch", "own"):
        dt = torch.float64 if tag == "f64" else torch.float32
        xx, ww, bb = (t.detach().to(dt).requires_grad_(True) for t in (x, w, b))
        if tag == "own":
            y = tc_conv(xx, ww, bb, kw)
            assert y is not None
        elif len(ws) == 4:
            y = F.conv2d(xx, ww, bb, **kw)
        else:
            y = F.linear(xx, ww, bb)
        if gy is None:
            gy = (torch.randn(y.shape, generator=g) * gscale).to(dev)
        gx, gw, gb = torch.autograd.grad(y, (xx, ww, bb), gy.to(dt))
        res[tag] = (y.detach(), gx, gw, gb)
    for tag in ("torch", "own"):
        print(f"{name:34s} {tag:5s} " + "  ".join(f"{n} {rel(a, r):.2e}" for n, a, r in zip(("y", "dx", "dW", "db"), res[tag], res["f64"])))


case("conv3x3 32->64 @16x16 b8", (8, 32, 16, 16), (64, 32, 3, 3), dict(stride=(1, 1), padding=(1, 1)), 1e-4)
case("conv3x3 224->224 @32x32 b8", (8, 224, 32, 32), (224, 224, 3, 3), dict(stride=(1, 1), padding=(1, 1)), 1e-6)
case("conv1x1 64->128 @16x16 b8", (8, 64, 16, 16), (128, 64, 1, 1), dict(stride=(1, 1), padding=(0, 0)), 1e-3)
case("linear 96->384, 8x256 tokens", (8, 256, 96), (384, 96), {}, 1e-5)
case("linear 128->512, 32 rows (TIB)", (32, 128), (512, 128), {}, 1.0)
