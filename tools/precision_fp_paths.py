"""Accuracy of the two fp32-accurate conv paths (tf32x3, fp16 hi/lo x3) against float64, same data."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tfmq-dm_b200")]
import torch  # noqa: E402

from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (n, h, w, cin, cout) in [(2, 16, 16, 64, 64), (2, 32, 32, 224, 448), (2, 16, 16, 896, 896), (2, 16, 16, 1792, 896),
                             (2, 8, 8, 8064, 256)]:
    x = torch.randn(n, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
    ref = (x.double().reshape(-1, cin) @ wt.double().t()).reshape(n, h, w, cout)
    xd = x.to(dev)
    hi, lo = ops.split_tf32(wt.to(dev))
    o1 = torch.zeros((n, h, w, cout), device=dev)
    ops.conv_fp(xd, 1, 1, 0, hi, lo, o1, passes=3)
    whi, wlo, sc = ops.split_h16(wt.to(dev))
    xh = torch.empty(xd.shape, dtype=torch.float16, device=dev)
    xl = torch.empty_like(xh)
    ops.act_prepare(xd, dst_h16=(xh, xl))
    o2 = torch.zeros_like(o1)
    ops.conv_h16(xh, xl, 1, 1, 0, whi, wlo, o2, wscale=sc)
    # what exact arithmetic on the split operands would give (isolates the tensor core's accumulation)
    xs = (xh.double() + xl.double()).cpu().reshape(-1, cin)
    ws = ((whi.double() + wlo.double()) * sc.double()[:, None]).cpu()
    ref_split = (xs @ ws.t()).reshape(n, h, w, cout)
    torch.cuda.synchronize()
    e1 = (o1.cpu().double() - ref).abs()
    e2 = (o2.cpu().double() - ref).abs()
    e3 = (ref_split - ref).abs()
    print(f"K={cin:5d}: tf32x3 max {e1.max():.2e} rms {e1.pow(2).mean().sqrt():.2e} | h16x3 max {e2.max():.2e} rms "
          f"{e2.pow(2).mean().sqrt():.2e} | split-only max {e3.max():.2e} | ref rms {ref.pow(2).mean().sqrt():.2f}", flush=True)
