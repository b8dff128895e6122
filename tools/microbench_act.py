"""Activation producer (tfmq_act_prepare, GroupNorm + SiLU + u8 quantise with halo / fp16 split) on the LDM-4 batch-16
shapes: back-to-back launches (throughput) and launches separated by a dependent tiny kernel (latency), CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
hbm = 6551.0
for (n, h, w, c) in [(16, 64, 64, 224), (16, 64, 64, 672), (16, 32, 32, 448), (16, 16, 16, 672), (16, 8, 8, 896)]:
    x = torch.randn(n, h, w, c, device=dev)
    stats = ops.gn_stats(x, 32)
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
    aq = torch.tensor([0.02, 14.0], device=dev)
    u8 = torch.empty((n, h + 2, w + 2, c), dtype=torch.uint8, device=dev)
    hi = torch.empty((n, h, w, c), dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi)
    for tag, fn, nbytes in (
            ("GN+SiLU+u8", lambda: ops.act_prepare(x, aq=aq, dst_u8=u8, halo=1, gn_stats_t=stats, gamma=gamma, beta=beta, silu=True),
             x.numel() * 4 + u8.numel()),
            ("h16 split", lambda: ops.act_prepare(x, dst_h16=(hi, lo)), x.numel() * 8)):
        g = torch.cuda.CUDAGraph()
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / 100
        print(f"{tag:12s} {str((n, h, w, c)):22s} {nbytes / 1e6:7.1f} MB  {us:7.1f} us per launch in a 20-launch graph  "
              f"{nbytes / us / 1e3:6.0f} GB/s = {nbytes / us / 1e3 / hbm:.2f} of the copy bandwidth", flush=True)
