"""Replay one LDM-4 batch-16 denoising step many times from the same latent and count distinct results: a latent race in
the warp-specialised / CTA-pair kernels would show up as a changing output (or a hang -> run under `timeout`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
qnn, eng, ts = bench.build_quantised(dev, bench.BATCH)
x = torch.randn(bench.BATCH, 3, 64, 64, device=dev, generator=torch.Generator(dev).manual_seed(5))
ref = None
distinct = 0
worst = 0.0
for i in range(n):
    eng.select_step(0)
    e = eng.forward(x, ts[0])
    if ref is None:
        ref = e.clone()
    elif not torch.equal(e, ref):
        distinct += 1
        worst = max(worst, (e - ref).abs().max().item())
torch.cuda.synchronize()
print(f"{n} replays of one step: {distinct} differ from the first (max-abs {worst:.3e}); finite: {bool(torch.isfinite(ref).all())}")
