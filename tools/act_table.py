"""Per-launch table of the activation producer (tfmq_act_prepare) inside one LDM-4 batch-16 step: shape, mode, device time
(CUDA events around each launch in an eager pass), algorithmic HBM bytes and GB/s against the measured copy bandwidth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
qnn, eng, ts = bench.build_quantised(dev, bench.BATCH)
rows = []
orig = ops.act_prepare


def timed(src, **kw):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    orig(src, **kw)
    e.record()
    n, h, w, c = src.shape
    up = 4 if kw.get("upsample") else 1
    if kw.get("dst_u8") is not None:
        mode, wb = "u8", kw["dst_u8"].numel()
    elif kw.get("dst_h16") is not None:
        mode, wb = "h16", kw["dst_h16"][0].numel() * 4
    else:
        mode, wb = "f32", kw["dst_f32"].numel() * 4
    tag = ("GN+" if kw.get("gn_stats_t") is not None else "") + ("SiLU+" if kw.get("silu") else "") + mode + ("(x2 up)" if up == 4 else "")
    rows.append([tag, (n, h, w, c), src.numel() * 4 + wb, s, e])


ops.act_prepare = timed
saved = eng.x_in.clone()
for rep in range(3):
    rows.clear()
    eng._run_program(True)
    torch.cuda.synchronize()
eng.x_in.copy_(saved)
ops.act_prepare = orig
hbm = bench.peaks()["hbm"]
tot_us = tot_b = 0.0
agg = {}
print(f"{'mode':18s} {'shape':22s} {'MB':>8s} {'us':>8s} {'GB/s':>8s} {'of copy bw':>10s}")
for tag, shape, b, s, e in rows:
    us = s.elapsed_time(e) * 1e3
    tot_us += us
    tot_b += b
    a = agg.setdefault(tag, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += b
    print(f"{tag:18s} {str(shape):22s} {b / 1e6:8.1f} {us:8.1f} {b / us / 1e3:8.0f} {b / us / 1e3 / hbm:10.2f}")
print()
for tag, (n, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{tag:18s} n={n:3d} {us:8.1f} us {b / us / 1e3:8.0f} GB/s = {b / us / 1e3 / hbm:.2f} of the measured copy bandwidth ({hbm:.0f} GB/s)")
print(f"all {len(rows)} launches: {tot_us:.1f} us per step, {tot_b / tot_us / 1e3:.0f} GB/s = {tot_b / tot_us / 1e3 / hbm:.2f} of the copy bandwidth")
