"""Step time of the SpatialTransformer UNets (BASELINE configs[2] / [4]) on one B200, guided sampling:
  python tools/bench_sd.py sd_v14 2      (1 prompt x 2 guidance halves per GPU = configs[2]'s per-GPU share)
  python tools/bench_sd.py cin256 16     (8 classes x 2 per GPU = configs[4]'s per-GPU share)
Prints ms per denoising step (CUDA events over graph replays) and the per-kernel shares of one eager step."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from helpers import synth  # noqa: E402
from tfmq_b200 import ops  # noqa: E402
from tfmq_b200.host import ldm_unet as H  # noqa: E402
from tfmq_b200.quant.quant_layer import QMODE, Scaler  # noqa: E402
from tfmq_b200.quant.quant_model import QuantModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sd_v14"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(sd_v14=H.sd_v14_config, cin256=H.cin256_config)[name]()
tk = 77 if name == "sd_v14" else 1
scale = 7.5 if name == "sd_v14" else 3.0
dev = torch.device("cuda:0")
fp = H.UNetModel(**cfg).eval()
synth.fill_state_dict(fp, 7)
fp = fp.to(dev)
wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).eval()
x = synth.latents((nb, cfg["in_channels"], 64, 64), 21).to(dev)
t = torch.full((nb,), 601.0, device=dev)
ctx = synth.latents((nb, tk, cfg["context_dim"]), 22).to(dev)
qnn.set_quant_state(True, True)
qnn.disable_out_quantization()
with torch.no_grad():
    qnn(x[:2], t[:2], ctx[:2])
    eng = qnn.build_engine(batch=nb, context_shape=(tk, cfg["context_dim"]))
eng.set_schedule([601.0] * 8, None, [[0.9, 0.4, 0.92, 0.39, 0.0]] * 8)
eng.set_guidance(scale)
eng.ctx_in.copy_(ctx)
eng.x_in.copy_(x)
for k in range(4):
    eng.step(k)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
n = 20
for k in range(n):
    eng.step(k % 8)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
print(f"{name}: engine batch {nb} (guidance halves included): {ms:.2f} ms per step, {eng.launches_per_step} launches; "
      f"{nb // 2} images per {50 if name == 'sd_v14' else 250} steps -> {nb // 2 / (ms * 1e-3 * (50 if name == 'sd_v14' else 250)):.2f} images/s")
# eager per-op timing by kernel family
names = ["act_prepare", "conv_w4a8", "conv_h16", "conv_fp", "attention", "gn_stats_part", "linear_small", "conv_in", "conv_out"]
acc = collections.defaultdict(lambda: [0, []])
attn_calls = []
orig = {k: getattr(ops, k) for k in names}


def wrap(k):
    def f(*a, **kw):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        r = orig[k](*a, **kw)
        e_.record()
        acc[k][0] += 1
        acc[k][1].append((s_, e_))
        if k == "attention":
            attn_calls.append(((a[4], a[5], a[6], a[7], a[8]), s_, e_))      # (b, heads, tq, tk, d)
        return r
    return f


for k in names:
    setattr(ops, k, wrap(k))
eng.use_graph = False
eng.step(0)
torch.cuda.synchronize()
tot = 0.0
rows = []
for k, (cnt, evs) in acc.items():
    us = sum(a.elapsed_time(b) for a, b in evs) * 1e3
    rows.append((us, k, cnt))
    tot += us
for us, k, cnt in sorted(rows, reverse=True):
    print(f"  {k:16s} n={cnt:4d} {us:9.1f} us  {100 * us / tot:5.1f} %")

by_shape = collections.defaultdict(lambda: [0, 0.0])
for shape, a_, b_ in attn_calls:
    by_shape[shape][0] += 1
    by_shape[shape][1] += a_.elapsed_time(b_) * 1e3
print("  attention by (batch, heads, Tq, Tk, d):")
for shape, (cnt, us) in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
    print(f"    {shape}  n={cnt:2d} {us:9.1f} us")
