"""Per-layer table of the w4a8 conv launches of one LDM-4 step (batch 16): shape, algorithmic int8 ops, the duration of the
corresponding launch in an ncu launch list (`--metrics gpu__time_duration.sum`, program order), achieved TOP/s and the
fraction of the int8 peak bench.py uses.  Runs on the CPU: shapes come from one FP forward of the host model with hooks.
  python tools/layer_table.py profiles/r1j_launches_step.csv > profiles/r1j_conv_layers.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402

from helpers import fp_model  # noqa: E402
from tfmq_b200.quant.quant_layer import QMODE, QuantLayer, Scaler  # noqa: E402
from tfmq_b200.quant.quant_model import QuantModel  # noqa: E402

BATCH = 16
path = sys.argv[1]
# the int8 dense burst peak measured with the MEASURED_PEAKS.json protocol (profiles/r2a_int8_peak.json), as bench.py uses it
_pk = os.path.join(ROOT, "profiles", "r2a_int8_peak.json")
peak = json.load(open(_pk))["int8_tops"] if os.path.exists(_pk) else 3200.0

wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
fp = fp_model("ldm").eval()
order = []
hooks = []
for n, m in fp.named_modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv1d)):
        hooks.append(m.register_forward_hook(lambda mod, inp, out, n=n: order.append((n, tuple(inp[0].shape), tuple(out.shape),
                                                                                     tuple(mod.weight.shape)))))
with torch.no_grad():
    fp(torch.zeros(1, 3, 64, 64), torch.zeros(1))       # FP host model on the CPU: shapes only
for h in hooks:
    h.remove()
qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).eval()
qnn.set_quant_state(True, True)
qnn.disable_out_quantization()
flags = {n: (m.use_wq, m.use_aq and not m.disable_aq) for n, m in qnn.model.named_modules() if isinstance(m, QuantLayer)}
layers = [(n, i, o, w) for n, i, o, w in order if flags.get(n) == (True, True)]
# everything else that is a tensor-core conv runs on the fp16-split path: un-wrapped convs (skip / op / Conv1d qkv, proj_out) and
# weight-only-quantised layers; the first / last conv (3 or fewer channels on one side) have their own FFMA kernels
fp_layers = [(n, i, o, w) for n, i, o, w in order if flags.get(n) != (True, True) and min(w[0], w[1]) > 4]

rows = [l for l in open(path) if not l.startswith("==")]
durs = []
for r in csv.DictReader(rows):
    if "igemm_kernel<0" in r["Kernel Name"] or "igemm_kernel<(int)0" in r["Kernel Name"]:
        v = float(r["Metric Value"].replace(",", ""))
        durs.append(v / 1e3 if r["Metric Unit"] == "ns" else v)
assert len(durs) == len(layers), (len(durs), len(layers))
print(f"# w4a8 conv launches of one LDM-4 step, batch {BATCH} ({os.path.basename(path)}; ncu durations are cold-cache and serialised)")
print(f"# int8 peak used: {peak:.0f} TOP/s (measured burst, 8192^3 s8; profiles/r2a_int8_peak.json)\n")
print("| # | layer | conv | map | K | GOP | us | TOP/s | of peak |")
print("|---:|---|---|---|---:|---:|---:|---:|---:|")
tot_op = tot_us = 0.0
by_map = {}
for k, ((n, i, o, w), us) in enumerate(zip(layers, durs)):
    cout, cin, kh, kw = w
    hh, ww = o[2], o[3]
    gop = 2.0 * BATCH * hh * ww * cout * cin * kh * kw / 1e9
    tot_op += gop
    tot_us += us
    a = by_map.setdefault(f"{hh}x{ww}", [0, 0.0, 0.0])
    a[0] += 1
    a[1] += gop
    a[2] += us
    print(f"| {k} | {n} | {cin}->{cout} {kh}x{kw} | {hh}x{ww} | {cin * kh * kw} | {gop:.1f} | {us:.1f} | {gop / us * 1e3:.0f} | "
          f"{gop / us * 1e3 / peak:.2f} |")
print(f"\ntotal {tot_op / 1e3:.2f} TOP in {tot_us:.0f} us = {tot_op / tot_us * 1e3:.0f} TOP/s = {tot_op / tot_us * 1e3 / peak:.3f} of peak")
print("\n| map | launches | GOP | us | share of kernel time | TOP/s | of peak |")
print("|---|---:|---:|---:|---:|---:|---:|")
for m, (cnt, g, u) in by_map.items():
    print(f"| {m} | {cnt} | {g:.0f} | {u:.0f} | {100 * u / tot_us:.0f} % | {g / u * 1e3:.0f} | {g / u * 1e3 / peak:.2f} |")

fdurs = []
for r in csv.DictReader(rows):
    if "igemm_kernel<3" in r["Kernel Name"] or "igemm_kernel<(int)3" in r["Kernel Name"]:
        v = float(r["Metric Value"].replace(",", ""))
        fdurs.append(v / 1e3 if r["Metric Unit"] == "ns" else v)
if len(fdurs) == len(fp_layers):
    fpeak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    print(f"\n# fp32-accurate conv launches (kind::f16 on fp16 hi/lo planes, 3 products): algorithmic GFLOP, and the executed "
          f"3 x GFLOP against the measured burst f16/bf16 rate ({fpeak:.0f} TFLOP/s)\n")
    print("| # | layer | conv | map | K | GFLOP | us | algorithmic TFLOP/s | executed (x3) of peak |")
    print("|---:|---|---|---|---:|---:|---:|---:|---:|")
    tg = tu = 0.0
    for k, ((n, i, o, w), us) in enumerate(zip(fp_layers, fdurs)):
        cout, cin = w[0], w[1]
        taps = w[2] * (w[3] if len(w) == 4 else 1)
        pix = o[2] * (o[3] if len(o) == 4 else 1)
        gf = 2.0 * BATCH * pix * cout * cin * taps / 1e9
        tg += gf
        tu += us
        shape = f"{o[2]}x{o[3]}" if len(o) == 4 else f"{o[2]} tok"
        print(f"| {k} | {n} | {cin}->{cout} k{w[2]} | {shape} | {cin * taps} | {gf:.1f} | {us:.1f} | {gf / us * 1e3:.0f} | "
              f"{3 * gf / us * 1e3 / fpeak:.2f} |")
    print(f"\ntotal {tg:.0f} GFLOP algorithmic in {tu:.0f} us = {tg / tu * 1e3:.0f} TFLOP/s; executed x3 = {3 * tg / tu * 1e3 / fpeak:.2f} of peak")
else:
    print(f"\n# fp conv table skipped: {len(fdurs)} launches vs {len(fp_layers)} layers")
