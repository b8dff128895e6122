"""Times the attention core on the LDM-4 (batch 16) and SD v1.4 shapes: tcgen05 kernel (tfmq_attention_h16) and the
mma.sync kernel (tfmq_attention) on the same data.  CUDA events, 20 launches after 3 warm-ups."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=200):
    # enough back-to-back work that the SM clock has ramped up before the timed region
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


import subprocess  # noqa: E402


def smclk():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader"], capture_output=True,
                              text=True, timeout=5).stdout.strip()
    except Exception:      # noqa: BLE001
        return "?"


cases = [(16, 14, 1024, 32), (16, 21, 256, 32), (16, 28, 64, 32)]
if len(sys.argv) > 1 and sys.argv[1] == "sd":
    cases += [(2, 8, 4096, 40), (2, 8, 1024, 80)]
for b, heads, t, d in cases:
    c = heads * d
    qkv = torch.randn(b, t, 1, 3 * c, device=dev)
    hi = torch.empty(qkv.shape, dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi)
    ops.act_prepare(qkv, dst_h16=(hi, lo))
    fh, fl, ff = hi.view(-1), lo.view(-1), qkv.view(-1)
    st = (t * 3 * c, 3 * d, 3 * c)
    oh = torch.empty((b, t, c), dtype=torch.float16, device=dev)
    ol = torch.empty_like(oh)
    strides = dict(q=st, k=st, v=st, o=(t * c, d, c))
    flops = 4.0 * t * t * d * heads * b
    line = f"b{b} h{heads} T{t} d{d}:"
    if d in ops.ATTN_TC_DIMS:
        us = timeit(lambda: ops.attention_h16((fh, fl), (fh[d:], fl[d:]), (fh[2 * d:], fl[2 * d:]), None, b, heads, t, t, d,
                                              d ** -0.5, strides, o_h16=(oh, ol)))
        line += f" tcgen05 {us:7.1f} us ({flops / us / 1e6:6.1f} TFLOP/s algorithmic, x3 executed)"
    us2 = timeit(lambda: ops.attention(ff, ff[d:], ff[2 * d:], None, b, heads, t, t, d, d ** -0.5, strides, o_h16=(oh, ol)))
    line += f" | mma.sync {us2:7.1f} us"
    print(line, "| sm clock", smclk(), flush=True)
