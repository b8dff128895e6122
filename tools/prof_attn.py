"""In-kernel phase counters of the tcgen05 attention kernel on the LDM-4 top-level shape (run with TFMQ_ATTN_PROF=1,
optionally TFMQ_ATTN_DBG / TFMQ_ATTN_KT)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from tfmq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
b, heads, t, d = 16, 14, 1024, 32
c = heads * d
qkv = torch.randn(b, t, 1, 3 * c, device=dev)
hi = torch.empty(qkv.shape, dtype=torch.float16, device=dev)
lo = torch.empty_like(hi)
ops.act_prepare(qkv, dst_h16=(hi, lo))
fh, fl = hi.view(-1), lo.view(-1)
st = (t * 3 * c, 3 * d, 3 * c)
oh = torch.empty((b, t, c), dtype=torch.float16, device=dev)
ol = torch.empty_like(oh)
for _ in range(2):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.attention_h16((fh, fl), (fh[d:], fl[d:]), (fh[2 * d:], fl[2 * d:]), None, b, heads, t, t, d, d ** -0.5,
                      dict(q=st, k=st, v=st, o=(t * c, d, c)), o_h16=(oh, ol))
    e.record()
    torch.cuda.synchronize()
    print("launch", s.elapsed_time(e) * 1e3, "us (includes the profile read-back when TFMQ_ATTN_PROF is set)", flush=True)
