"""int8 dense tensor peak by the protocol of MEASURED_PEAKS.json: 8192^3 s8 x s8 -> s32 (cuBLASLt through torch._int_mm),
best of 10 (burst) and back to back for 4 s (sustained); prints one JSON line."""
import json
import time

import torch

dev = torch.device("cuda:0")
n = 8192
a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
for _ in range(3):
    torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    torch._int_mm(a, b)
    e.record()
    torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
ops = 2 * n ** 3
burst = ops / (best * 1e-3) / 1e12
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
cnt = 0
s.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        torch._int_mm(a, b)
    cnt += 20
    torch.cuda.synchronize()
e.record()
torch.cuda.synchronize()
sustained = ops * cnt / (s.elapsed_time(e) * 1e-3) / 1e12
print(json.dumps({"int8_tops": burst, "int8_tops_sustained": sustained, "how": "torch._int_mm 8192^3 s8, best of 10 / 4 s back to back"}))
