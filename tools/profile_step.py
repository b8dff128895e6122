"""One eager (un-graphed) denoising step of the bench workload between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/profile_step.py
and for   ncu --profile-from-start off --set full -k regex:igemm ...   captures of single kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

batch = int(os.environ.get("TFMQ_BATCH", bench.BATCH))
dev = torch.device("cuda:0")
qnn, eng, ts = bench.build_quantised(dev, batch)
eng.use_graph = False
x = torch.randn(batch, 3, 64, 64, device=dev)
eng.x_in.copy_(x)
for k in range(2):
    eng.step(k)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, launches/step =", eng.launches_per_step)
