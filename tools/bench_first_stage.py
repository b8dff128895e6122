"""Time of the first-stage decode on one B200:
  python tools/bench_first_stage.py vq_f4 16     (LDM-4 CelebA-HQ, BASELINE configs[1]: 16 latents 3x64x64 -> 3x256x256)
  python tools/bench_first_stage.py kl_f8 1      (SD v1.4, configs[2] per-GPU share: 1 latent 4x64x64 -> 3x512x512)
Prints ms per decode (CUDA events over graph replays), launches, and the per-kernel-family times of one eager decode."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from helpers import synth  # noqa: E402
from tfmq_b200 import first_stage as FS  # noqa: E402
from tfmq_b200 import ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vq_f4"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = dict(vq_f4=FS.vq_f4_config, kl_f8=FS.kl_f8_config)[name]()
dev = torch.device("cuda:0")
m = FS.FirstStageModel(**cfg).eval()
synth.fill_state_dict(m, 7)
if cfg["n_embed"]:
    m.quantize.embedding.weight.data.copy_(synth.latents((cfg["n_embed"], cfg["embed_dim"]), 91))
m = m.to(dev)
z = synth.latents((nb, cfg["embed_dim"], 64, 64), 21).to(dev)
eng = m.engine(nb, 64, 64, dev)
for _ in range(3):
    img = m.decode_first_stage(z)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
s.record()
for _ in range(n):
    img = m.decode_first_stage(z)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
print(f"{name}: batch {nb}, latent {tuple(z.shape[1:])} -> image {tuple(img.shape[1:])}: {ms:.2f} ms per decode "
      f"({ms / nb:.2f} ms per image, {nb / ms * 1e3:.1f} images/s), {eng.launches_per_decode} launches; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

names = ["first_stage_input", "act_prepare", "conv_h16", "attention", "gn_stats_part", "conv_in", "conv_out"]
acc = collections.defaultdict(list)
orig = {k: getattr(ops, k) for k in names}


def wrap(k):
    def f(*a, **kw):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        r = orig[k](*a, **kw)
        e_.record()
        acc[k].append((s_, e_))
        return r
    return f


for k in names:
    setattr(ops, k, wrap(k))
eng.use_graph = False
eng.decode(z)
torch.cuda.synchronize()
acc.clear()
eng.decode(z)
torch.cuda.synchronize()
tot = 0.0
rows = []
for k, evs in acc.items():
    t = sum(a.elapsed_time(b) for a, b in evs)
    tot += t
    rows.append((t, k, len(evs)))
for t, k, c in sorted(rows, reverse=True):
    print(f"  {k:20s} n={c:3d} {t:8.3f} ms  {100 * t / tot:5.1f} %")
print(f"  eager total {tot:.2f} ms")
