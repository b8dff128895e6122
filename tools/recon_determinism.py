"""How reproducible is a reconstruction run?  30 iterations of block_reconstruction on one CIFAR block, run twice eagerly, twice
through the captured iteration graph: loss-trace and alpha differences between the runs (split-K weight gradients are reduced by
TMA adds in arrival order, and Adam turns noise-level gradients into +-lr steps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import tfmq_b200.quant.reconstruction as R  # noqa: E402
from helpers import fp_model, synth  # noqa: E402
from tfmq_b200.quant.quant_layer import QMODE, Scaler  # noqa: E402
from tfmq_b200.quant.quant_model import QuantModel  # noqa: E402
from tfmq_b200.quant.reconstruction_util import RLOSS  # noqa: E402

dev = torch.device("cuda:0")
w_cali = (synth.latents((64, 3, 32, 32), 51), torch.randint(0, 1000, (64,), generator=torch.Generator().manual_seed(2)).float())


def run(graph):
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp_model("cifar").to(dev), wq, aq, cali=True, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*(d[:8].to(dev) for d in w_cali))
    qnn.disable_out_quantization()
    blk = qnn.model.down[1].block[0]
    if os.environ.get("TFMQ_KEEP_DROPOUT") is None:
        blk.dropout.p = 0.0      # the unit runs in train() mode (data_utill.py:72): with dropout, eager and replayed runs draw their
        #                          masks from different positions of the generator stream (both valid, not comparable bit for bit)
    R.RECON_GRAPH = graph
    R.LOSS_TRACE = []
    torch.manual_seed(0)
    R.block_reconstruction(qnn, blk, w_cali, batch_size=32, iters=30, w=0.01, opt_mode=RLOSS.MSE, asym=False, b_range=(20, 2),
                           warmup=0.2, multi_gpu=False)
    al = [m.wqtizer.alpha.detach().clone() for m in blk.modules() if hasattr(m, "wqtizer") and hasattr(m.wqtizer, "alpha")]
    return R.LOSS_TRACE, al


def cmp(name, a, b):
    (ta, aa), (tb, ab) = a, b
    dt = max(abs(x - y) for x, y in zip(ta, tb)) / max(abs(v) for v in tb)
    tot = sum(x.numel() for x in aa)
    for thr in (5e-4, 2.5e-3, 1e-2):
        print(f"{name}: loss trace max rel diff {dt:.2e}; alpha elements further apart than {thr:g}: "
              f"{sum(int(((x - y).abs() > thr).sum()) for x, y in zip(aa, ab))} of {tot}")


e1, e2, g1, g2 = run(False), run(False), run(True), run(True)
cmp("eager vs eager", e1, e2)
cmp("graph vs graph", g1, g2)
cmp("graph vs eager", g1, e1)
print("trace eager:", " ".join(f"{v:.4f}" for v in e1[0][:8]), "...", f"{e1[0][-1]:.4f}")
print("trace graph:", " ".join(f"{v:.4f}" for v in g1[0][:8]), "...", f"{g1[0][-1]:.4f}")
