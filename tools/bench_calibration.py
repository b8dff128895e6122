"""Wall-clock of the PTQ weight-reconstruction phase (BASELINE configs[3]: TIAR + AdaRound block reconstruction) on one
B200: every reconstruction unit of the LDM-4 UNet (22 QuantResBlocks + 3 upsample convs + the Temporal Information Block)
is run through `block_/layer_/tib_reconstruction` for K2 iterations; device-synchronised stamps at iterations K1 and K2 give the
steady-state time per iteration, the rest of the run is the input/output caching (+ the graph capture).  The reference runs 20 000 iterations per unit
(sample_diffusion_ldm.py:506-538), so  sum(per-iteration) x 20 000 + caching  is the projected W1 wall-clock.

  python tools/bench_calibration.py [n_calibration_samples=256]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from helpers import fp_model, synth  # noqa: E402
from tfmq_b200.quant.quant_block import BaseQuantBlock  # noqa: E402
from tfmq_b200.quant.quant_layer import QMODE, QuantLayer, Scaler  # noqa: E402
from tfmq_b200.quant.quant_model import QuantModel  # noqa: E402
from tfmq_b200.quant.reconstruction import block_reconstruction, layer_reconstruction, tib_reconstruction  # noqa: E402
from tfmq_b200.quant.reconstruction_util import RLOSS  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
# cali_model runs inside QuantModel.calibrating(), which switches TF32 off (fp32 parity with the reference's CPU graph); the
# units are called directly here, so set the same state.  TFMQ_BENCH_TF32=1 times torch's kernels with TF32 on instead.
_tf32 = os.environ.get("TFMQ_BENCH_TF32", "0") == "1"
torch.backends.cudnn.allow_tf32 = _tf32
torch.backends.cuda.matmul.allow_tf32 = _tf32
print(f"contractions: {'own tcgen05 kernels (quant/tc_autograd.py)' if os.environ.get('TFMQ_TC_RECON', '1') != '0' else 'torch kernels'}"
      f", TF32 {'on' if _tf32 else 'off'}")
K1, K2 = (int(v) for v in os.environ.get("TFMQ_BENCH_ITERS", "20,220").split(","))     # iterations of the two runs per unit
dev = torch.device("cuda:0")
wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
qnn = QuantModel(fp_model("ldm").to(dev), wq, aq, cali=True, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
qnn.eval()
g = torch.Generator().manual_seed(0)
cali = (synth.latents((n, 3, 64, 64), 71), torch.randint(0, 1000, (n,), generator=g).float())
qnn.set_quant_state(True, False)
with torch.no_grad():
    qnn(*(d[:8].to(dev) for d in cali))
qnn.disable_out_quantization()
kw = dict(batch_size=32, w=0.01, asym=True, warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)


def units(model, out):
    for name, m in model.named_children():
        if name == "tib":
            continue
        if name in ("time_embed", "temb"):
            out.append(("TIB", "tib", qnn.tib))
        elif isinstance(m, QuantLayer):
            if not m.ignore_recon:
                out.append((name, "layer", m))
        elif isinstance(m, BaseQuantBlock):
            if not m.ignore_recon:
                out.append((name, "block", m))
        else:
            units(m, out)
    return out


import tfmq_b200.quant.reconstruction as _R  # noqa: E402

_stamp = {}


def _hook(it):
    """device-synchronised wall-clock stamps at iterations K1 and K2 of a run: the steady state of the loop, without the input /
    output caching before it and the CUDA-graph capture of its first iteration"""
    if it in (K1, K2):
        torch.cuda.synchronize()
        _stamp[it] = time.time()


_R.ITER_HOOK = _hook


def run(kind, m, iters):
    torch.cuda.synchronize()
    t0 = time.time()
    if kind == "tib":
        tib_reconstruction(m, cali_data=cali, iters=iters, **kw)
    elif kind == "layer":
        layer_reconstruction(qnn, m, cali_data=cali, iters=iters, **kw)
    else:
        block_reconstruction(qnn, m, cali_data=cali, iters=iters, **kw)
    torch.cuda.synchronize()
    return time.time() - t0


names = {id(m): n_ for n_, m in qnn.model.named_modules()}
tot_iter = tot_cache = 0.0
todo = units(qnn.model, [])
print(f"{len(todo)} reconstruction units, {n} calibration samples, batch 32")
only = os.environ.get("TFMQ_BENCH_UNIT")        # e.g. output_blocks.9.0: time that unit alone and print its kernel profile
if only:
    todo = [u for u in todo if names.get(id(u[2]), "tib") == only]
    _, kind, m = todo[0]
    run(kind, m, 5)
    R = _R
    from torch.profiler import ProfilerActivity, profile
    loop = R._adaround_loop

    def profiled_loop(*a, **k):            # the iteration loop alone (not the input / output caching before it)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            r = loop(*a, **k)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
        return r
    R._adaround_loop = profiled_loop
    run(kind, m, 10)
    R._adaround_loop = loop
for _, kind, m in todo:
    has_layers = kind == "tib" or any(isinstance(x, QuantLayer) and not x.quant_emb for x in m.modules())
    if not has_layers:
        continue
    total = run(kind, m, K2)
    per = (_stamp[K2] - _stamp[K1]) / (K2 - K1)
    cache = max(total - K2 * per, 0.0)   # input / output caching (+ the graph capture)
    tot_iter += per
    tot_cache += cache
    print(f"  {names.get(id(m), 'tib'):28s} {kind:5s} {per * 1e3:7.2f} ms / iteration   caching {cache:5.2f} s", flush=True)
print(f"sum over units: {tot_iter * 1e3:.1f} ms per iteration round, caching {tot_cache:.1f} s at {n} samples")
print(f"projected weight-reconstruction wall-clock at 20000 iterations per unit: {tot_iter * 20000 / 3600:.2f} h "
      f"(+ caching {tot_cache * 1024 / n / 60:.1f} min at 1024 samples)")
