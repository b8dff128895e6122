"""CPU baseline for tools/bench_calibration.py: the REFERENCE's own block_reconstruction (quant/reconstruction.py:86-209)
on one LDM-4 unit, on this container's host cores (no GPU here), timed at two iteration counts.
Run once in the build container:  python tests/golden/time_reference_recon.py [unit=input_blocks.4.0]
It reuses make_golden.py's environment shims (type-only imports, .cuda() -> identity)."""
import sys
import time

import torch

import make_golden as G  # noqa: F401  (installs the shims, imports the reference)
from quant.reconstruction import block_reconstruction
from quant.reconstruction_util import RLOSS

unit = sys.argv[1] if len(sys.argv) > 1 else "input_blocks.4.0"
qnn, _ = G.build_ref_qnn("ldm")
n = 32
g = torch.Generator().manual_seed(0)
cali = (G.synth.latents((n, 3, 64, 64), 71), torch.randint(0, 1000, (n,), generator=g).float())
qnn.set_quant_state(True, False)
with torch.no_grad():
    qnn(*(d[:8] for d in cali))
qnn.disable_out_quantization()
blk = dict(qnn.model.named_modules())[unit]
kw = dict(batch_size=32, w=0.01, asym=True, warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False, keep_gpu=True)
times = {}
block_reconstruction(qnn, blk, cali_data=cali, iters=1, **kw)        # warm-up: swaps the quantisers to AdaRound
for iters in (2, 10):
    t0 = time.time()
    block_reconstruction(qnn, blk, cali_data=cali, iters=iters, **kw)
    times[iters] = time.time() - t0
per = (times[10] - times[2]) / 8
print(f"reference block_reconstruction on {unit}, CPU {torch.get_num_threads()} threads, batch 32: "
      f"{per:.2f} s / iteration (runs: {times})")
