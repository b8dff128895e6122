"""Trace the first-stage DecoderEngine programs WITHOUT a GPU: every C-ABI call is replaced by a recorder and tensor
allocations are redirected to the CPU, so the host logic (program order, shapes, the asserts of ops.py) is exercised here;
nothing is computed.  Used by tests/test_host_logic_cpu.py (in a subprocess: it monkey-patches torch).
Prints one line per decoder: <kind> <ops in the program> <launch count by entry point as JSON>.
With --emulate the calls are not only recorded: torch stand-ins with the kernels' semantics (tests/emulated_ops.py) execute the
program on the CPU and the image is compared with the oracle and the reference fixture; a fourth field is the max abs error."""
import contextlib
import json
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch  # noqa: E402


def _strip_device(f):
    def g(*a, **k):
        k.pop("device", None)
        return f(*a, **k)
    return g


torch.zeros, torch.empty = _strip_device(torch.zeros), _strip_device(torch.empty)
_orig_to = torch.Tensor.to


def _to(self, *a, **k):
    a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda"))
    if isinstance(k.get("device"), torch.device):
        k.pop("device")
    return _orig_to(self, *a, **k) if (a or k) else self


torch.Tensor.to = _to
torch.cuda.device = lambda d: contextlib.nullcontext()
from tfmq_b200 import ops  # noqa: E402

calls = []


class Recorder:
    launches = 0

    def call(self, name, *a):
        calls.append(name)


ops._ctx = lambda t: Recorder()
ops._stream = lambda: None
from helpers import first_stage_model  # noqa: E402
from tfmq_b200.first_stage import DecoderEngine  # noqa: E402

EMULATE = "--emulate" in sys.argv
if EMULATE:
    import emulated_ops
    from helpers import load_golden
    from oracle import first_stage_ref as FS
    emulated_ops.install(ops)
    golden = load_golden("first_stage.pt")

for kind in ("vq", "vq-attn", "kl"):
    m, cfg = first_stage_model(kind)
    eng = DecoderEngine(m, 2, 16, 16, device=torch.device("cuda"), use_graph=False)
    calls.clear()
    quant = cfg["n_embed"] is not None
    if not EMULATE:
        eng._run(quant, True)
        assert tuple(eng.image.shape) == (2, 3, 32, 32)
        print(kind, len(eng.ops), json.dumps(Counter(calls)))
        continue
    z = golden[kind]["z"]
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    errs = []
    for force in ((False, True) if quant else (False,)):
        eng.z_in.copy_(z)
        eng._run(quant and not force, True)
        ref = FS.decode_first_stage(z, sd, cfg["scale_factor"], quantize=quant, force_not_quantize=force)
        fix = golden[kind]["image_not_quantized" if force else "image"]
        errs += [(eng.image - ref).abs().max().item(), (eng.image - fix).abs().max().item()]
    print(kind, len(eng.ops), json.dumps({}), max(errs))
