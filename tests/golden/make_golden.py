"""Generate the golden fixtures by running the REFERENCE's own code (ModelTC/TFMQ-DM at
/root/reference) on CPU.  Run once in the build container:  python tests/golden/make_golden.py
The fixtures are small (.pt, outputs + quantiser parameters only; weights are re-created from the
seeded fill in synth.py) and are what tests/test_oracle_golden.py and the GPU parity tests check."""
import os
import sys
import tempfile
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFMQ_REFERENCE", "/root/reference")
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path[:0] = [REF, os.path.join(REF, "stable-diffusion"), HERE, os.path.join(HERE, ".."), ROOT,
                os.path.join(ROOT, "tfmq-dm_b200")]

# ---- environment shims (type-only imports and hard-coded .cuda()) -------------------------------
pl = types.ModuleType("pytorch_lightning")
pl.LightningModule = torch.nn.Module
pl.seed_everything = lambda s: torch.manual_seed(s)
plu = types.ModuleType("pytorch_lightning.utilities")
plud = types.ModuleType("pytorch_lightning.utilities.distributed")
plud.rank_zero_only = lambda f: f
sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": plu,
                    "pytorch_lightning.utilities.distributed": plud})
oc = types.ModuleType("omegaconf")          # openaimodel.py:509 imports ListConfig only to test `type(context_dim)`
ocl = types.ModuleType("omegaconf.listconfig")
ocl.ListConfig = type("ListConfig", (list,), {})
sys.modules.update({"omegaconf": oc, "omegaconf.listconfig": ocl})
ddpm_stub = types.ModuleType("ldm.models.diffusion.ddpm")
ddpm_stub.LatentDiffusion = torch.nn.Module
sys.modules["ldm.models.diffusion.ddpm"] = ddpm_stub
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
_orig_to = torch.Tensor.to


def _to(self, *a, **k):
    def is_cuda(x):
        return (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda")
    a = tuple("cpu" if is_cuda(x) else x for x in a)
    if is_cuda(k.get("device")):
        k["device"] = "cpu"
    return _orig_to(self, *a, **k)


torch.Tensor.to = _to
torch.set_flush_denormal(True)
torch.set_num_threads(8)

import synth  # noqa: E402
from quant.quant_layer import QMODE, QuantLayer, Scaler, UniformAffineQuantizer, minmax, mse  # noqa: E402
from quant.adaptive_rounding import AdaRoundQuantizer, RMODE  # noqa: E402
from quant.quant_model import QuantModel  # noqa: E402
from quant.calibration import load_cali_model  # noqa: E402
from quant import quant_block as rqb  # noqa: E402

SEED = 1234


def wq_aq(scaler=Scaler.MINMAX):
    wq = dict(bits=4, channel_wise=True, scaler=scaler)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    return wq, aq


# ---------------------------------------------------------------------------- G1/G2/G4 KATs
def quantizer_kats():
    out = {}
    g = torch.Generator().manual_seed(7)
    w = torch.randn(24, 16, 3, 3, generator=g) * 0.07
    x = torch.randn(4, 16, 12, 12, generator=g) * 1.7 + 0.2
    # G2 scalers
    d, z = minmax(x, False, 256, False)
    out["minmax_x"] = (d.clone(), torch.as_tensor(z).clone())
    d, z = mse(x, False, 256, False)
    out["mse_x"] = (d.clone(), torch.as_tensor(z).clone())
    q = UniformAffineQuantizer(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    wdq = q(w)
    out["w_minmax"] = (q.delta.clone(), q.zero_point.clone(), wdq.clone())
    q2 = UniformAffineQuantizer(bits=4, channel_wise=True, scaler=Scaler.MSE)
    wdq2 = q2(w)
    out["w_mse"] = (q2.delta.clone(), q2.zero_point.clone(), wdq2.clone())
    # G1 act quantiser incl. running-stat update
    qa = UniformAffineQuantizer(bits=8, channel_wise=False, scaler=Scaler.MSE, leaf_param=True)
    xdq = qa(x)
    out["x_mse_fq"] = (qa.delta.detach().clone(), torch.as_tensor(qa.zero_point).clone(), xdq.detach().clone())
    qa.running_stat = True
    trace = []
    for i in range(3):
        xi = torch.randn(4, 16, 12, 12, generator=g) * (1.0 + 0.5 * i)
        qa(xi)
        trace.append((xi, qa.x_min.clone(), qa.x_max.clone(), qa.delta.detach().clone(), qa.zero_point.clone()))
    out["running_stat"] = trace
    qs = UniformAffineQuantizer(bits=8, channel_wise=False, scaler=Scaler.MINMAX, always_zero=True)
    p = torch.softmax(torch.randn(2, 8, 8, generator=g), -1)
    out["softmax_always_zero"] = (p, qs(p).clone(), qs.delta.clone())
    # G4 AdaRound
    ar = AdaRoundQuantizer(q, w, RMODE.LEARNED_HARD_SIGMOID)
    alpha0 = ar.alpha.detach().clone()
    hard0 = ar(w).detach().clone()
    ar.alpha.data += torch.randn(w.shape, generator=g)
    hard1 = ar(w).detach().clone()
    ar.soft_tgt = True
    soft1 = ar(w).detach().clone()
    out["adaround"] = dict(alpha0=alpha0, hard0=hard0, alpha1=ar.alpha.detach().clone(), hard1=hard1, soft1=soft1)
    out["inputs"] = dict(w=w, x=x)
    # G3 QuantLayer KATs: conv3x3, conv1x1, linear
    layers = {}
    for name, mod, inp in (("conv3", torch.nn.Conv2d(16, 24, 3, padding=1), x),
                           ("conv1", torch.nn.Conv2d(16, 8, 1), x),
                           ("lin", torch.nn.Linear(32, 12), torch.randn(5, 32, generator=g))):
        synth.fill_state_dict(mod, 5)
        wq, aq = wq_aq()
        ql = QuantLayer(mod, wq, aq)
        ql.set_quant_state(True, True)
        y = ql(inp)
        layers[name] = dict(x=inp, y=y.detach().clone(), wd=ql.wqtizer.delta.clone(), wz=ql.wqtizer.zero_point.clone(),
                            ad=ql.aqtizer.delta.detach().clone(), az=torch.as_tensor(ql.aqtizer.zero_point).clone())
    out["quant_layer"] = layers
    torch.save(out, os.path.join(HERE, "kats.pt"))
    print("kats.pt written")


# ---------------------------------------------------------------------------- model goldens
def build_ref_qnn(kind):
    if kind == "cifar":
        from ddim.models.diffusion import Model
        sys.path.insert(0, os.path.join(HERE, "..", "..", "tfmq-dm_b200"))
        from tfmq_b200.host.ddim_unet import cifar10_config
        fp = Model(cifar10_config())
        x = synth.latents((1, 3, 32, 32), 11)
    elif kind == "sdmini":
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        from tfmq_b200.host.ldm_unet import sd_mini_config
        fp = UNetModel(**sd_mini_config())
        x = synth.latents((2, 4, 16, 16), 13)
    else:
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        sys.path.insert(0, os.path.join(HERE, "..", "..", "tfmq-dm_b200"))
        from tfmq_b200.host.ldm_unet import celebahq_ldm4_config
        fp = UNetModel(**celebahq_ldm4_config())
        x = synth.latents((1, 3, 64, 64), 12)
    fp.eval()
    synth.fill_state_dict(fp, SEED)
    wq, aq = wq_aq()
    qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    return qnn, x


def attach_alpha(qnn, init):
    """Run the reference's load_cali_model on a synthetic AdaRound checkpoint."""
    # first pass: initialise weight quantisers to learn delta (needed for the synthetic alpha)
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*init)
    weight = {}
    for name, m in qnn.model.named_modules():
        if isinstance(m, QuantLayer):
            weight[f"model.{name}.wqtizer.alpha"] = synth.synth_alpha(name, m.original_w, m.wqtizer.delta, SEED)
    for m in qnn.model.modules():           # undo the init so load_cali_model starts from scratch
        if isinstance(m, QuantLayer):
            m.wqtizer.init = False
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ckpt.pth")
        torch.save({"weight": weight}, path)
        with torch.no_grad():
            load_cali_model(qnn, init, use_aq=True, path=path)
    qnn.set_quant_state(True, True)


def reset_aq(qnn):
    for name, module in qnn.model.named_modules():
        if "aqtizer" in name and isinstance(module, UniformAffineQuantizer):
            if module.delta is not None:   # registered Parameters must be deleted first (calibration.py:115-121)
                del module.delta
                del module.zero_point
            module.delta = None
            module.zero_point = None
            module.init = False


def collect_aq(qnn):
    d = {}
    for name, module in qnn.model.named_modules():
        if "aqtizer" in name and isinstance(module, UniformAffineQuantizer) and module.delta is not None:
            d["model." + name + ".delta"] = module.delta.detach().clone().reshape(())
            d["model." + name + ".zero_point"] = torch.as_tensor(module.zero_point).detach().clone().reshape(())
    return d


def pack_act(act):
    """list of act_k dicts -> (sorted layer names, float tensor [steps, layers, 2]) to keep the fixture small."""
    names = sorted({k[len("model."):-len(".aqtizer.delta")] for k in act[0] if k.endswith(".aqtizer.delta")})
    tab = torch.zeros(len(act), len(names), 2)
    for k, a in enumerate(act):
        for i, n in enumerate(names):
            tab[k, i, 0] = a[f"model.{n}.aqtizer.delta"]
            tab[k, i, 1] = a[f"model.{n}.aqtizer.zero_point"]
    return names, tab


def promote_zero_points(qnn):
    for module in qnn.model.modules():
        if isinstance(module, UniformAffineQuantizer) and module.delta is not None and \
                not isinstance(module.zero_point, torch.nn.Parameter):
            module.zero_point = torch.nn.Parameter(torch.as_tensor(module.zero_point).float())


def alt_arithmetic():
    """The oracle with every conv / linear accumulated in float64 (oracle/quant_ref.py::float64_accumulation): measures how
    far fp re-association alone moves the reference's outputs (flip cascade)."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import quant_ref as Q
    return Q.float64_accumulation()


def oracle_alt_cifar(g):
    """The oracle's float64-accumulation evaluation of the CIFAR golden: eps at step 0 and the 50-step latent."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    sys.path.insert(0, os.path.join(HERE, "..", "..", "tfmq-dm_b200"))
    from helpers import CIFAR_CFG, fp_model, oracle_spec
    from oracle import unet_ref as U
    sd = fp_model("cifar", g["seed"]).state_dict()
    spec = oracle_spec(sd, g["seed"])
    names, tab = g["act_names"], g["act_table"]
    with torch.no_grad(), alt_arithmetic():
        x, t, _ = g["eps"][0]
        e = U.ddim_unet_forward(sd, CIFAR_CFG, x, t, spec, U.ActParams(names, tab[0]))
        fn = lambda xt, tt, k: U.ddim_unet_forward(sd, CIFAR_CFG, xt, tt, spec, U.ActParams(names, tab[k]))  # noqa
        xs, _ = U.generalized_steps(g["x_T"], g["seq"], fn, synth.ddim_betas(), eta=0.0)
    return e, xs[-1]


def oracle_alt_ldm(g):
    sys.path.insert(0, os.path.join(HERE, ".."))
    sys.path.insert(0, os.path.join(HERE, "..", "..", "tfmq-dm_b200"))
    from helpers import LDM4_CFG, fp_model, oracle_spec
    from oracle import unet_ref as U
    sd = fp_model("ldm", g["seed"]).state_dict()
    spec = oracle_spec(sd, g["seed"])
    with torch.no_grad(), alt_arithmetic():
        return U.ldm_unet_forward(sd, LDM4_CFG, g["x"], g["t"], spec, U.ActParams(g["act_names"], g["act_table"][0]))


def cifar_golden(steps=50):
    from ddim.functions.denoising import generalized_steps
    t0 = time.time()
    qnn, x = build_ref_qnn("cifar")
    seq = list(range(0, 1000, 1000 // steps))
    betas = synth.ddim_betas()
    attach_alpha(qnn, (x, torch.tensor([float(seq[-1])])))
    print("cifar: load_cali_model done", time.time() - t0)
    # pass 1: FSC tables -- at every step re-initialise the activation quantisers on the current latent
    act, xt = [], x
    seq_next = [-1] + seq[:-1]
    with torch.no_grad():
        for i, j in zip(reversed(seq), reversed(seq_next)):
            reset_aq(qnn)
            t = torch.ones(1) * i
            et = qnn(xt, t)
            act.append(collect_aq(qnn))
            at = (1 - betas).cumprod(0)[i]
            an = (1 - betas).cumprod(0)[j] if j >= 0 else torch.tensor(1.0)
            x0 = (xt - et * (1 - at).sqrt()) / at.sqrt()
            xt = an.sqrt() * x0 + (1 - an).sqrt() * et
    print("cifar: pass 1 done", time.time() - t0)
    promote_zero_points(qnn)
    ckpt = {f"act_{k}": a for k, a in enumerate(act)}
    # pass 2: the reference sampler with the FSC switch
    eps = {}
    hook_k = [0]
    orig_fwd = qnn.forward

    def rec_fwd(xx, tt=None, context=None):
        out = orig_fwd(xx, tt) if context is None else orig_fwd(xx, tt, context)
        if hook_k[0] in (0, 25, 49):
            eps[hook_k[0]] = (xx.clone(), tt.clone(), out.clone())
        hook_k[0] += 1
        return out
    qnn.forward = rec_fwd
    xs, x0_preds, _, _ = generalized_steps(x, seq, qnn, betas, eta=0.0, tot=1000 // steps, cali_ckpt=ckpt,
                                           t_max=steps - 1)
    qnn.forward = orig_fwd
    print("cifar: pass 2 done", time.time() - t0)
    names, tab = pack_act(act)
    g = dict(seed=SEED, seq=seq, x_T=x, act_names=names, act_table=tab, eps=eps)
    alt_eps0, alt_last = oracle_alt_cifar(g)
    print("cifar: float64-accumulation oracle: eps0 dev", (alt_eps0 - eps[0][2]).abs().max().item(),
          "final latent dev", (alt_last - xs[-1]).abs().max().item(), time.time() - t0)
    torch.save(dict(seed=SEED, steps=steps, seq=seq, x_T=x, act_names=names, act_table=tab, eps=eps,
                    alt_eps0=alt_eps0, alt_last=alt_last, xs_last=xs[-1], x_mid=xs[25],
                    x0_last=x0_preds[-1]), os.path.join(HERE, "cifar_w4a8.pt"))
    print("cifar_w4a8.pt written")


def ldm_golden():
    t0 = time.time()
    qnn, x = build_ref_qnn("ldm")
    t = torch.tensor([501.0])
    attach_alpha(qnn, (x, t))
    print("ldm: load_cali_model done", time.time() - t0)
    with torch.no_grad():
        reset_aq(qnn)
        e = qnn(x, t)
        act = collect_aq(qnn)
        e2 = qnn(x, t)     # second call with frozen parameters = what sampling sees
    print("ldm: forward done", time.time() - t0)
    names, tab = pack_act([act])
    alt = oracle_alt_ldm(dict(seed=SEED, x=x, t=t, act_names=names, act_table=tab))
    print("ldm: float64-accumulation oracle: eps dev", (alt - e2).abs().max().item(), time.time() - t0)
    torch.save(dict(seed=SEED, x=x, t=t, act_names=names, act_table=tab, eps=e2, eps_init=e, alt_eps=alt),
               os.path.join(HERE, "ldm4_w4a8.pt"))
    print("ldm4_w4a8.pt written")


def sdmini_golden():
    """SpatialTransformer UNet (structure of SD v1.4 at a size the CPU oracle runs in seconds): the reference's
    QuantModel with QuantBasicTransformerBlock / cross_attn_forward, batch 2, a 7-token context per sample."""
    t0 = time.time()
    qnn, x = build_ref_qnn("sdmini")
    t = torch.tensor([801.0, 801.0])
    ctx = synth.latents((2, 7, 96), 14)
    attach_alpha(qnn, (x, t, ctx))
    with torch.no_grad():
        reset_aq(qnn)
        e = qnn(x, t, ctx)
        act = collect_aq(qnn)
        e2 = qnn(x, t, ctx)
    # the block-level attention quantisers must be inert: record the evidence with the fixture
    inert = all(not getattr(m, "use_aq", False) for m in qnn.model.modules()
                if m.__class__.__name__ == "QuantBasicTransformerBlock")
    names, tab = pack_act([act])
    torch.save(dict(seed=SEED, x=x, t=t, context=ctx, act_names=names, act_table=tab, eps=e2, eps_init=e,
                    block_attention_quantisers_inert=inert),
               os.path.join(HERE, "sdmini_w4a8.pt"))
    print("sdmini_w4a8.pt written", len(names), "act-quantised layers; inert:", inert, time.time() - t0)


def full_size_golden(name):
    """BASELINE configs[2] / [4] at FULL size through the reference's own QuantModel (w4a8, synthetic AdaRound checkpoint,
    classifier-free-guidance batch of 2): SD v1.4 (configs/stable-diffusion/v1-inference.yaml:29-44, 77-token context of 768)
    and cin256-v2 (configs/latent-diffusion/cin256-v2.yaml:19-39, one class token of 512).  Inputs are seeded (the tests
    regenerate them); the fixture holds the activation-quantiser table and eps only."""
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from tfmq_b200.host import ldm_unet as H
    t0 = time.time()
    cfg = dict(sd_v14=H.sd_v14_config, cin256=H.cin256_config)[name]()
    tk = 77 if name == "sd_v14" else 1
    seed = 7
    fp = UNetModel(**cfg).eval()
    synth.fill_state_dict(fp, seed)
    wq, aq = wq_aq()
    qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    x = synth.latents((2, cfg["in_channels"], 64, 64), 21)
    t = torch.tensor([601.0, 601.0])
    ctx = synth.latents((2, tk, cfg["context_dim"]), 22)
    global SEED
    saved, SEED = SEED, seed
    try:
        attach_alpha(qnn, (x, t, ctx))
    finally:
        SEED = saved
    print(name, ": load_cali_model done", time.time() - t0)
    with torch.no_grad():
        reset_aq(qnn)
        e = qnn(x, t, ctx)
        act = collect_aq(qnn)
        t1 = time.time()
        e2 = qnn(x, t, ctx)
        print(name, ": one reference w4a8 forward at batch 2 on", torch.get_num_threads(), "threads:", time.time() - t1, "s")
    names, tab = pack_act([act])
    torch.save(dict(seed=seed, x_seed=21, ctx_seed=22, t=t, tokens=tk, act_names=names, act_table=tab, eps=e2),
               os.path.join(HERE, f"{name}_w4a8.pt"))
    print(f"{name}_w4a8.pt written", len(names), "act-quantised layers", time.time() - t0)


def recon_golden(iters=50):
    """the reference's own `tib_reconstruction` and `block_reconstruction` (quant/reconstruction.py:86-318) for
    `iters` iterations with a fixed torch seed (it drives `randperm`): the loss of every iteration and a strided sample of
    every alpha afterwards.  Sequence as in cali_model (quant/calibration.py:71-124): weight-quantiser initialisation,
    first / last layer exemptions, TIB first, then the blocks.  Units: CIFAR UNet -- TIB (DDIM flavour) and an AttnBlock;
    small SpatialTransformer UNet -- TIB (LDM flavour), a ResBlock and a BasicTransformerBlock.  (The reference leaves the
    model in train() mode during reconstruction, so CIFAR's ResnetBlocks run with dropout 0.1 drawn from the CPU
    generator: those are not reproducible on another device and are left out; the LDM configs have dropout 0.)"""
    from ddim.models.diffusion import Model
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from quant import reconstruction_util as RU
    from quant.reconstruction import block_reconstruction, tib_reconstruction
    from tfmq_b200.host.ddim_unet import cifar10_config
    from tfmq_b200.host.ldm_unet import sd_mini_config
    trace = []
    for cls in (RU.LossFunc, RU.LossFuncTimeEmbedding):
        orig = cls.__call__

        def rec_call(self, *a, _orig=orig, **k):
            out = _orig(self, *a, **k)
            trace.append(float(out.detach()))
            return out
        cls.__call__ = rec_call
    kw = dict(batch_size=8, iters=iters, w=0.01, opt_mode=RU.RLOSS.MSE, asym=True, b_range=(20, 2), warmup=0.2,
              multi_gpu=False)
    stride = 37
    out = dict(seed=SEED, iters=iters, n=32, kw={k: v for k, v in kw.items() if k != "opt_mode"}, stride=stride, models={})

    def sample(a):
        a = a.detach()
        return a.flatten()[::stride].clone(), float(a.double().sum()), float(a.double().abs().sum())

    for kind, blocks in (("cifar", ["down.1.attn.0"]),
                         ("sdmini", ["input_blocks.1.0", "input_blocks.1.1.transformer_blocks.0"])):
        if kind == "cifar":
            fp = Model(cifar10_config()).eval()
            cali = (synth.latents((32, 3, 32, 32), 71),
                    torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(72)).float())
        else:
            fp = UNetModel(**sd_mini_config()).eval()
            cali = (synth.latents((32, 4, 16, 16), 73),
                    torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(74)).float(),
                    synth.latents((32, 7, 96), 75))
        synth.fill_state_dict(fp, SEED)
        wq, aq = wq_aq()
        qnn = QuantModel(fp, wq, aq, cali=True, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
        qnn.eval()
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(*(d[:8] for d in cali))
        qnn.disable_out_quantization()
        rec = {}
        trace.clear()
        torch.manual_seed(0)
        tib_reconstruction(qnn.tib, cali_data=cali, **kw)
        rec["tib_loss"] = torch.tensor(trace)
        rec["tib_alpha"] = {name: sample(m.wqtizer.alpha) for name, m in qnn.model.named_modules()
                            if isinstance(m, QuantLayer) and isinstance(m.wqtizer, AdaRoundQuantizer)}
        rec["blocks"] = {}
        mods = dict(qnn.model.named_modules())
        for i, bn in enumerate(blocks):
            trace.clear()
            torch.manual_seed(1 + i)
            block_reconstruction(qnn, mods[bn], cali_data=cali, **kw)
            rec["blocks"][bn] = dict(loss=torch.tensor(trace),
                                     alpha={n: sample(m.wqtizer.alpha) for n, m in mods[bn].named_modules()
                                            if isinstance(m, QuantLayer) and m.quant_emb is False})
            print(kind, bn, "loss", rec["blocks"][bn]["loss"][:3].tolist(), "...", rec["blocks"][bn]["loss"][-2:].tolist(),
                  "layers", list(rec["blocks"][bn]["alpha"]))
        print(kind, "tib layers", len(rec["tib_alpha"]), "loss", rec["tib_loss"][:3].tolist(), "...", rec["tib_loss"][-2:].tolist())
        out["models"][kind] = rec
    torch.save(out, os.path.join(HERE, "recon_trace.pt"))
    print("recon_trace.pt written")


def inout_golden():
    """The reference's own `save_inout` (quant/data_utill.py:13-55, GetLayerInpOut :109-169) on the small SpatialTransformer
    UNet after weight-quantiser initialisation: cached inputs and FP outputs of a ResBlock (x, emb), a BasicTransformerBlock
    (x, context) and the TIB-facing time-embedding layer, symmetric (FP inputs) and asymmetric (inputs seen with the preceding
    layers weight-quantised), as strided samples + sums."""
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from quant.data_utill import save_inout
    from tfmq_b200.host.ldm_unet import sd_mini_config
    fp = UNetModel(**sd_mini_config()).eval()
    synth.fill_state_dict(fp, SEED)
    wq, aq = wq_aq()
    qnn = QuantModel(fp, wq, aq, cali=True, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    cali = (synth.latents((16, 4, 16, 16), 173),
            torch.randint(0, 1000, (16,), generator=torch.Generator().manual_seed(174)).float(), synth.latents((16, 7, 96), 175))
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*(d[:8] for d in cali))
    qnn.disable_out_quantization()
    stride = 53
    mods = dict(qnn.model.named_modules())

    def pack(t):
        t = t.detach().float()
        return tuple(t.shape), t.flatten()[::stride].clone(), float(t.double().sum()), float(t.double().abs().sum())
    out = dict(seed=SEED, stride=stride, units={})
    for name in ("input_blocks.1.0", "input_blocks.1.1.transformer_blocks.0", "middle_block.0", "output_blocks.0.0"):
        for asym in (False, True):
            ins, outs = save_inout(qnn, mods[name], cali, asym=asym, use_act=False, batch_size=8, keep_gpu=True)
            outs = outs if isinstance(outs, tuple) else (outs,)
            out["units"][(name, asym)] = dict(ins=[pack(t) for t in ins], outs=[pack(t) for t in outs])
            print(name, "asym" if asym else "sym", [tuple(t.shape) for t in ins], "->", [tuple(t.shape) for t in outs])
    torch.save(out, os.path.join(HERE, "inout_sdmini.pt"))
    print("inout_sdmini.pt written")


def unet_keys():
    """state_dict keys / shapes of the REFERENCE UNetModel for every supported LDM config (meta device: no weights are
    materialised), and the module list of the reference QuantModel on the small transformer UNet."""
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from tfmq_b200.host import ldm_unet as H
    out = {}
    for name, cfg in (("ldm4", H.celebahq_ldm4_config()), ("sd_mini", H.sd_mini_config()), ("sd_v14", H.sd_v14_config()),
                      ("cin256", H.cin256_config())):
        with torch.device("meta"):
            m = UNetModel(**cfg)
        out[name] = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    qnn, _ = build_ref_qnn("sdmini")
    out["sd_mini_quant_modules"] = [(n, type(m).__name__) for n, m in qnn.named_modules()]
    torch.save(out, os.path.join(HERE, "unet_keys.pt"))
    print("unet_keys.pt written", {k: len(v) for k, v in out.items()})


def stub_eps(x, t, c=None):
    """A UNet stand-in made of exactly-rounded fp32 ops only (roll, mul, sub, add): the same bits on CPU and GPU, so a
    sampler trajectory pins the UPDATE RULE, not a network."""
    e = 0.3 * x.roll(1, -1) - (0.2 * x) * (t.float() / 1000.0)[:, None, None, None]
    if c is not None:
        e = e + 0.05 * c.reshape(c.shape[0], -1)[:, :1, None, None]
    return e


def plms_golden():
    """Trajectories of the reference's PLMSSampler / DDIMSampler (ldm/models/diffusion/{plms,ddim}.py) driven by `stub_eps`
    through a minimal LatentDiffusion stand-in (betas / alphas_cumprod buffers + apply_model), incl. guidance and the
    `untill_fake_t` early stop."""
    import numpy as np
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.plms import PLMSSampler
    from ldm.modules.diffusionmodules.util import make_beta_schedule

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
            ac = np.cumprod(1.0 - betas, axis=0)
            f = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
            self.num_timesteps = 1000
            self.register_buffer("betas", f(betas))
            self.register_buffer("alphas_cumprod", f(ac))
            self.register_buffer("alphas_cumprod_prev", f(np.append(1.0, ac[:-1])))
            self.register_buffer("sqrt_alphas_cumprod", f(np.sqrt(ac)))
            self.register_buffer("sqrt_one_minus_alphas_cumprod", f(np.sqrt(1.0 - ac)))
            self.register_buffer("log_one_minus_alphas_cumprod", f(np.log(1.0 - ac)))
            self.register_buffer("sqrt_recip_alphas_cumprod", f(np.sqrt(1.0 / ac)))
            self.register_buffer("sqrt_recipm1_alphas_cumprod", f(np.sqrt(1.0 / ac - 1)))
            self.device = torch.device("cpu")

        def apply_model(self, x, t, c):
            return stub_eps(x, t, c)

    m = Stub()
    x_T = synth.latents((2, 4, 8, 8), 81)
    c, uc = synth.latents((2, 3, 5), 82), synth.latents((2, 3, 5), 83)
    out = dict(x_T=x_T, c=c, uc=uc, linear_start=0.00085, linear_end=0.012)
    for name, cls in (("plms", PLMSSampler), ("ddim", DDIMSampler)):
        s10, _ = cls(m).sample(S=10, batch_size=2, shape=[4, 8, 8], eta=0.0, x_T=x_T.clone(), verbose=False)
        s10c, _ = cls(m).sample(S=10, batch_size=2, shape=[4, 8, 8], eta=0.0, x_T=x_T.clone(), verbose=False, conditioning=c,
                               unconditional_conditioning=uc, unconditional_guidance_scale=3.0)
        s4, _ = cls(m).sample(S=10, batch_size=2, shape=[4, 8, 8], eta=0.0, x_T=x_T.clone(), verbose=False, untill_fake_t=5)
        out[name] = dict(full=s10, guided=s10c, stop5=s4)
    # stochastic DDIM (eta = 1, the README's lsun / ffhq commands): sigma_t * noise_like(...) per step from the global CPU
    # generator, seeded here; the test feeds the product the same stream
    out["eta_seed"] = 77
    torch.manual_seed(out["eta_seed"])
    e1, _ = DDIMSampler(m).sample(S=10, batch_size=2, shape=[4, 8, 8], eta=1.0, x_T=x_T.clone(), verbose=False)
    torch.manual_seed(out["eta_seed"])
    e1c, _ = DDIMSampler(m).sample(S=10, batch_size=2, shape=[4, 8, 8], eta=1.0, x_T=x_T.clone(), verbose=False, conditioning=c,
                                   unconditional_conditioning=uc, unconditional_guidance_scale=3.0)
    out["ddim"]["eta1"], out["ddim"]["eta1_guided"] = e1, e1c
    torch.save(out, os.path.join(HERE, "samplers_stub.pt"))
    print("samplers_stub.pt written", {k: (v["full"].abs().max().item() if isinstance(v, dict) else None) for k, v in out.items()
                                        if isinstance(v, dict)})


def first_stage_golden():
    """First-stage decode through the reference's own `VQModelInterface.decode` / `AutoencoderKL.decode`
    (ldm/models/autoencoder.py:274-283,330-333 -> Decoder, ldm/modules/diffusionmodules/model.py:462-568) on small seeded
    models.  `ldm.models.autoencoder` imports taming-transformers' VectorQuantizer2, which the reference does not vendor:
    a stand-in module with its published forward (the same restatement as oracle/first_stage_ref.py::vq_lookup) is
    registered, so the lookup itself is NOT pinned by this fixture -- post_quant_conv and the decoder are."""
    from tfmq_b200.first_stage import first_stage_mini_config

    class VectorQuantizer2(torch.nn.Module):
        def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
            super().__init__()
            self.n_e, self.e_dim = n_e, e_dim
            self.embedding = torch.nn.Embedding(n_e, e_dim)
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

        def forward(self, z):
            z = z.permute(0, 2, 3, 1).contiguous()
            zf = z.view(-1, self.e_dim)
            d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(self.embedding.weight ** 2, dim=1) - 2 * \
                torch.einsum('bd,dn->bn', zf, self.embedding.weight.t())
            idx = torch.argmin(d, dim=1)
            z_q = self.embedding(idx).view(z.shape)
            z_q = z + (z_q - z).detach()
            return z_q.permute(0, 3, 1, 2).contiguous(), None, (None, None, idx)

    for name in ("taming", "taming.modules", "taming.modules.vqvae"):
        sys.modules[name] = types.ModuleType(name)
    tq = types.ModuleType("taming.modules.vqvae.quantize")
    tq.VectorQuantizer2 = VectorQuantizer2
    sys.modules["taming.modules.vqvae.quantize"] = tq
    from ldm.models.autoencoder import AutoencoderKL, VQModelInterface

    t0 = time.time()
    out = {}
    for kind in ("vq", "vq-attn", "kl"):
        cfg = first_stage_mini_config(kind)
        loss = {"target": "torch.nn.Identity"}
        if cfg["n_embed"] is not None:
            m = VQModelInterface(embed_dim=cfg["embed_dim"], n_embed=cfg["n_embed"], ddconfig=dict(cfg["ddconfig"]),
                                 lossconfig=loss)
        else:
            m = AutoencoderKL(ddconfig=dict(cfg["ddconfig"]), lossconfig=loss, embed_dim=cfg["embed_dim"])
        m.eval()
        synth.fill_state_dict(m, SEED)
        if cfg["n_embed"] is not None:       # codes of the latent's own scale, so that the lookup moves the latent a little
            m.quantize.embedding.weight.data.copy_(synth.latents((cfg["n_embed"], cfg["embed_dim"]), 91))
        z = synth.latents((2, cfg["embed_dim"], 16, 16), 92)
        zs = 1. / cfg["scale_factor"] * z                       # ddpm.py:713
        with torch.no_grad():
            rec = dict(z=z, image=m.decode(zs))
            if cfg["n_embed"] is not None:
                rec["image_not_quantized"] = m.decode(zs, force_not_quantize=True)
        rec["decoder_keys"] = [(k, tuple(v.shape)) for k, v in m.state_dict().items()
                               if k.startswith(("decoder.", "post_quant_conv.", "quantize."))]
        out[kind] = rec
    torch.save(out, os.path.join(HERE, "first_stage.pt"))
    print("first_stage.pt written", {k: tuple(v["image"].shape) for k, v in out.items()}, time.time() - t0)


def runner_golden():
    """The reference's DDIM runner (ddim/runners/diffusion.py): beta schedules, logvar, and `Diffusion.sample_image` driven by
    `stub_eps` on the uniform and quadratic timestep sequences, with and without the `untill_fake_t` early stop.
    (`ddim.datasets` imports lmdb for a dataset class that is never touched: an empty stand-in module is registered.)"""
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    from ddim.runners.diffusion import Diffusion, get_beta_schedule
    NS = types.SimpleNamespace
    out = {"schedules": {k: torch.from_numpy(get_beta_schedule(k, beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000))
                         for k in ("linear", "quad", "const", "jsd", "sigmoid")}}
    x = synth.latents((2, 3, 8, 8), 61)
    for var in ("fixedlarge", "fixedsmall"):
        cfg = NS(model=NS(var_type=var), diffusion=NS(beta_schedule="linear", beta_start=1e-4, beta_end=0.02,
                                                       num_diffusion_timesteps=1000))
        for skip in ("uniform", "quad"):
            r = Diffusion(NS(skip_type=skip, timesteps=10, sample_type="generalized", eta=0.0), cfg, device=torch.device("cpu"))
            full, _, _ = r.sample_image(x.clone(), stub_eps)
            _, x_t, t_t = r.sample_image(x.clone(), stub_eps, untill_fake_t=4)
            out[(var, skip)] = dict(betas=r.betas.clone(), logvar=r.logvar.clone(), full=full, x_t=x_t, t_t=t_t)
    out["x"] = x
    torch.save(out, os.path.join(HERE, "runner_ddim.pt"))
    print("runner_ddim.pt written", {k: (v["full"].abs().max().item(), v["t_t"].tolist()) for k, v in out.items()
                                      if isinstance(k, tuple)})


def cali_schema():
    """G9: run the reference's cali_model on a tiny synthetic set and record the checkpoint's key set and
    shapes (the on-disk format the drop-in must read and write)."""
    from ddim.models.diffusion import Model
    from quant.calibration import cali_model
    from quant.reconstruction_util import RLOSS
    from tfmq_b200.host.ddim_unet import cifar10_config
    t0 = time.time()
    fp = Model(cifar10_config())
    fp.eval()
    synth.fill_state_dict(fp, SEED)
    wq, aq = wq_aq()
    qnn = QuantModel(fp, wq, aq, cali=True, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    w_cali = (synth.latents((8, 3, 32, 32), 31), torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(1)).float())
    a_cali = (synth.latents((32, 3, 32, 32), 32), torch.cat([torch.full((16,), 980.0), torch.full((16,), 960.0)]))
    kwargs = dict(iters=2, batch_size=4, w=0.01, asym=True, warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.pth")
        cali_model(qnn, w_cali, a_cali, use_aq=True, path=path, running_stat=True, interval=16, **kwargs)
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    schema = {k: {kk: tuple(v.shape) for kk, v in d.items()} for k, d in ckpt.items()}
    torch.save(schema, os.path.join(HERE, "cali_schema.pt"))
    print("cali_schema.pt written", {k: len(v) for k, v in schema.items()}, time.time() - t0)


if __name__ == "__main__":
    what = sys.argv[1:] or ["kats", "cifar", "ldm"]
    if "kats" in what:
        quantizer_kats()
    if "cifar" in what:
        cifar_golden()
    if "ldm" in what:
        ldm_golden()
    if "runner" in what:
        runner_golden()
    if "first_stage" in what:
        first_stage_golden()
    if "schema" in what:
        cali_schema()
    if "sdmini" in what:
        sdmini_golden()
    if "keys" in what:
        unet_keys()
    if "plms" in what:
        plms_golden()
    if "recon" in what:
        recon_golden()
    if "inout" in what:
        inout_golden()
    for full in ("sd_v14", "cin256"):
        if full in what:
            full_size_golden(full)
