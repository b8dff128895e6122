"""GPU: the PTQ calibration passes through the drop-in API (cali_model -> checkpoint -> load_cali_model ->
step engine), against the reference's checkpoint schema (tests/golden/cali_schema.pt, produced by the
reference's own cali_model on the same tiny synthetic set)."""
import os
import tempfile

import pytest
import torch

from helpers import fp_model, load_golden, synth

pytestmark = pytest.mark.gpu


def _data():
    w_cali = (synth.latents((8, 3, 32, 32), 31),
              torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(1)).float())
    a_cali = (synth.latents((32, 3, 32, 32), 32), torch.cat([torch.full((16,), 980.0), torch.full((16,), 960.0)]))
    return w_cali, a_cali


def _qnn(dev, cali):
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    q = QuantModel(fp_model("cifar").to(dev), wq, aq, cali=cali, softmax_a_bit=8,
                   aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    q.eval()
    return q


def test_cali_model_checkpoint_schema_and_reload(dev):
    from tfmq_b200.quant.calibration import act_tables_from_ckpt, cali_model, load_cali_model
    from tfmq_b200.quant.reconstruction_util import RLOSS
    from tfmq_b200.samplers import generalized_steps
    schema = load_golden("cali_schema.pt")
    w_cali, a_cali = _data()
    qnn = _qnn(dev, cali=True)
    torch.manual_seed(0)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.pth")
        ckpt = cali_model(qnn, w_cali, a_cali, use_aq=True, path=path, running_stat=True, interval=16, iters=2,
                          batch_size=4, w=0.01, asym=True, warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
        on_disk = torch.load(path, map_location="cpu", weights_only=False)
    mine = {k: {kk: tuple(v.shape) for kk, v in d.items()} for k, d in on_disk.items()}
    assert set(mine) == set(schema)
    for part in schema:
        missing = set(schema[part]) - set(mine[part])
        extra = set(mine[part]) - set(schema[part])
        assert not missing and not extra, (part, sorted(missing)[:5], sorted(extra)[:5])
        for k, shp in schema[part].items():
            assert mine[part][k] == shp, (part, k, mine[part][k], shp)
    # the checkpoint drives sampling through a freshly built model: AdaRound detected by 'alpha' keys
    q2 = _qnn(dev, cali=False)
    x = synth.latents((2, 3, 32, 32), 40)
    load_cali_model(q2, (x, torch.full((2,), 980.0)), use_aq=True, ckpt=ckpt)
    from tfmq_b200.quant.adaptive_rounding import AdaRoundQuantizer
    assert isinstance(q2.model.down[0].block[0].conv2.wqtizer, AdaRoundQuantizer)
    assert torch.equal(q2.model.down[0].block[0].conv2.wqtizer.alpha.detach().cpu(),
                       ckpt["weight"]["model.down.0.block.0.conv2.wqtizer.alpha"])
    assert len(act_tables_from_ckpt(ckpt)) == 2
    seq = [960, 980]
    xs, x0, _, _ = generalized_steps(x.to(dev), seq, q2, synth.ddim_betas().to(dev), eta=0.0, tot=20, cali_ckpt=ckpt,
                                     t_max=1)
    assert torch.isfinite(xs[-1]).all() and xs[-1].shape == x.shape


def test_block_reconstruction_reduces_the_loss(dev):
    """AdaRound on one QuantResnetBlock with the fused kernels: the reconstruction error of the hard-rounded
    block after optimisation is lower than with round-to-nearest."""
    from tfmq_b200.quant.data_utill import save_inout
    from tfmq_b200.quant.quant_layer import lp_loss
    from tfmq_b200.quant.reconstruction import block_reconstruction
    from tfmq_b200.quant.reconstruction_util import RLOSS
    qnn = _qnn(dev, cali=True)
    w_cali = (synth.latents((64, 3, 32, 32), 51),
              torch.randint(0, 1000, (64,), generator=torch.Generator().manual_seed(2)).float())
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*(d[:8].to(dev) for d in w_cali))
    qnn.disable_out_quantization()
    blk = qnn.model.down[1].block[0]
    ins, outs = save_inout(qnn, blk, w_cali, asym=False, use_act=False, batch_size=32)

    def err():
        qnn.set_quant_state(False, False)
        blk.set_quant_state(True, False)
        blk.eval()
        with torch.no_grad():
            return lp_loss(blk(*ins), outs).item()
    before = err()
    torch.manual_seed(0)
    block_reconstruction(qnn, blk, w_cali, batch_size=32, iters=300, w=0.01, opt_mode=RLOSS.MSE, asym=False,
                         b_range=(20, 2), warmup=0.2, multi_gpu=False)
    after = err()
    print(f"block reconstruction: lp_loss nearest {before:.5f} -> AdaRound {after:.5f}")
    assert after < before


def test_cali_model_spatial_transformer_unet(dev):
    """cali_model on a conditional UNet (SD / cin256 structure): calibration data are (x, t, context) triples
    (quant/calibration.py:62-67, sample_diffusion_ldm.py:486-538), QuantBasicTransformerBlock units go through
    block_reconstruction, and the checkpoint drives guided sampling through a freshly built model."""
    from tfmq_b200.quant.calibration import act_tables_from_ckpt, cali_model, load_cali_model
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    from tfmq_b200.quant.reconstruction_util import RLOSS
    from tfmq_b200.samplers import DDIMSampler

    def build(cali):
        wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
        aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
        q = QuantModel(fp_model("sdmini").to(dev), wq, aq, cali=cali, softmax_a_bit=8,
                       aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
        q.eval()
        return q

    g = torch.Generator().manual_seed(3)
    w_cali = (synth.latents((8, 4, 16, 16), 61), torch.randint(0, 1000, (8,), generator=g).float(),
              synth.latents((8, 7, 96), 62))
    a_cali = (synth.latents((16, 4, 16, 16), 63), torch.cat([torch.full((8,), 751.0), torch.full((8,), 501.0)]),
              synth.latents((16, 7, 96), 64))
    qnn = build(True)
    torch.manual_seed(0)
    ckpt = cali_model(qnn, w_cali, a_cali, use_aq=True, path=None, running_stat=True, interval=8, iters=2, batch_size=4,
                      w=0.01, asym=True, warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
    alphas = [k for k in ckpt["weight"] if k.endswith("wqtizer.alpha")]
    assert any("transformer_blocks.0.attn2.to_k" in k for k in alphas) and any("ff.net.0.proj" in k for k in alphas)
    assert len(act_tables_from_ckpt(ckpt)) == 2
    q2 = build(False)
    x = synth.latents((2, 4, 16, 16), 65)
    c, uc = synth.latents((2, 7, 96), 66), synth.latents((2, 7, 96), 67)
    load_cali_model(q2, (x, torch.full((2,), 751.0), c), use_aq=True, ckpt=ckpt)
    out, _ = DDIMSampler(q2, ckpt=ckpt).sample(2, 2, (4, 16, 16), conditioning=c, unconditional_conditioning=uc,
                                              unconditional_guidance_scale=3.0, x_T=x)
    assert torch.isfinite(out).all() and out.shape == x.shape


@pytest.mark.parametrize("tc", [True, False], ids=["own-kernels", "torch-kernels"])
@pytest.mark.parametrize("kind", ["cifar", "sdmini"])
def test_reconstruction_trace_matches_the_reference(dev, kind, tc, monkeypatch):
    """50 iterations of `tib_reconstruction` and `block_reconstruction` with a fixed torch seed against the trace of
    the reference's own functions (quant/reconstruction.py:86-318, run on the CPU by tests/golden/make_golden.py::recon_golden):
    the loss of every iteration and the learned alpha of every layer.  Same sequence as cali_model: weight-quantiser
    initialisation, exemptions, TIB, then the blocks (whose asymmetric input cache sees the reconstructed TIB).
    cifar: DDIM-flavour TIB + an AttnBlock; sdmini: LDM-flavour TIB + a ResBlock + a BasicTransformerBlock."""
    import tfmq_b200.quant.quant_layer as QL
    import tfmq_b200.quant.reconstruction as R
    from tfmq_b200.quant.adaptive_rounding import AdaRoundQuantizer
    # tc: the unit's contractions (forward, dgrad, wgrad) on this library's tcgen05 kernels (quant/tc_autograd.py, the
    # default) or on torch's fp32 kernels
    monkeypatch.setattr(QL, "TC_RECONSTRUCTION", tc)
    # Adam divides by sqrt(v): where a gradient sits at the fp32 noise floor its sign, and so the direction of the 1e-3 step, is
    # chance.  torch's fp32 kernels happen to round like the reference's CPU kernels (no element moves); the tcgen05 path (fp16
    # split, 3 products) rounds differently, and up to ~1 % of a layer's elements end one or two steps apart.  Gates: the loss
    # trace, a bound on the largest deviation in steps, the fraction of elements beyond one step, and the sum of alpha.
    # (the loss responds to those elements: its trace follows the reference to 3e-3 with the own kernels, to 5e-6 with torch's)
    # measured with the own kernels: <= 1.1 % of a conv layer, 6-14 % of the self-attention output projection of the transformer
    # block (during the 10 warm-up iterations there is no rounding regulariser, and the reconstruction gradient of that layer
    # is mostly at the noise floor) beyond one step, never more than 6 of the 50 steps
    far_frac, far_max, trace_tol = (0.20, 1e-2, 5e-3) if tc else (2e-3, 5e-3, 1e-4)
    from tfmq_b200.quant.quant_layer import QMODE, QuantLayer, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    from tfmq_b200.quant.reconstruction_util import RLOSS
    g = load_golden("recon_trace.pt")
    rec = g["models"][kind]
    torch.backends.cudnn.allow_tf32 = False          # the unit's forward / backward through torch stays fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp_model(kind, g["seed"]).to(dev), wq, aq, cali=True, softmax_a_bit=8,
                     aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).eval()
    if kind == "cifar":
        cali = (synth.latents((32, 3, 32, 32), 71),
                torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(72)).float())
    else:
        cali = (synth.latents((32, 4, 16, 16), 73),
                torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(74)).float(),
                synth.latents((32, 7, 96), 75))
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*(d[:8].to(dev) for d in cali))
    qnn.disable_out_quantization()
    kw = dict(g["kw"], opt_mode=RLOSS.MSE)

    def compare(tag, trace, want_trace, alphas, want_alphas):
        trace = torch.tensor(trace)
        assert trace.shape == want_trace.shape
        rel = ((trace - want_trace).abs() / want_trace.abs().clamp_min(1e-6)).max().item()
        worst, worst_frac = 0.0, 0.0
        for name, (sample, s1, s2) in want_alphas.items():
            a = alphas[name].detach().float().cpu()
            d = (a.flatten()[::g["stride"]] - sample).abs()
            frac_far = (d > 1e-3).float().mean().item()
            worst, worst_frac = max(worst, d.max().item()), max(worst_frac, frac_far)
            assert frac_far < far_frac and d.max().item() < far_max, (tag, name, frac_far, d.max().item())
            assert abs(float(a.double().sum()) - s1) <= 5e-4 * max(1.0, s2), (tag, name)
        print(f"[{kind} {'own' if tc else 'torch'} kernels, {tag}] {len(trace)} iterations: loss trace max relative deviation {rel:.3e} (first {want_trace[0]:.5f}, "
              f"last {want_trace[-1]:.2f}); alpha of {len(want_alphas)} layers: worst element deviation {worst:.3e}, at most "
              f"{worst_frac:.1e} of a layer beyond 1e-3")
        assert rel < trace_tol, (tag, rel)

    R.LOSS_TRACE = []
    try:
        torch.manual_seed(0)
        R.tib_reconstruction(qnn.tib, cali_data=cali, **kw)
        tib_trace, R.LOSS_TRACE = R.LOSS_TRACE, []
        mods = dict(qnn.model.named_modules())
        alphas = {nm: m.wqtizer.alpha for nm, m in mods.items()
                  if isinstance(m, QuantLayer) and isinstance(m.wqtizer, AdaRoundQuantizer)}
        assert sorted(alphas) == sorted(rec["tib_alpha"])
        compare("tib_reconstruction", tib_trace, rec["tib_loss"], alphas, rec["tib_alpha"])
        for i, (bn, want) in enumerate(rec["blocks"].items()):
            R.LOSS_TRACE = []
            torch.manual_seed(1 + i)
            R.block_reconstruction(qnn, mods[bn], cali_data=cali, **kw)
            blk_alphas = {nm: m.wqtizer.alpha for nm, m in mods[bn].named_modules()
                          if isinstance(m, QuantLayer) and not m.quant_emb}
            assert sorted(blk_alphas) == sorted(want["alpha"])
            compare("block_reconstruction " + bn, R.LOSS_TRACE, want["loss"], blk_alphas, want["alpha"])
    finally:
        R.LOSS_TRACE = None


@pytest.mark.parametrize("xs,ws,kw,gscale", [
    ((8, 32, 16, 16), (64, 32, 3, 3), dict(stride=(1, 1), padding=(1, 1)), 1e-4),
    ((4, 224, 32, 32), (224, 224, 3, 3), dict(stride=(1, 1), padding=(1, 1)), 1e-6),
    ((8, 64, 16, 16), (128, 64, 1, 1), dict(stride=(1, 1), padding=(0, 0)), 1e-3),
    ((8, 256, 96), (384, 96), {}, 1e-5),
    ((32, 128), (512, 128), {}, 1.0),
], ids=["conv3x3-32-64", "conv3x3-224-224", "conv1x1-64-128", "linear-tokens", "linear-tib"])
def test_tc_autograd_forward_dgrad_wgrad_against_float64(dev, xs, ws, kw, gscale):
    """The three contractions of a reconstructed layer on the library's tcgen05 kernel (quant/tc_autograd.py: forward, dgrad on
    the flipped / transposed weights, wgrad as a pixel-reduction GEMM) against torch in float64 -- what `out_quant = block(...)`
    and `err.backward()` (quant/reconstruction.py:182-198) compute for a QuantLayer.  Gradients as small as 1e-6 (the scale of
    back-propagated reconstruction errors) must keep fp32-class accuracy: tolerance 3e-5 of the tensor's largest magnitude
    (measured 1e-7 ... 1e-5, the same band as torch's own fp32 kernels with TF32 off; tools/precision_tc_autograd.py)."""
    import torch.nn.functional as F
    from tfmq_b200.quant.tc_autograd import tc_conv
    g = torch.Generator().manual_seed(1)
    x = torch.randn(xs, generator=g).to(dev)
    w = (torch.randn(ws, generator=g) * 0.05).to(dev)
    b = torch.randn(ws[0], generator=g).to(dev)
    xx, ww, bb = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y64 = F.conv2d(xx, ww, bb, **kw) if len(ws) == 4 else F.linear(xx, ww, bb)
    gy = (torch.randn(y64.shape, generator=g) * gscale).to(dev)
    want = (y64.detach(),) + torch.autograd.grad(y64, (xx, ww, bb), gy.double())
    x1, w1, b1 = (t.detach().requires_grad_(True) for t in (x, w, b))
    y = tc_conv(x1, w1, b1, kw)
    assert y is not None
    got = (y.detach(),) + torch.autograd.grad(y, (x1, w1, b1), gy)
    for name, a, r in zip(("y", "dx", "dW", "db"), got, want):
        err = ((a.double() - r).abs().max() / r.abs().max()).item()
        assert err < 3e-5, (name, err)


def test_tc_autograd_declines_unsupported_shapes(dev):
    """Stride-2 down-sampling convs, channel counts that are not a multiple of 16 and non power-of-two maps stay on torch's op:
    tc_conv returns None and QuantLayer.forward falls through."""
    from tfmq_b200.quant.tc_autograd import tc_conv
    x = torch.randn(2, 32, 16, 16, device=dev)
    assert tc_conv(x, torch.randn(32, 32, 3, 3, device=dev), None, dict(stride=(2, 2), padding=(1, 1))) is None
    assert tc_conv(torch.randn(2, 3, 16, 16, device=dev), torch.randn(32, 3, 3, 3, device=dev), None,
                   dict(stride=(1, 1), padding=(1, 1))) is None
    assert tc_conv(torch.randn(2, 32, 12, 12, device=dev), torch.randn(32, 32, 3, 3, device=dev), None,
                   dict(stride=(1, 1), padding=(1, 1))) is None


def test_plain_convs_of_a_unit_run_on_the_own_kernel_inside_the_loop(dev):
    """A ResBlock's skip_connection is a plain nn.Conv2d (never a QuantLayer): inside the reconstruction loop its forward goes
    through the fp32-accurate tcgen05 kernel (forward only), outside it is torch's op again -- also after an exception."""
    from tfmq_b200.quant.reconstruction import _plain_convs_on_own_kernels
    torch.manual_seed(0)
    unit = torch.nn.Sequential(torch.nn.Conv2d(64, 32, 1), torch.nn.Conv2d(32, 32, 3, stride=2, padding=1)).to(dev)
    x = torch.randn(4, 64, 16, 16, device=dev)
    with torch.no_grad():
        want = torch.nn.functional.conv2d(x.double(), unit[0].weight.double(), unit[0].bias.double()).float()
        with _plain_convs_on_own_kernels(unit):
            assert "forward" in unit[0].__dict__ and "forward" in unit[1].__dict__
            got = unit[0](x)
            y2 = unit[1](got)                      # stride 2: the kernel declines, torch's op answers
        assert (got - want).abs().max().item() < 3e-5 * want.abs().max().item()
        assert y2.shape == (4, 32, 8, 8)
    assert "forward" not in unit[0].__dict__ and "forward" not in unit[1].__dict__
    with pytest.raises(RuntimeError):
        with _plain_convs_on_own_kernels(unit):
            raise RuntimeError("boom")
    assert "forward" not in unit[0].__dict__


def test_reconstruction_graph_replay_equals_the_eager_loop(dev, monkeypatch):
    """The captured iteration graph runs the same kernels in the same order as the eager loop: identical loss trace and alpha."""
    import tfmq_b200.quant.reconstruction as R
    from tfmq_b200.quant.reconstruction_util import RLOSS
    w_cali = (synth.latents((64, 3, 32, 32), 51),
              torch.randint(0, 1000, (64,), generator=torch.Generator().manual_seed(2)).float())
    res = {}
    for mode in (True, False):
        qnn = _qnn(dev, cali=True)
        qnn.set_quant_state(True, False)
        with torch.no_grad():
            qnn(*(d[:8].to(dev) for d in w_cali))
        qnn.disable_out_quantization()
        blk = qnn.model.down[1].block[0]
        blk.dropout.p = 0.0      # the unit runs in train() mode (data_utill.py:72): eager and replayed masks come from different
        #                          positions of the generator stream, everything else is the same arithmetic
        monkeypatch.setattr(R, "RECON_GRAPH", mode)
        trace = []
        monkeypatch.setattr(R, "LOSS_TRACE", trace)
        torch.manual_seed(0)
        R.block_reconstruction(qnn, blk, w_cali, batch_size=32, iters=30, w=0.01, opt_mode=RLOSS.MSE, asym=False,
                               b_range=(20, 2), warmup=0.2, multi_gpu=False)
        alphas = [m.wqtizer.alpha.detach().clone() for m in blk.modules() if hasattr(m, "wqtizer") and hasattr(m.wqtizer, "alpha")]
        res[mode] = (trace, alphas)
    tr_g, al_g = res[True]
    tr_e, al_e = res[False]
    assert len(tr_g) == len(tr_e) == 30
    # split-K weight gradients are reduced by TMA adds in arrival order, so single runs differ in the last bits
    assert max(abs(a - b) for a, b in zip(tr_g, tr_e)) <= 1e-3 * max(abs(v) for v in tr_e)
    moved = sum(int(((a - b).abs() > 2.5e-3).sum()) for a, b in zip(al_g, al_e))
    total = sum(a.numel() for a in al_g)
    print(f"graph vs eager: loss trace max diff {max(abs(a - b) for a, b in zip(tr_g, tr_e)):.3e}; alpha elements more than two "
          f"Adam steps apart: {moved} of {total}")
    assert moved <= 0.02 * total


def test_save_inout_matches_the_reference(dev):
    """`save_inout` / `GetLayerInpOut` (quant/data_utill.py:13-55, :109-169) -- the cached inputs and FP targets of
    a reconstruction unit -- against the reference's own function on the same model and data (tests/golden/inout_sdmini.pt, made
    on the CPU by make_golden.py::inout_golden): a ResBlock (x, emb), a BasicTransformerBlock (x, context), the middle block and
    an output block over a skip concatenation; symmetric (FP inputs) and asymmetric (inputs seen through the weight-quantised
    layers before the unit).  Tolerance 2e-5 of the tensor's largest magnitude on every sampled element (fp32 accumulation
    order of GPU vs CPU kernels; TF32 is off inside QuantModel.calibrating), 1e-6 relative on the sums."""
    from tfmq_b200.quant.data_utill import save_inout
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    g = load_golden("inout_sdmini.pt")
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp_model("sdmini", g["seed"]).to(dev), wq, aq, cali=True, softmax_a_bit=8,
                     aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).eval()
    cali = (synth.latents((16, 4, 16, 16), 173),
            torch.randint(0, 1000, (16,), generator=torch.Generator().manual_seed(174)).float(), synth.latents((16, 7, 96), 175))
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        qnn(*(d[:8].to(dev) for d in cali))
    qnn.disable_out_quantization()
    mods = dict(qnn.model.named_modules())
    worst = 0.0
    for (name, asym), want in g["units"].items():
        ins, outs = save_inout(qnn, mods[name], cali, asym=asym, use_act=False, batch_size=8)
        outs = outs if isinstance(outs, tuple) else (outs,)
        assert len(ins) == len(want["ins"]) and len(outs) == len(want["outs"]), (name, asym)
        for tag, got, ref in [("in", a, b) for a, b in zip(ins, want["ins"])] + [("out", a, b) for a, b in zip(outs, want["outs"])]:
            shape, sample, s1, s2 = ref
            t = got.detach().float().cpu()
            assert tuple(t.shape) == shape, (name, asym, tag, tuple(t.shape), shape)
            scale = max(sample.abs().max().item(), 1e-6)
            err = (t.flatten()[::g["stride"]] - sample).abs().max().item() / scale
            worst = max(worst, err)
            assert err < 2e-5, (name, asym, tag, err)
            assert abs(float(t.double().sum()) - s1) <= 1e-6 * max(1.0, s2), (name, asym, tag)
    print(f"[save_inout] {len(g['units'])} (unit, asym) cases against the reference: worst sampled deviation {worst:.2e} of the "
          "tensor's largest magnitude")
