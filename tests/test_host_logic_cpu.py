"""CPU: host-side logic of the drop-in surface (module surgery, state conventions, schedules,
checkpoint plumbing) -- no kernel is launched."""
import math

import pytest
import torch

from helpers import fp_model, synth
from oracle import quant_ref as Q
from oracle import unet_ref as U


def _qnn(kind, cali=True):
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    return QuantModel(fp_model(kind), wq, aq, cali=cali, softmax_a_bit=8,
                      aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])


@pytest.mark.parametrize("kind,n_layers,n_emb", [("cifar", 97, 22), ("ldm", 73, 22)])
def test_surgery_matches_reference_rules(kind, n_layers, n_emb):
    from tfmq_b200.quant.quant_layer import QuantLayer
    qnn = _qnn(kind)
    names = [n for n, m in qnn.model.named_modules() if isinstance(m, QuantLayer)]
    sd = fp_model(kind).state_dict()
    assert names == U.wrapped_layer_names(sd)          # same leaves, same order as the oracle's restatement
    assert len(names) == n_layers
    emb = [n for n, m in qnn.model.named_modules() if isinstance(m, QuantLayer) and m.quant_emb]
    assert len(emb) == n_emb
    tib_layers = qnn.tib.temb_projs if kind == "cifar" else qnn.tib.emb_layers
    assert len(tib_layers) == n_emb
    # nothing named skip / op / shortcut / downsample.conv is wrapped; Conv1d stays a Conv1d
    for n, m in qnn.model.named_modules():
        leaf = n.split(".")[-1]
        if isinstance(m, QuantLayer):
            assert "skip" not in leaf and "op" not in leaf and "shortcut" not in leaf
            assert not n.endswith("downsample.conv")
    qnn.set_quant_state(True, True)
    qnn.disable_out_quantization()
    ql = qnn.quant_layers()
    assert [l.ignore_recon for l in (ql[0], ql[1], ql[2], ql[3], ql[-1])] == [True, False, True, False, True]
    assert ql[1].disable_aq and ql[3].disable_aq and not ql[4].disable_aq
    qnn.set_quant_state(True, True)
    assert not ql[0].use_wq and not ql[-1].use_wq and ql[1].use_wq       # ignore_recon overrides the toggle


def test_state_dict_keys_follow_reference_checkpoint_format():
    qnn = _qnn("cifar", cali=False)
    keys = set(qnn.state_dict().keys())
    assert "model.conv_in.w" in keys and "model.conv_in.b" in keys
    assert "model.down.0.block.0.conv1.w" in keys and "model.down.1.attn.0.q.w" in keys
    assert "model.down.0.downsample.conv.weight" in keys          # un-wrapped conv keeps torch's names
    assert not any("original_w" in k for k in keys)                # plain tensor attribute, not a buffer


def test_no_cpu_fallback():
    qnn = _qnn("cifar", cali=False)
    qnn.set_quant_state(True, True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qnn(torch.zeros(1, 3, 32, 32), torch.zeros(1))
    from tfmq_b200.engine import StepEngine
    with pytest.raises(RuntimeError, match="no CPU path"):
        StepEngine(qnn, batch=1)


def test_ddim_coefficients_match_oracle():
    from tfmq_b200 import samplers
    betas = synth.ddim_betas()
    seq = list(range(0, 1000, 20))
    assert samplers.ddim_coefficients(seq, betas) == U.ddim_coef_table(seq, betas)


def test_ldm_ddim_schedule_matches_reference_formulas():
    import numpy as np
    from tfmq_b200.samplers import DDIMSampler, make_beta_schedule, make_ddim_timesteps
    s = DDIMSampler(None)
    s.make_schedule(200)
    assert list(s.ddim_timesteps[:3]) == [1, 6, 11] and s.ddim_timesteps[-1] == 996
    betas = make_beta_schedule("linear", 1000, 0.0015, 0.0195)
    ac = np.cumprod(1 - betas)
    assert abs(float(s.ddim_alphas[0]) - ac[1]) < 1e-7 and abs(float(s.ddim_alphas_prev[0]) - ac[0]) < 1e-7
    rows = s.coefficient_rows()
    assert len(rows) == 200
    sa, s1, sp, c2, c1, order = rows[0]                # first sampling step = largest t; order 1 = the LDM summation order
    assert abs(sa - math.sqrt(ac[996])) < 1e-6 and abs(s1 - math.sqrt(1 - ac[996])) < 1e-6
    assert abs(sp - math.sqrt(ac[991])) < 1e-6 and abs(c2 - math.sqrt(1 - ac[991])) < 1e-6 and c1 == 0.0
    assert list(make_ddim_timesteps(50)) == [i + 1 for i in range(0, 1000, 20)]


def test_temp_decay_and_round_loss_match_oracle():
    from tfmq_b200.quant.reconstruction_util import LinearTempDecay
    d = LinearTempDecay(20000, 0.2, 20, 2)
    for t in (1, 3999, 4000, 4001, 12000, 20000):
        assert d(t) == Q.temp_decay(t, 20000, 0.2, 20, 2)
    from tfmq_b200.quant.quant_layer import lp_loss, REDUCTION
    a, b = synth.latents((4, 3, 8, 8), 1), synth.latents((4, 3, 8, 8), 2)
    assert torch.equal(lp_loss(a, b), Q.lp_loss(a, b))
    assert torch.equal(lp_loss(a, b, 2.4, REDUCTION.ALL), Q.lp_loss(a, b, 2.4, True))


def test_act_tables_from_ckpt_and_fsc_index():
    from tfmq_b200.quant.calibration import act_tables_from_ckpt
    ckpt = {"weight": {}, "act_0": {"a": 0}, "act_1": {"a": 1}, "act_2": {"a": 2}}
    assert [t["a"] for t in act_tables_from_ckpt(ckpt)] == [0, 1, 2]


def test_adaround_quantizer_matches_oracle_on_cpu_math():
    """AdaRoundQuantizer's pure-torch forward (hard / soft / init) is device-agnostic arithmetic."""
    from tfmq_b200.quant.adaptive_rounding import RMODE, AdaRoundQuantizer
    from tfmq_b200.quant.quant_layer import UniformAffineQuantizer
    w = synth.latents((8, 4, 3, 3), 5) * 0.1
    d, z = Q.channel_wise(Q.minmax_scale, w, 16)
    u = UniformAffineQuantizer(bits=4, channel_wise=True)
    u.delta, u.zero_point, u.init = d, z, True
    a = AdaRoundQuantizer(u, w, RMODE.LEARNED_HARD_SIGMOID)
    assert torch.equal(a.alpha.detach(), Q.adaround_init_alpha(w, d))
    assert torch.equal(a(w), Q.adaround_fake_quant(w, d, z, a.alpha.detach(), 16))
    a.soft_tgt = True
    assert torch.equal(a(w).detach(), Q.adaround_fake_quant(w, d, z, a.alpha.detach(), 16, soft=True))


def test_ldm_unets_have_the_reference_state_dict_keys():
    """Reference checkpoints must load unchanged: same keys, same shapes, same order, for every supported LDM config
    (fixture written by tests/golden/make_golden.py from the reference's own UNetModel)."""
    import torch
    from helpers import load_golden
    from tfmq_b200.host import ldm_unet as H
    g = load_golden("unet_keys.pt")
    for name, cfg in (("ldm4", H.celebahq_ldm4_config()), ("sd_mini", H.sd_mini_config()), ("sd_v14", H.sd_v14_config()),
                      ("cin256", H.cin256_config())):
        with torch.device("meta"):
            m = H.UNetModel(**cfg)
        mine = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert mine == g[name], name


def test_quant_model_surgery_on_transformer_unet_matches_reference():
    """QuantModel's module surgery on a SpatialTransformer UNet: the same QuantLayer / QuantBasicTransformerBlock /
    QuantResBlock placement as the reference's QuantModel (module names and classes; Identity placement aside)."""
    from helpers import fp_model, load_golden
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    g = load_golden("unet_keys.pt")
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp_model("sdmini"), wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    mine = {n: type(m).__name__ for n, m in qnn.named_modules()}
    ref = dict(g["sd_mini_quant_modules"])
    interesting = ("QuantLayer", "QuantBasicTransformerBlock", "QuantResBlock", "UniformAffineQuantizer")
    want = {n: c for n, c in ref.items() if c in interesting}
    got = {n: c for n, c in mine.items() if c in interesting}
    assert got == want
    assert sum(c == "QuantLayer" for c in got.values()) > 100


def test_first_stage_configs_and_decoder_program():
    """First-stage decode, host side: the full-size configs build the reference's module tree, and the
    DecoderEngine program traces with every kernel call recorded instead of launched (tools/dry_trace_first_stage.py, in a
    subprocess because it redirects torch allocations)."""
    import json
    import os
    import subprocess
    import sys
    from tfmq_b200.first_stage import FirstStageModel, kl_f8_config, vq_f4_config
    with torch.device("meta"):
        vq, kl = FirstStageModel(**vq_f4_config()), FirstStageModel(**kl_f8_config())
    sd = vq.state_dict()
    assert tuple(sd["quantize.embedding.weight"].shape) == (8192, 3)
    assert tuple(sd["decoder.conv_in.weight"].shape) == (512, 3, 3, 3)
    assert tuple(sd["decoder.mid.attn_1.q.weight"].shape) == (512, 512, 1, 1)
    assert tuple(sd["decoder.up.0.block.2.conv2.weight"].shape) == (128, 128, 3, 3)
    assert tuple(sd["decoder.up.1.block.0.nin_shortcut.weight"].shape) == (256, 512, 1, 1)
    assert "decoder.up.0.upsample.conv.weight" not in sd and "decoder.up.2.upsample.conv.weight" in sd
    assert tuple(sd["decoder.conv_out.weight"].shape) == (3, 128, 3, 3)
    sk = kl.state_dict()
    assert "quantize.embedding.weight" not in sk and tuple(sk["post_quant_conv.weight"].shape) == (4, 4, 1, 1)
    assert len([k for k in sk if k.startswith("decoder.up.3.")]) > 0 and kl.scale_factor == 0.18215
    # 3 (vq-f4) / 4 (kl-f8) levels x 3 ResnetBlocks + 2 mid blocks, 2 convs each, + shortcuts + attention + conv_in / out
    assert sum(1 for k in sd if k.endswith("conv1.weight")) == 11 and sum(1 for k in sk if k.endswith("conv1.weight")) == 14
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "dry_trace_first_stage.py")
    out = subprocess.run([sys.executable, tool], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = {ln.split(" ", 2)[0]: (int(ln.split(" ", 2)[1]), json.loads(ln.split(" ", 2)[2])) for ln in out.stdout.strip().splitlines()}
    for kind, n_attn in (("vq", 1), ("vq-attn", 3), ("kl", 1)):
        n_ops, c = got[kind]
        assert c["tfmq_first_stage_input"] == 1 and c["tfmq_conv_in"] == 1 and c["tfmq_conv_out"] == 1
        assert c["tfmq_attention"] == n_attn
        # 6 ResnetBlocks (2 convs each) + 1 nin_shortcut + 1 Upsample conv + (stacked qkv + proj_out) per attention
        assert c["tfmq_conv_h16"] == 6 * 2 + 1 + 1 + 2 * n_attn
        # the attention output reaches proj_out as fp16 planes written by the attention kernel: no split launch for it
        assert c["tfmq_act_prepare"] == c["tfmq_conv_h16"] - n_attn + 1      # + the final GN + SiLU before conv_out
    # the same program EXECUTED on the CPU by torch stand-ins with the kernels' semantics (tests/emulated_ops.py): the host
    # logic (op order, strides, residual aliasing, GroupNorm statistics through conv epilogues, fp16 planes) reproduces the
    # oracle and the reference fixture
    out = subprocess.run([sys.executable, tool, "--emulate"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    errs = {ln.split()[0]: float(ln.split()[-1]) for ln in out.stdout.strip().splitlines()}
    assert set(errs) == {"vq", "vq-attn", "kl"} and max(errs.values()) < 2e-5, errs


def test_ddim_runner_schedules_and_sample_image_plumbing(monkeypatch):
    """runners.Diffusion against the reference's own runner (fixture runner_ddim.pt): beta schedules and logvar bit for bit;
    `sample_image`'s timestep sequence / argument plumbing by routing its `generalized_steps` call to the oracle's CPU
    restatement with the same stand-in UNet the fixture used (the product's generalized_steps itself needs the GPU engine and
    is covered by the GPU tests)."""
    from types import SimpleNamespace as NS
    from helpers import load_golden
    from tfmq_b200 import runners as R
    g = load_golden("runner_ddim.pt")
    for name, want in g["schedules"].items():
        got = R.get_beta_schedule(name, beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)
        assert torch.equal(torch.from_numpy(got), want), name
    with pytest.raises(NotImplementedError):
        R.get_beta_schedule("cosine", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=10)

    def stub_eps(x, t, c=None):          # tests/golden/make_golden.py::stub_eps
        return 0.3 * x.roll(1, -1) - (0.2 * x) * (t.float() / 1000.0)[:, None, None, None]

    calls = []

    def oracle_steps(x, seq, model, b, **kw):
        calls.append(dict(kw, seq=list(seq)))
        seq = list(seq)
        stop = kw.get("untill_fake_t")
        xs, x0 = U.generalized_steps(x, seq, lambda xt, t, k: model(xt, t), b, eta=kw.get("eta", 0.0))
        x_t = t_t = None
        if stop is not None and stop <= len(seq):            # the reference breaks before the stop-th UNet call
            x_t, t_t = xs[stop - 1], torch.ones(x.size(0)) * list(reversed(seq))[stop - 1]
        return xs, x0, x_t, t_t
    monkeypatch.setattr(R, "generalized_steps", oracle_steps)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            R.Diffusion(NS(), NS(diffusion=NS(), model=NS()))
    for var in ("fixedlarge", "fixedsmall"):
        cfg = NS(model=NS(var_type=var), diffusion=NS(beta_schedule="linear", beta_start=1e-4, beta_end=0.02,
                                                       num_diffusion_timesteps=1000))
        for skip in ("uniform", "quad"):
            want = g[(var, skip)]
            r = R.Diffusion(NS(skip_type=skip, timesteps=10, sample_type="generalized", eta=0.0), cfg, device="cpu")
            assert torch.equal(r.betas, want["betas"]) and torch.equal(r.logvar, want["logvar"]) and r.num_timesteps == 1000
            full, x_t, t_t = r.sample_image(g["x"].clone(), stub_eps)
            assert torch.equal(full, want["full"]) and x_t is None
            _, x_t, t_t = r.sample_image(g["x"].clone(), stub_eps, untill_fake_t=4, tot=7, cali_ckpt={"k": 1}, t_max=3)
            assert torch.equal(x_t, want["x_t"]) and torch.equal(t_t, want["t_t"])
            assert calls[-1]["tot"] == 7 and calls[-1]["t_max"] == 3 and calls[-1]["cali_ckpt"] == {"k": 1}
            pair, _, _ = r.sample_image(g["x"].clone(), stub_eps, last=False)
            assert len(pair) == 2 and torch.equal(pair[0][-1], want["full"]) and len(pair[1]) == 10
    assert calls[0]["seq"] == list(range(0, 1000, 100))
    r.args.sample_type = "ddpm_noisy"
    with pytest.raises(NotImplementedError):
        r.sample_image(g["x"], stub_eps)
    # quantize() without --ptq hands the FP model back; image-space transform
    r.args.ptq = False
    m = torch.nn.Identity()
    assert r.quantize(m) == (m, None, None, None)
    img = R.inverse_data_transform(NS(data=NS(rescaled=True, logit_transform=False)), torch.tensor([-3.0, -1.0, 0.0, 1.0, 2.0]))
    assert torch.equal(img, torch.tensor([0.0, 0.0, 0.5, 1.0, 1.0]))


def test_ldm_script_shells():
    """runners.LatentDiffusion / DiffusionWrapper: what sample_diffusion_ldm.py:438-479 and the samplers read."""
    from types import SimpleNamespace as NS
    import numpy as np
    from helpers import first_stage_model
    from tfmq_b200 import runners as R
    from tfmq_b200.samplers import DDIMSampler, PLMSSampler
    seen = []

    def unet(x, t, context=None):
        seen.append(context)
        return x * 2

    ld = R.LatentDiffusion(unet, scale_factor=0.18215, linear_start=0.00085, linear_end=0.012, conditioning_key="crossattn",
                           image_size=64, channels=4)
    x, t = torch.ones(2, 4, 8, 8), torch.tensor([5.0, 5.0])
    assert torch.equal(ld.apply_model(x, t, [torch.zeros(2, 3, 7), torch.ones(2, 4, 7)]), x * 2)
    assert tuple(seen[-1].shape) == (2, 7, 7)                       # c_crossattn concatenated along the token axis
    # FSC attributes as the script installs them for a 50-table checkpoint (:473-479)
    ckpt = {"weight": {}, **{f"act_{k}": {"k": k} for k in range(50)}}
    n_tables = len(ckpt) - 1
    ld.model.tot, ld.model.t_max, ld.model.ckpt, ld.model.iter = 1000 // n_tables, n_tables - 1, ckpt, 0
    assert [ld.model.fsc_index(t_) for t_ in (981, 961, 21, 1)] == [0, 1, 48, 49]
    with pytest.raises(RuntimeError, match="DDIMSampler"):
        ld.model(x, t)
    # the reference's call shape: DDIMSampler(<LatentDiffusion>) picks up UNet, FSC tables and noise schedule
    for cls in (DDIMSampler, PLMSSampler):
        s = cls(ld)
        direct = cls(unet, linear_start=0.00085, linear_end=0.012, timesteps=1000, ckpt=ckpt)
        assert s.model is unet and s.ckpt is ckpt and s.ddpm_num_timesteps == 1000
        assert np.array_equal(s.alphas_cumprod, direct.alphas_cumprod)
    assert DDIMSampler(unet).ddpm_num_timesteps == 1000 and DDIMSampler(unet).ckpt is None
    # decode goes to the first stage with the model's scale factor (and there is no CPU path behind it)
    with pytest.raises(RuntimeError, match="first_stage_model"):
        ld.decode_first_stage(x)
    fs, _ = first_stage_model("kl")
    ld.first_stage_model = fs
    with pytest.raises(RuntimeError):
        ld.decode_first_stage(torch.zeros(1, 4, 16, 16))
    assert fs.scale_factor == 0.18215
    assert R.quantize_ldm(NS(ptq=False), ld) is ld
    with pytest.raises(NotImplementedError):
        R.DiffusionWrapper(unet, "concat")


def test_fsc_table_index_follows_the_number_of_calibrated_tables():
    """DiffusionWrapper.forward picks act_k with k = t_max - (t - 1) // tot where the scripts set tot = 1000 // n_tables and
    t_max = n_tables - 1 (sample_diffusion_ldm.py:475-477, ldm/models/diffusion/ddpm.py:1402-1405): the index depends on the
    number of CALIBRATED tables, not on the number of sampling steps S."""
    from tfmq_b200.samplers import DDIMSampler

    class _Stub:                      # only .model / attribute lookups are touched before sampling
        pass
    ckpt = {"weight": {}}
    for k in range(4):
        ckpt[f"act_{k}"] = {"k": k}
    smp = DDIMSampler(_Stub(), ckpt=ckpt)
    for S in (4, 8, 20, 2):
        smp.make_schedule(S)
        ts = [float(t) for t in reversed(smp.ddim_timesteps.tolist())]
        got = [d["k"] for d in smp.fsc_tables(ts)]
        want = [int(3 - (int(t) - 1) // 250) for t in ts]        # tot = 1000 // 4, t_max = 3
        assert got == want and all(0 <= k <= 3 for k in got), (S, got, want)
    # the attributes the scripts install on the wrapper win over the table count
    smp.tot, smp.t_max = 500, 1
    smp.make_schedule(4)
    ts = [float(t) for t in reversed(smp.ddim_timesteps.tolist())]
    assert [d["k"] for d in smp.fsc_tables(ts)] == [int(1 - (int(t) - 1) // 500) for t in ts]


def test_quant_model_engine_invalidation_and_calibration_context():
    """QuantModel.forward has two routes and no silent third one: the torch module graph inside `calibrating()` / under
    autograd / for the lazy-initialisation forward, the sm_100a step engine otherwise (a CPU tensor then raises).  Anything
    that changes what the engine froze at trace time drops it."""
    qnn = _qnn("cifar", cali=False)
    qnn.set_quant_state(False, False)
    x, t = synth.latents((1, 3, 32, 32), 5), torch.tensor([10.0])
    with qnn.calibrating(), torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        qnn(x, t)                                       # module graph: its QuantLayers refuse CPU tensors themselves
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        qnn(x, t)                                       # outside: the engine route, which needs a CUDA tensor
    sentinel = object()
    for mutate in (lambda: qnn.set_quant_state(False, False), lambda: qnn.disable_out_quantization(),
                   lambda: qnn.set_running_stat(False), lambda: qnn.load_state_dict(qnn.state_dict()),
                   lambda: qnn.float()):
        qnn._engine = sentinel
        mutate()
        assert qnn._engine is None
    qnn._engine = sentinel
    with qnn.calibrating():
        pass
    assert qnn._engine is None


def test_timestep_embedding_frequency_rows_are_cached_and_unchanged():
    """The frequency row of the sinusoidal embedding is evaluated on the host exactly as the reference does (ldm util.py:151-171,
    ddim/models/diffusion.py:6-24) and kept per device, so a captured reconstruction iteration contains no pageable upload."""
    from tfmq_b200.host import ddim_unet, ldm_unet
    t = torch.tensor([0.0, 1.0, 500.0, 999.0])
    half = 64
    f_ldm = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    a = t[:, None].float() * f_ldm[None]
    assert torch.equal(ldm_unet.timestep_embedding(t, 2 * half), torch.cat([a.cos(), a.sin()], dim=-1))
    assert ldm_unet._frequencies(half, 10000, t.device) is ldm_unet._frequencies(half, 10000, t.device)
    f_ddim = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    b = t.float()[:, None] * f_ddim[None, :]
    assert torch.equal(ddim_unet.get_timestep_embedding(t, 2 * half), torch.cat([b.sin(), b.cos()], dim=1))
    assert ddim_unet._frequencies(half, t.device) is ddim_unet._frequencies(half, t.device)
    # odd widths keep the reference's zero padding
    assert ldm_unet.timestep_embedding(t, 2 * half + 1).shape == (4, 2 * half + 1)
    assert ddim_unet.get_timestep_embedding(t, 2 * half + 1).shape == (4, 2 * half + 1)


def test_weight_gradient_buffers_are_trimmed_between_units_only():
    """A captured iteration graph holds the addresses of the cached plane buffers: the cache never shrinks inside a unit
    (`_wgrad_buffers`), only through `trim_buffers` at the start of the next one."""
    from tfmq_b200.quant import tc_autograd as T
    T._WG_BUF.clear()
    for i in range(12):
        T._wgrad_buffers(torch.device("cpu"), 16, 16, 1, 4, 4 + 4 * i)
    assert len(T._WG_BUF) == 12
    first = T._wgrad_buffers(torch.device("cpu"), 16, 16, 1, 4, 4)
    assert first is T._wgrad_buffers(torch.device("cpu"), 16, 16, 1, 4, 4)
    T.trim_buffers(keep=16)
    assert len(T._WG_BUF) == 12
    T.trim_buffers()
    assert len(T._WG_BUF) == 0
