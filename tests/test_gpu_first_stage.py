"""GPU parity of the first-stage decode: DecoderEngine (library kernels through the C ABI) against the
CPU oracle (oracle/first_stage_ref.py) and the fixture produced by the reference's own VQModelInterface / AutoencoderKL
decode (tests/golden/first_stage.pt)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import first_stage_model, load_golden, synth

pytestmark = pytest.mark.gpu

TOL_IMAGE = 1e-4    # stated fp tolerance on the decoded image (values are O(1)): all layers are fp32-accurate kernels


def _ops():
    from tfmq_b200 import ops
    return ops


@pytest.mark.parametrize("n,c,c_out,h,w,n_embed", [
    (2, 3, 3, 16, 16, 96),        # the fixture's shape
    (1, 3, 3, 64, 64, 8192),      # vq-f4: 8192 codes of 3 dims
    (3, 4, 4, 7, 9, 0),           # kl-f8: no codebook, ragged pixel count
    (2, 2, 4, 5, 5, 1500),        # codebook size that is not a multiple of the staging chunk, c != c_out
])
def test_first_stage_input_kernel(dev, n, c, c_out, h, w, n_embed):
    from oracle import first_stage_ref as FS
    ops = _ops()
    z = synth.latents((n, c, h, w), 5)
    cb = synth.latents((n_embed, c), 6) if n_embed else None
    pw = synth.latents((c_out, c, 1, 1), 7)
    pb = synth.latents((c_out,), 8)
    inv = float(torch.tensor(1.0 / 0.18215, dtype=torch.float32))
    zs = 1. / 0.18215 * z
    out = torch.empty((n, c_out, h, w), device=dev)
    idx = torch.zeros((n * h * w,), dtype=torch.int32, device=dev)
    ops.first_stage_input(z.to(dev), inv, out, codebook=cb.to(dev) if n_embed else None, w=pw.to(dev), bias=pb.to(dev),
                          indices=idx if n_embed else None)
    if n_embed:
        zq, ridx = FS.vq_lookup(zs, cb)
        got = idx.cpu().long()
        bad = (got != ridx).nonzero().flatten()
        if bad.numel():      # only exact near-ties may differ (the distance is a rounded fp32 expression on both sides)
            zf = zs.permute(0, 2, 3, 1).reshape(-1, c).double()
            d = torch.cdist(zf[bad], cb.double()) ** 2
            gap = (d.gather(1, got[bad, None]) - d.gather(1, ridx[bad, None])).abs().max().item()
            assert gap < 1e-5 and bad.numel() <= max(1, got.numel() // 1000), (bad.numel(), gap)
        zq_dev = cb[got].view(n, h, w, c).permute(0, 3, 1, 2)
        zq_dev = zs + (zq_dev - zs)
    else:
        zq_dev = zs
    ref = F.conv2d(zq_dev.double(), pw.double(), pb.double())
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5
    # without post_quant_conv the kernel passes the (looked-up) latent through
    if c == c_out:
        out2 = torch.empty((n, c, h, w), device=dev)
        ops.first_stage_input(z.to(dev), inv, out2, codebook=cb.to(dev) if n_embed else None)
        assert (out2.cpu() - zq_dev).abs().max().item() < 1e-6


def test_conv_in_wide_output(dev):
    """The decoder's conv_in: 4 latent channels -> 512 (72 KB of transposed weights: opt-in shared memory)."""
    ops = _ops()
    x = synth.latents((2, 4, 16, 24), 11)
    w = synth.latents((512, 4, 3, 3), 12) * 0.2
    b = synth.latents((512,), 13)
    out = torch.empty((2, 16, 24, 512), device=dev)
    ops.conv_in(x.to(dev), w.to(dev), b.to(dev), out)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    assert (out.cpu().permute(0, 3, 1, 2).double() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("kind", ["vq", "vq-attn", "kl"])
def test_decoder_engine_matches_oracle_and_reference(dev, kind):
    from oracle import first_stage_ref as FS
    from tfmq_b200 import _lib
    g = load_golden("first_stage.pt")[kind]
    m, cfg = first_stage_model(kind)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    z = g["z"]
    ctx = _lib.context(0)
    before = ctx.launches
    img = m.decode_first_stage(z.to(dev)).cpu()
    assert ctx.launches > before, "decode did not go through the library's kernels"
    ref = FS.decode_first_stage(z, sd, cfg["scale_factor"], quantize=cfg["n_embed"] is not None)
    err_o = (img - ref).abs().max().item()
    err_g = (img - g["image"]).abs().max().item()
    print(f"first stage {kind}: |image| max {ref.abs().max():.3f}; vs oracle {err_o:.2e}; vs reference fixture {err_g:.2e}")
    assert err_o < TOL_IMAGE and err_g < TOL_IMAGE
    # graph replay == eager program; repeated decode is bit-identical
    eng = m.engine(z.shape[0], z.shape[2], z.shape[3], dev)
    assert torch.equal(m.decode_first_stage(z.to(dev)).cpu(), img)
    eng.use_graph = False
    assert torch.equal(eng.decode(z.to(dev)).cpu(), img)
    eng.use_graph = True
    if cfg["n_embed"] is not None:
        img2 = m.decode_first_stage(z.to(dev), force_not_quantize=True).cpu()
        assert (img2 - g["image_not_quantized"]).abs().max().item() < TOL_IMAGE
        # first_stage_model.decode: the same without the 1 / scale_factor step (scale_factor is 1 for the VQ config)
        assert torch.equal(m.decode(z.to(dev)).cpu(), img)
        # the chosen codes are the oracle's
        _, ridx = FS.vq_lookup(1. / cfg["scale_factor"] * z, sd["quantize.embedding.weight"])
        m.decode_first_stage(z.to(dev))
        assert torch.equal(eng.indices.cpu().long(), ridx)
    with pytest.raises(RuntimeError):
        eng.decode(z)                      # CPU latent
    with pytest.raises(RuntimeError):
        eng.decode(z[:1].to(dev))          # other batch than the program was built for
    # a scale_factor assigned after the first decode (the scripts set it on the LatentDiffusion object) must not replay the
    # graph captured with the old value
    if cfg["n_embed"] is None:
        old = m.scale_factor
        m.scale_factor = 2.0 * old
        img_half = m.decode_first_stage((2.0 * z).to(dev)).cpu()       # (2 z) / (2 s) = z / s, power-of-two scaling is exact
        m.scale_factor = old
        assert torch.equal(img_half, img)


def test_full_size_vq_f4_decode(dev):
    """LDM-4 CelebA-HQ's first stage at full size (vq-f4: 3x64x64 latent -> 3x256x256 image, 512-channel single-head
    attention over 4096 tokens, 8192-code lookup), one image against the CPU oracle."""
    from oracle import first_stage_ref as FS
    from tfmq_b200.first_stage import FirstStageModel, vq_f4_config
    cfg = vq_f4_config()
    m = FirstStageModel(**cfg)
    m.eval()
    synth.fill_state_dict(m, 1234)
    m.quantize.embedding.weight.data.copy_(synth.latents((cfg["n_embed"], cfg["embed_dim"]), 91))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(dev)
    z = synth.latents((1, 3, 64, 64), 93)
    img = m.decode_first_stage(z.to(dev)).cpu()
    assert tuple(img.shape) == (1, 3, 256, 256) and torch.isfinite(img).all()
    ref = FS.decode_first_stage(z, sd, cfg["scale_factor"], quantize=True)
    err = (img - ref).abs().max().item()
    print(f"vq-f4 full size: |image| max {ref.abs().max():.3f}, max abs err vs oracle {err:.2e}, "
          f"{m.engine(1, 64, 64, dev).launches_per_decode} launches per decode")
    assert err < 2 * TOL_IMAGE
