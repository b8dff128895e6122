"""Shared test fixtures: seeded FP models (product host classes + tests/golden/synth.py), oracle specs."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)
import synth  # noqa: E402

CIFAR_CFG = dict(ch=128, ch_mult=[1, 2, 2, 2], num_res_blocks=2)
LDM4_CFG = dict(model_channels=224, num_head_channels=32)
SDMINI_CFG = dict(model_channels=64, num_heads=2)
SD_V14_CFG = dict(model_channels=320, num_heads=8)       # BASELINE configs[2]
CIN256_CFG = dict(model_channels=192, num_heads=1)       # BASELINE configs[4]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def fp_model(kind: str, seed: int = 1234):
    if kind == "cifar":
        from tfmq_b200.host.ddim_unet import Model, cifar10_config
        m = Model(cifar10_config())
    elif kind == "sdmini":
        from tfmq_b200.host.ldm_unet import UNetModel, sd_mini_config
        m = UNetModel(**sd_mini_config())
    elif kind in ("sd_v14", "cin256"):
        from tfmq_b200.host import ldm_unet as H
        m = H.UNetModel(**dict(sd_v14=H.sd_v14_config, cin256=H.cin256_config)[kind]())
    else:
        from tfmq_b200.host.ldm_unet import UNetModel, celebahq_ldm4_config
        m = UNetModel(**celebahq_ldm4_config())
    m.eval()
    synth.fill_state_dict(m, seed)
    return m


def first_stage_model(kind: str, seed: int = 1234):
    """Seeded first stage (product parameter container) matching tests/golden/make_golden.py::first_stage_golden."""
    from tfmq_b200.first_stage import FirstStageModel, first_stage_mini_config
    cfg = first_stage_mini_config(kind)
    m = FirstStageModel(**cfg)
    m.eval()
    synth.fill_state_dict(m, seed)
    if cfg["n_embed"] is not None:
        m.quantize.embedding.weight.data.copy_(synth.latents((cfg["n_embed"], cfg["embed_dim"]), 91))
    return m, cfg


def oracle_spec(sd, seed: int = 1234):
    from oracle import unet_ref
    return unet_ref.build_spec(sd, alpha_fn=lambda n, w, d: synth.synth_alpha(n, w, d, seed))


def full_size_inputs(name: str, g: dict):
    """(x, t, context) of the full-size SpatialTransformer goldens (tests/golden/make_golden.py::full_size_golden): seeded,
    regenerated here instead of stored."""
    from tfmq_b200.host import ldm_unet as H
    cfg = dict(sd_v14=H.sd_v14_config, cin256=H.cin256_config)[name]()
    x = synth.latents((2, cfg["in_channels"], 64, 64), g["x_seed"])
    ctx = synth.latents((2, g["tokens"], cfg["context_dim"]), g["ctx_seed"])
    return x, g["t"], ctx
