"""The DDIM runner surface of the sampling scripts: `Diffusion(args, config)` with `sample_image`, the quantise-or-calibrate
plumbing of `Diffusion.sample`, and batched sample generation -- the drop-in for the parts of the reference's
`ddim/runners/diffusion.py` that sit on the hot path:

    get_beta_schedule            :37-68    (float64 numpy schedules -> fp32 betas)
    Diffusion.__init__           :71-106   (betas, num_timesteps, logvar)
    Diffusion.sample, PTQ block  :245-312  (QuantModel + load_cali_model, or calibration-data generation + cali_model)
    Diffusion.sample_fid         :326-364  (here `sample_batches`: the same loop without the PNG / npz file output)
    Diffusion.sample_image       :429-476  (timestep sequence, generalized_steps with the FSC arguments)
    inverse_data_transform       ddim/datasets/__init__.py:206-215

and the LDM side (`sample_diffusion_ldm.py:438-547`, `ldm/models/diffusion/ddpm.py:1385-1405`): `LatentDiffusion` /
`DiffusionWrapper` shells that carry what the samplers and the decode read (`.model.diffusion_model`, the FSC attributes
`.tot .t_max .ckpt .iter`, `.first_stage_model`, `.scale_factor`, the noise schedule) and `quantize_ldm`, the script's `--ptq` block.

`args` / `config` are the namespaces the reference's `sample_diffusion_ddim.py` builds (argparse + YAML): the attributes
read here are args.{ptq, wq, aq, use_aq, cali, cali_ckpt, cali_save_path, softmax_a_bit, q_mode, timesteps, interval_length,
skip_type, sample_type, eta, asym, running_stat} and config.{diffusion.*, model.var_type, data.{channels, image_size,
rescaled, logit_transform}, sampling.batch_size}.  Checkpoint download, datasets, training and image files are outside the
path.  Sampling itself runs on the fused step engine (samplers.generalized_steps); there is no CPU path.
"""
from __future__ import annotations

import logging
import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import dist_utils
from .quant.calibration import cali_model, load_cali_model
from .quant.data_generate import generate_cali_data_ddim, generate_cali_data_ldm
from .quant.quant_layer import Scaler
from .quant.quant_model import QuantModel
from .quant.reconstruction_util import RLOSS
from .samplers import generalized_steps

logger = logging.getLogger(__name__)

_SCHEDULES = {
    "quad": lambda lo, hi, n: np.linspace(lo ** 0.5, hi ** 0.5, n, dtype=np.float64) ** 2,
    "linear": lambda lo, hi, n: np.linspace(lo, hi, n, dtype=np.float64),
    "const": lambda lo, hi, n: hi * np.ones(n, dtype=np.float64),
    "jsd": lambda lo, hi, n: 1.0 / np.linspace(n, 1, n, dtype=np.float64),            # 1/T, 1/(T-1), ..., 1
    "sigmoid": lambda lo, hi, n: 1 / (np.exp(-np.linspace(-6, 6, n)) + 1) * (hi - lo) + lo,
}


def get_beta_schedule(beta_schedule: str, *, beta_start: float, beta_end: float, num_diffusion_timesteps: int) -> np.ndarray:
    if beta_schedule not in _SCHEDULES:
        raise NotImplementedError(beta_schedule)
    betas = _SCHEDULES[beta_schedule](beta_start, beta_end, num_diffusion_timesteps)
    assert betas.shape == (num_diffusion_timesteps,)
    return betas


def inverse_data_transform(config, x: torch.Tensor) -> torch.Tensor:
    """Model space -> [0, 1] images."""
    if hasattr(config, "image_mean"):
        x = x + config.image_mean.to(x.device)[None, ...]
    if getattr(config.data, "logit_transform", False):
        x = torch.sigmoid(x)
    elif getattr(config.data, "rescaled", False):
        x = (x + 1.0) / 2.0
    return torch.clamp(x, 0.0, 1.0)


class Diffusion:
    def __init__(self, args, config, device=None):
        self.args, self.config = args, config
        config.split_shortcut = True
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("tfmq_b200.runners.Diffusion needs an sm_100a GPU (no CPU path); pass device= explicitly "
                                   "only to build the schedule on the host")
            device = torch.device("cuda")
        self.device = torch.device(device)
        d = config.diffusion
        betas = get_beta_schedule(d.beta_schedule, beta_start=d.beta_start, beta_end=d.beta_end,
                                  num_diffusion_timesteps=d.num_diffusion_timesteps)
        self.betas = torch.from_numpy(betas).float().to(self.device)
        self.num_timesteps = self.betas.shape[0]
        self.model_var_type = config.model.var_type
        cum = (1.0 - self.betas).cumprod(dim=0)
        cum_prev = torch.cat([torch.ones(1, device=self.device), cum[:-1]], dim=0)
        if self.model_var_type == "fixedlarge":
            self.logvar = self.betas.log()
        elif self.model_var_type == "fixedsmall":
            self.logvar = (self.betas * (1.0 - cum_prev) / (1.0 - cum)).clamp(min=1e-20).log()

    # ------------------------------------------------------------------ quantise (sampling) or calibrate
    def quantize(self, model) -> Tuple[torch.nn.Module, Optional[int], Optional[dict], Optional[int]]:
        """The `if self.args.ptq:` block of the reference's `sample`: returns (model, tot, cali_ckpt, t_max), the three FSC
        arguments `sample_image` forwards to `generalized_steps`.  With args.cali the model is calibrated and saved to
        args.cali_save_path instead (the reference then exits; here the calibrated QuantModel is returned)."""
        a = self.args
        if not getattr(a, "ptq", False):
            return model, None, None, None
        scaler = Scaler.MSE if a.cali else Scaler.MINMAX
        wq_params = dict(bits=a.wq, channel_wise=True, scaler=scaler)
        aq_params = dict(bits=a.aq, channel_wise=False, scaler=scaler, leaf_param=a.use_aq)
        qnn = QuantModel(model=model, wq_params=wq_params, aq_params=aq_params, cali=bool(a.cali),
                         softmax_a_bit=a.softmax_a_bit, aq_mode=a.q_mode)
        qnn.to(self.device)
        qnn.eval()
        ch, size = self.config.data.channels, self.config.data.image_size
        if not a.cali:
            init = (torch.randn(1, ch, size, size), torch.randint(0, 1000, (1,)))
            load_cali_model(qnn, init, use_aq=a.use_aq, path=a.cali_ckpt)
            tot = cali_ckpt = t_max = None
            if a.use_aq:
                cali_ckpt = torch.load(a.cali_ckpt, map_location="cpu", weights_only=False)
                n_tables = len(cali_ckpt) - 1                        # every key but 'weight' is one act_k table
                tot, t_max = 1000 - n_tables, n_tables - 1
            return qnn, tot, cali_ckpt, t_max
        logger.info("Generating calibration data...")
        per_step = 256
        # the FP sampler runs on the same engine in its all-floating-point state, so the QuantModel is what samples here
        xs, ts = generate_cali_data_ddim(runnr=self, model=qnn, T=a.timesteps, c=1, batch_size=per_step,
                                         shape=(ch, size, size))
        kept = [slice(i * per_step, (i + 1) * per_step) for i in range(0, a.timesteps, a.interval_length)]
        w_cali_data = [torch.cat([xs[s] for s in kept]), torch.cat([ts[s] for s in kept])]
        logger.info("Calibration data generated.")
        cali_model(qnn=qnn, use_aq=a.use_aq, path=a.cali_save_path, running_stat=a.running_stat, interval=per_step,
                   w_cali_data=w_cali_data, a_cali_data=(xs, ts), iters=20000, batch_size=32, w=0.01, asym=a.asym,
                   warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
        return qnn, None, None, None

    # ------------------------------------------------------------------ sampling
    def timestep_sequence(self):
        """The DDPM times the sampler visits (ascending), `args.timesteps` of them."""
        kind = self.args.skip_type
        if kind == "uniform":
            return range(0, self.num_timesteps, self.num_timesteps // self.args.timesteps)
        if kind == "quad":
            return [int(s) for s in np.linspace(0, np.sqrt(self.num_timesteps * 0.8), self.args.timesteps) ** 2]
        raise NotImplementedError(kind)

    def sample_image(self, x, model, last=True, untill_fake_t=114514, tot=None, cali_ckpt=None, t_max=None):
        if self.args.sample_type != "generalized":
            # the reference's other branch ("ddpm_noisy") imports a module its tree does not contain
            raise NotImplementedError(f"sample_type {self.args.sample_type}")
        xs, x0_preds, x_t, t_t = generalized_steps(x, self.timestep_sequence(), model, self.betas, eta=self.args.eta,
                                                   untill_fake_t=untill_fake_t, tot=tot, cali_ckpt=cali_ckpt, t_max=t_max)
        out = (xs, x0_preds)
        if last:
            out = out[0][-1]
        return out, x_t, t_t

    @torch.no_grad()
    def sample_batches(self, model, total: int, untill_fake_t=114514, tot=None, cali_ckpt=None, t_max=None,
                       generator: Optional[torch.Generator] = None) -> np.ndarray:
        """`sample_fid` without the files: `total` images in rounds of config.sampling.batch_size, as uint8 [n, H, W, C].
        Under torch.distributed every rank takes its contiguous share of the rounds (independent batches, no collective).
        With a `generator` seeded identically on every rank, each rank draws the x_T of ALL rounds and keeps its own, so
        the union over the ranks is bit-identical to a single-GPU run with that seed; without one the ranks
        draw independently."""
        n = self.config.sampling.batch_size
        ch, size = self.config.data.channels, self.config.data.image_size
        n_rounds = math.ceil(total / n)
        mine = dist_utils.shard_range(n_rounds, dist_utils.rank(), dist_utils.world())
        out = []
        for r in (range(n_rounds) if generator is not None else mine):
            x = torch.randn(n, ch, size, size, device=self.device, generator=generator)
            if r not in mine:
                continue
            x = self.sample_image(x, model, untill_fake_t=untill_fake_t, tot=tot, cali_ckpt=cali_ckpt, t_max=t_max)[0]
            x = inverse_data_transform(self.config, x)
            keep = min(n, total - r * n)
            out.append((x[:keep].permute(0, 2, 3, 1).cpu().numpy() * 255.).round().astype(np.uint8))
        return np.concatenate(out, axis=0) if out else np.zeros((0, size, size, ch), dtype=np.uint8)

# ---------------------------------------------------------------------------------------------- LDM scripts
class DiffusionWrapper(torch.nn.Module):
    """Holder with the reference's attribute names: `.diffusion_model` (the UNet; the QuantModel after `--ptq`) and,
    once activation quantisation is calibrated, the FSC attributes the script installs (`.tot .t_max .ckpt .iter`).
    The per-call `load_state_dict` switch of the reference's forward (ddpm.py:1402-1405) is not reproduced: the samplers
    read `.ckpt` and switch the activation parameters on the device, one row copy per step."""

    def __init__(self, diffusion_model, conditioning_key: Optional[str] = None):
        super().__init__()
        if conditioning_key not in (None, "crossattn"):
            raise NotImplementedError(f"conditioning_key {conditioning_key}: the configs of the path use None / crossattn")
        self.diffusion_model, self.conditioning_key = diffusion_model, conditioning_key

    def fsc_index(self, t: int) -> int:
        """The `act_k` table the reference loads for DDPM time t: t_max - (t - 1) // tot."""
        return int(self.t_max - (int(t) - 1) // self.tot)

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None):
        if hasattr(self, "tot"):
            raise RuntimeError("DiffusionWrapper.forward with FSC tables: sample through DDIMSampler / PLMSSampler, which "
                               "select the per-step activation parameters on the device")
        if self.conditioning_key is None:
            return self.diffusion_model(x, t)
        return self.diffusion_model(x, t, context=torch.cat(c_crossattn, 1))


class LatentDiffusion(torch.nn.Module):
    """What the sampling scripts hold as `model`: `.model` (DiffusionWrapper), `.first_stage_model`, the noise schedule
    and `decode_first_stage` -- without Lightning, the encoders, the losses or training (outside the path)."""

    def __init__(self, unet, first_stage_model=None, scale_factor: float = 1.0, timesteps: int = 1000,
                 linear_start: float = 0.0015, linear_end: float = 0.0195, conditioning_key: Optional[str] = None,
                 image_size: Optional[int] = None, channels: Optional[int] = None):
        super().__init__()
        self.model = DiffusionWrapper(unet, conditioning_key)
        self.first_stage_model = first_stage_model
        self.scale_factor, self.num_timesteps = float(scale_factor), timesteps
        self.linear_start, self.linear_end = linear_start, linear_end
        self.image_size = image_size if image_size is not None else getattr(unet, "image_size", None)
        self.channels = channels if channels is not None else getattr(unet, "in_channels", None)

    def apply_model(self, x_noisy, t, cond=None):
        if cond is None:
            return self.model(x_noisy, t)
        return self.model(x_noisy, t, c_crossattn=cond if isinstance(cond, list) else [cond])

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids: bool = False, force_not_quantize: bool = False):
        fs = self.first_stage_model
        if fs is None:
            raise RuntimeError("LatentDiffusion.decode_first_stage: no first_stage_model (first_stage.FirstStageModel) was given")
        fs.scale_factor = self.scale_factor
        return fs.decode_first_stage(z, predict_cids=predict_cids, force_not_quantize=force_not_quantize)


def quantize_ldm(opt, model: LatentDiffusion, device="cuda", context_shape=None, cali_data=None):
    """The `if opt.ptq:` block of sample_diffusion_ldm.py:456-547 on a `LatentDiffusion` shell.  Sampling (`not opt.cali`):
    wraps the UNet in a QuantModel, loads opt.cali_ckpt, installs it as `model.model.diffusion_model` and, with
    opt.use_aq, the FSC attributes `.tot .t_max .ckpt .iter`.  Calibration (`opt.cali`): generates the calibration data
    with the FP sampler, runs `cali_model` (saves to opt.cali_save_path) and installs the calibrated QuantModel.
    opt: ptq, cali, wq, aq, use_aq, softmax_a_bit, q_mode, cali_ckpt, cali_save_path, custom_steps, interval_length, plms,
    eta, asym, running_stat[, no_grad_ckpt].
    Conditional UNets (txt2img.py:394-470, latent_imagenet_diffusion.py): context_shape = (tokens, context_dim) adds the
    conditioning tensor to the dummy forward of `load_cali_model`; their calibration data come from guided sampling with
    encoder outputs (quant.data_generate.generate_cali_data_conditional) and are passed in as `cali_data` = (x_t, t, c),
    used for both the weight and the activation phase as those scripts do."""
    if not getattr(opt, "ptq", False):
        return model
    scaler = Scaler.MSE if opt.cali else Scaler.MINMAX
    wq_params = dict(bits=opt.wq, channel_wise=True, scaler=scaler)
    aq_params = dict(bits=opt.aq, channel_wise=False, scaler=scaler, leaf_param=opt.use_aq)
    unet = model.model.diffusion_model
    setattr(unet, "split", True)
    qnn = QuantModel(model=unet, wq_params=wq_params, aq_params=aq_params, cali=bool(opt.cali),
                     softmax_a_bit=opt.softmax_a_bit, aq_mode=opt.q_mode)
    qnn.to(device)
    qnn.eval()
    if getattr(opt, "no_grad_ckpt", False):
        qnn.set_grad_ckpt(False)
    shape = [model.channels, model.image_size, model.image_size]
    if not opt.cali:
        init = (torch.randn(1, *shape), torch.randint(0, 1000, (1,)))
        if context_shape is not None:
            init += (torch.randn(1, *context_shape),)
        load_cali_model(qnn, init, use_aq=opt.use_aq, path=opt.cali_ckpt)
        model.model.diffusion_model = qnn
        if opt.use_aq:
            cali_ckpt = torch.load(opt.cali_ckpt, map_location="cpu", weights_only=False)
            n_tables = len(cali_ckpt) - 1
            model.model.tot, model.model.t_max = 1000 // n_tables, n_tables - 1
            model.model.ckpt, model.model.iter = cali_ckpt, 0
        return model
    if cali_data is not None:
        cali_model(qnn=qnn, use_aq=opt.use_aq, path=opt.cali_save_path, running_stat=opt.running_stat, interval=256,
                   w_cali_data=cali_data, a_cali_data=cali_data, iters=20000, batch_size=32, w=0.01, asym=opt.asym,
                   warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
        model.model.diffusion_model = qnn
        return model
    if context_shape is not None:
        raise ValueError("quantize_ldm: calibrating a conditional UNet needs cali_data=(x_t, t, c) "
                         "(quant.data_generate.generate_cali_data_conditional)")
    logger.info("Generating calibration data...")
    per_step = 256
    xs, ts = generate_cali_data_ldm(qnn, T=opt.custom_steps, c=1, batch_size=per_step, shape=shape,
                                    plms=getattr(opt, "plms", False), eta=opt.eta, linear_start=model.linear_start,
                                    linear_end=model.linear_end, timesteps=model.num_timesteps)
    qnn._engine = None
    kept = [slice(i * per_step, (i + 1) * per_step) for i in range(0, opt.custom_steps, opt.interval_length)]
    w_cali_data = [torch.cat([xs[s] for s in kept]), torch.cat([ts[s] for s in kept])]
    logger.info("Calibration data generated.")
    cali_model(qnn=qnn, use_aq=opt.use_aq, path=opt.cali_save_path, running_stat=opt.running_stat, interval=per_step,
               w_cali_data=w_cali_data, a_cali_data=(xs, ts), iters=20000, batch_size=32, w=0.01, asym=opt.asym,
               warmup=0.2, opt_mode=RLOSS.MSE, multi_gpu=False)
    model.model.diffusion_model = qnn
    return model
