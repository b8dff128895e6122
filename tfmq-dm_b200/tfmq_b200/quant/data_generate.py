"""Calibration-data generation -- mirror of the reference's `quant/data_generate.py`.

The reference harvests (x_t, t[, c]) pairs by running the FULL-PRECISION sampler with `untill_fake_t` early stops
(data_generate.py:52-113; ddim.py:145-147, denoising.py:24-25).  Here that is the same step engine in its all-floating-
point state (every layer on the fp16-split tensor-core path), so data generation is accelerated by the same kernels.
`model` is the QuantModel (the reference passes the LatentDiffusion wrapper; text / class encoders are outside the hot path,
so the conditional variants take the conditioning TENSORS the reference would have obtained from
`model.get_learned_conditioning`)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from ..samplers import DDIMSampler, PLMSSampler, generalized_steps


def _cat(tmp) -> Tuple[torch.Tensor]:
    return tuple(torch.cat([x[i] for x in tmp]) for i in range(len(tmp[0])))


def _fp_state(model):
    model.set_quant_state(False, False)
    model._engine = None            # the engine freezes the quant state it was traced in


@torch.no_grad()
def generate_cali_data_ddim(model=None, betas: torch.Tensor = None, T: int = None, c: int = None, batch_size: int = None,
                            shape: List[int] = None, num_timesteps: int = 1000, generator: torch.Generator = None,
                            runnr=None) -> Tuple[torch.Tensor]:
    """reference :52-71 (`runnr.sample_image(x, model, untill_fake_t=i)[1:]` -> (x_t, t)).  Two call shapes: the
    reference's, `generate_cali_data_ddim(runnr=<runners.Diffusion>, model=qnn, T=, c=, batch_size=, shape=)`, which samples
    through the runner (its skip_type / eta / betas); or without a runner, `(model, betas, T, c, batch_size, shape)` on the
    uniform schedule.  `model` is the QuantModel; it is put in its all-floating-point state."""
    _fp_state(model)
    dev = next(model.parameters()).device
    if runnr is None and betas is None:
        raise ValueError("generate_cali_data_ddim: give the runner (runnr=) or the beta schedule (betas=)")
    seq = list(range(0, num_timesteps, num_timesteps // T)) if runnr is None else None
    tmp = []
    for i in range(1, T + 1):
        if i % c == 0:
            x = torch.randn((batch_size, *shape), device=dev, generator=generator)
            if runnr is not None:
                x_t, t_t = runnr.sample_image(x, model, untill_fake_t=i)[1:]
            else:
                _, _, x_t, t_t = generalized_steps(x, seq, model, betas, eta=0.0, untill_fake_t=i)
            tmp.append([x_t.cpu(), t_t.cpu()])
    model._engine = None                # the all-fp engine must not outlive the state it was traced in
    return _cat(tmp)


@torch.no_grad()
def generate_cali_data_ldm(model, T: int, c: int, batch_size: int, shape: List[int], plms: bool = False,
                           eta: float = 0.0, ddpm_time_num: int = 1000, x_T_fn=None, **sampler_kw) -> Tuple[torch.Tensor]:
    """reference :74-113: x_t after (t - 1) sampler steps, paired with the DDPM time the UNet sees next."""
    _fp_state(model)
    sampler = (PLMSSampler if plms else DDIMSampler)(model, **sampler_kw)
    tmp = []
    for t in range(1, T + 1):
        if t % c == 0:
            x_t, _ = sampler.sample(S=T, batch_size=batch_size, shape=shape, verbose=False, eta=eta, untill_fake_t=t,
                                    x_T=x_T_fn(t) if x_T_fn is not None else None)
            real_time = (T - t) * ddpm_time_num // T + 1
            tmp.append([x_t.cpu(), torch.full((batch_size,), real_time, dtype=torch.long)])
    return _cat(tmp)


@torch.no_grad()
def generate_cali_data_conditional(model, T: int, c: int, batch_size: int, shape: List[int],
                                   conditionings: Sequence[torch.Tensor], unconditional: torch.Tensor, scale: float,
                                   plms: bool = False, eta: float = 0.0, ddpm_time_num: int = 1000, x_T_fn=None,
                                   **sampler_kw) -> Tuple[torch.Tensor]:
    """reference :13-49 (text-guided, scale 7.5) and :116-154 (class-conditional ImageNet, scale 3.0): for every kept step
    and every conditioning, guided sampling up to that step; both (x_t, t, c) and (x_t, t, uc) are recorded."""
    _fp_state(model)
    sampler = (PLMSSampler if plms else DDIMSampler)(model, **sampler_kw)
    tmp = []
    for t in range(1, T + 1):
        if t % c == 0:
            for c_t in conditionings:
                x_t, _ = sampler.sample(S=T, batch_size=batch_size, shape=shape, conditioning=c_t, verbose=False, eta=eta,
                                        unconditional_guidance_scale=scale, unconditional_conditioning=unconditional,
                                        untill_fake_t=t, x_T=x_T_fn(t) if x_T_fn is not None else None)
                real_time = (T - t) * ddpm_time_num // T + 1
                t_t = torch.full((batch_size,), real_time, dtype=torch.long)
                tmp += [[x_t.cpu(), t_t, c_t.cpu()], [x_t.cpu(), t_t, unconditional.cpu()]]
    return _cat(tmp)
