"""Sampler entry points with the reference's call shape, running on the fused step engine.

* generalized_steps(x, seq, model, b, **kwargs) -> (xs, x0_preds, xt, t)
      reference ddim/functions/denoising.py:10-41 (used by Diffusion.sample_image, runners/diffusion.py:429-476)
* DDIMSampler(model).sample(S, batch_size, shape, eta=0., x_T=None, ...) -> (samples, intermediates)
      reference ldm/models/diffusion/ddim.py:57-212, unconditional or conditional with classifier-free guidance
* PLMSSampler(model).sample(...)  -- same call shape, Adams-Bashforth multistep update
      reference ldm/models/diffusion/plms.py:57-242 (the README's Stable Diffusion command samples with --plms)

What changes underneath: the latent stays resident on the GPU for the whole trajectory (the reference
hops GPU<->CPU every step, denoising.py:23,38), the FSC parameter switch is one device-side row
copy (instead of `load_state_dict` + `.item()`), and the UNet step + DDIM update is one CUDA-graph replay.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch


def compute_alpha(beta: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def _randn(shape, device) -> torch.Tensor:
    """One standard-normal draw of a stochastic (eta > 0) step, from the same generator, in the same order and with the same
    shape as the reference's `torch.randn_like(x)` (denoising.py:36) / `noise_like(x.shape, device)` (ddim.py:209)."""
    return torch.randn(tuple(shape), device=device)


def ddim_coefficients(seq: Sequence[int], betas: torch.Tensor, eta: float = 0.0):
    """Per sampling step (sqrt(a_t), sqrt(1-a_t), sqrt(a_next), c2, c1) as fp32 values computed with the
    same tensor ops as the reference loop (denoising.py:19-36), so the device update reproduces it."""
    betas = betas.detach().float().cpu()
    rows = []
    seq_next = [-1] + list(seq[:-1])
    for i, j in zip(reversed(seq), reversed(seq_next)):
        at = compute_alpha(betas, torch.tensor([i])).reshape(())
        an = compute_alpha(betas, torch.tensor([j])).reshape(())
        c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
        c2 = ((1 - an) - c1 ** 2).sqrt()
        rows.append([at.sqrt().item(), (1 - at).sqrt().item(), an.sqrt().item(), c2.item(), float(c1)])
    return rows


def _act_tables(cali_ckpt, steps: int):
    if cali_ckpt is None:
        return None
    return [cali_ckpt[f"act_{k}"] for k in range(steps)]


@torch.no_grad()
def generalized_steps(x, seq, model, b, **kwargs):
    """Drop-in for ddim/functions/denoising.py:generalized_steps.  `model` is a QuantModel whose
    calibrated state is loaded; kwargs: eta, tot / cali_ckpt / t_max (FSC, as the reference passes them),
    untill_fake_t.  Returns (xs, x0_preds, xt, t) with xs[0] = x and xs[-1] the denoised sample
    (intermediates are kept only with keep_trajectory=True)."""
    eta = kwargs.get("eta", 0)
    dev = next(model.parameters()).device
    n = x.size(0)
    seq = list(seq)
    steps = len(seq)
    fsc = kwargs.get("cali_ckpt") if kwargs.get("tot") is not None else None
    eng = getattr(model, "_engine", None)
    if eng is None or eng.batch != n:
        eng = model.build_engine(batch=n)
    eng.set_schedule(list(reversed(seq)), _act_tables(fsc, steps), ddim_coefficients(seq, b, eta))
    keep = kwargs.get("keep_trajectory", False)
    stop = kwargs.get("untill_fake_t")
    xs, x0_preds = [x], []
    eng.x_in.copy_(x.to(dev))
    xt = t = None
    for k in range(steps):
        t = torch.ones(n, device=dev) * list(reversed(seq))[k]
        if keep or (stop is not None and k == stop - 1):
            xt = eng.x_in.clone()
        if stop is not None and k == stop - 1:
            break
        eng.set_noise(_randn(x.shape, dev) if eta != 0 else None)      # c1 * randn_like(x), denoising.py:36
        eng.step(k)
        if keep:
            xs.append(eng.x_in.to("cpu"))
            x0_preds.append(eng.x0_pred.to("cpu"))
    if not keep:
        xs.append(eng.x_in.to("cpu"))
        x0_preds.append(eng.x0_pred.to("cpu"))
    return xs, x0_preds, xt, t


def make_beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2):
    """ldm/modules/diffusionmodules/util.py:20-43 ('linear' = linspace of sqrt(beta), squared)."""
    if schedule != "linear":
        raise NotImplementedError(schedule)
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def make_ddim_timesteps(num_ddim_timesteps: int, num_ddpm_timesteps: int = 1000) -> np.ndarray:
    """uniform discretisation, +1 (util.py:46-60)."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class DDIMSampler:
    """DDIM sampling for the LDM UNets, unconditional or conditional with classifier-free guidance
    (ldm/models/diffusion/ddim.py:57-212).
    `model` is the QuantModel (what the reference installs as model.model.diffusion_model);
    the noise schedule comes from linear_start / linear_end of the LDM config."""

    def __init__(self, model, linear_start: Optional[float] = None, linear_end: Optional[float] = None,
                 timesteps: Optional[int] = None, ckpt: Optional[dict] = None):
        wrapper = getattr(model, "model", None)
        if wrapper is not None and hasattr(wrapper, "diffusion_model"):
            # the reference's call shape, DDIMSampler(<LatentDiffusion>) (runners.LatentDiffusion here): the UNet is
            # model.model.diffusion_model, the FSC tables are the `.ckpt` the script installs on the wrapper, the noise
            # schedule is the model's
            ckpt = ckpt if ckpt is not None else getattr(wrapper, "ckpt", None)
            linear_start = linear_start if linear_start is not None else getattr(model, "linear_start", None)
            linear_end = linear_end if linear_end is not None else getattr(model, "linear_end", None)
            timesteps = timesteps if timesteps is not None else getattr(model, "num_timesteps", None)
            model = wrapper.diffusion_model
        linear_start = 0.0015 if linear_start is None else linear_start
        linear_end = 0.0195 if linear_end is None else linear_end
        timesteps = 1000 if timesteps is None else timesteps
        self.model = model
        self.ddpm_num_timesteps = timesteps
        betas = make_beta_schedule("linear", timesteps, linear_start, linear_end)
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.ckpt = ckpt         # FSC tables: {'act_k': ...} as saved by cali_model
        # FSC indexing constants as the scripts install them on the wrapper (sample_diffusion_ldm.py:473-479)
        self.tot = getattr(wrapper, "tot", None) if wrapper is not None else None
        self.t_max = getattr(wrapper, "t_max", None) if wrapper is not None else None

    def fsc_tables(self, ts):
        """The `act_k` dict DiffusionWrapper.forward would load before the UNet call at timestep t
        (ldm/models/diffusion/ddpm.py:1402-1405): k = t_max - (t - 1) // tot, where the scripts set tot = 1000 // n_tables and
        t_max = n_tables - 1 from the NUMBER OF CALIBRATED TABLES (sample_diffusion_ldm.py:475-477), not from S."""
        if self.ckpt is None:
            return None
        tot, t_max = self.tot, self.t_max
        if tot is None or t_max is None:
            n_tables = sum(1 for k in self.ckpt if str(k).startswith("act_"))
            tot, t_max = self.ddpm_num_timesteps // n_tables, n_tables - 1
        return [self.ckpt[f"act_{int(t_max - (int(t) - 1) // tot)}"] for t in ts]

    def make_schedule(self, ddim_num_steps: int, ddim_eta: float = 0.0):
        self.ddim_timesteps = make_ddim_timesteps(ddim_num_steps, self.ddpm_num_timesteps)
        ac = self.alphas_cumprod
        alphas = ac[self.ddim_timesteps]
        alphas_prev = np.asarray([ac[0]] + ac[self.ddim_timesteps[:-1]].tolist())
        sigmas = ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        f32 = lambda a: torch.from_numpy(np.asarray(a)).to(torch.float32)  # noqa: E731
        self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sigmas = f32(alphas), f32(alphas_prev), f32(sigmas)
        self.ddim_sqrt_one_minus_alphas = f32(np.sqrt(1.0 - alphas))

    def coefficient_rows(self):
        """p_sample_ddim's per-index scalars (ddim.py:196-211) in sampling order (index S-1 .. 0)."""
        rows = []
        for index in reversed(range(len(self.ddim_timesteps))):
            a_t, a_prev = self.ddim_alphas[index], self.ddim_alphas_prev[index]
            sigma = self.ddim_sigmas[index]
            rows.append([a_t.sqrt().item(), self.ddim_sqrt_one_minus_alphas[index].item(), a_prev.sqrt().item(),
                         (1.0 - a_prev - sigma ** 2).sqrt().item(), sigma.item(), 1.0])   # 1.0: (x0 term + dir_xt) + noise
        return rows

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, eta=0.0, x_T=None, verbose=False,
               unconditional_guidance_scale=1.0, unconditional_conditioning=None, untill_fake_t=None, **kwargs):
        self.make_schedule(S, eta)
        if not hasattr(self.model, "build_engine"):
            return self._sample_callable(S, batch_size, shape, conditioning, x_T, unconditional_guidance_scale,
                                         unconditional_conditioning, untill_fake_t, eta=eta)
        dev = next(self.model.parameters()).device
        C, H, W = shape
        img = torch.randn((batch_size, C, H, W), device=dev) if x_T is None else x_T.to(dev)
        # conditional sampling (ddim.py:171-180): with guidance the UNet sees [x | x], [uncond | cond] as one 2B batch
        cfg = conditioning is not None and unconditional_conditioning is not None and unconditional_guidance_scale != 1.0
        ctx = None
        if conditioning is not None:
            ctx = conditioning.to(dev).float()
            if cfg:
                ctx = torch.cat([unconditional_conditioning.to(dev).float(), ctx], dim=0)
        nb = batch_size * (2 if cfg else 1)
        eng = getattr(self.model, "_engine", None)
        want_ctx = tuple(ctx.shape) if ctx is not None else None
        have_ctx = tuple(eng.ctx_in.shape) if (eng is not None and eng.ctx_in is not None) else None
        if eng is None or eng.batch != nb or want_ctx != have_ctx:
            eng = self.model.build_engine(batch=nb, context_shape=ctx.shape[1:] if ctx is not None else None)
        eng.set_guidance(float(unconditional_guidance_scale) if cfg else None)
        if ctx is not None:
            eng.ctx_in.copy_(ctx)
        ts = [float(t) for t in np.flip(self.ddim_timesteps)]
        tables = self.fsc_tables(ts)
        eng.set_schedule(ts, tables, self.coefficient_rows())
        eng.x_in.copy_(torch.cat([img, img], dim=0) if cfg else img)
        n_run = S if not untill_fake_t else min(S, untill_fake_t - 1)
        for k in range(n_run):
            # sigma_t * noise_like(x.shape, device) (ddim.py:209): one draw per step, of the un-doubled batch
            eng.set_noise(_randn((batch_size, C, H, W), dev) if eta != 0 else None)
            eng.step(k)
        out = eng.x_in[:batch_size].clone()
        return out, {"x_inter": [img, out], "pred_x0": [img, eng.x0_pred[:batch_size].clone()]}


def _callable_eps(model, batch_size, ctx, cfg, scale):
    """eps(x, t) from any callable (x, t, context) -> eps on CUDA tensors, with the guidance formula of ddim.py:171-180."""
    from . import ops

    def fn(x, t, k):
        tt = torch.full((x.shape[0],), float(t), device=x.device)
        if not cfg:
            return model(x, tt, ctx)
        e = model(torch.cat([x, x]), torch.cat([tt, tt]), ctx)
        out = torch.empty_like(e[:batch_size])
        ops.cfg_combine(e[:batch_size].contiguous(), e[batch_size:].contiguous(), scale, out)
        return out
    return fn


def _sample_callable(self, S, batch_size, shape, conditioning, x_T, scale, uncond, untill_fake_t, eta=0.0):
    """DDIM with a plain callable as the UNet (update-rule parity tests): same kernels for guidance and the update."""
    from . import ops
    dev = x_T.device
    img = x_T.clone()
    cfg = conditioning is not None and uncond is not None and scale != 1.0
    ctx = None
    if conditioning is not None:
        ctx = torch.cat([uncond.to(dev).float(), conditioning.to(dev).float()]) if cfg else conditioning.to(dev).float()
    eps_fn = _callable_eps(self.model, batch_size, ctx, cfg, float(scale))
    rows = torch.tensor(self.coefficient_rows(), dtype=torch.float32, device=dev)
    ts = [float(t) for t in np.flip(self.ddim_timesteps)]
    x0 = torch.empty_like(img)
    for k in range(S if not untill_fake_t else min(S, untill_fake_t - 1)):
        e = eps_fn(img, ts[k], k).contiguous()
        ops.ddim_update(img, e, rows[k], img, x0, noise=_randn(img.shape, dev) if eta != 0 else None)
    return img, {"x_inter": [x_T, img], "pred_x0": [x_T, x0]}


DDIMSampler._sample_callable = _sample_callable


class PLMSSampler(DDIMSampler):
    """Pseudo linear multistep sampling (ldm/models/diffusion/plms.py:57-242), eta = 0: the UNet output of a step is
    combined with up to three stored ones (Adams-Bashforth, `tfmq_plms_eps`) before the DDIM-form update; the first
    step is the pseudo improved Euler step with a second UNet evaluation at x_prev.  Same schedule, FSC indexing,
    conditioning and guidance handling as DDIMSampler.  `model` is a QuantModel, or any callable
    (x, t, context) -> eps on CUDA tensors (used by the update-rule parity test)."""

    def _eps_fn(self, batch_size, ctx, cfg, scale):
        from . import ops
        model = self.model
        if hasattr(model, "build_engine"):
            nb = batch_size * (2 if cfg else 1)
            eng = getattr(model, "_engine", None)
            want = tuple(ctx.shape) if ctx is not None else None
            have = tuple(eng.ctx_in.shape) if (eng is not None and eng.ctx_in is not None) else None
            if eng is None or eng.batch != nb or want != have:
                eng = model.build_engine(batch=nb, context_shape=ctx.shape[1:] if ctx is not None else None)
            eng.set_guidance(None)
            self._eng = eng

            def fn(x, t, k):
                if self._tables is not None:
                    eng.select_step(k)
                e = eng.forward(torch.cat([x, x]) if cfg else x, torch.full((nb,), float(t)), ctx)
                if not cfg:
                    return e
                out = torch.empty_like(e[:batch_size])
                ops.cfg_combine(e[:batch_size], e[batch_size:], scale, out)
                return out
            return fn

        return _callable_eps(model, batch_size, ctx, cfg, scale)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, eta=0.0, x_T=None, verbose=False,
               unconditional_guidance_scale=1.0, unconditional_conditioning=None, untill_fake_t=None, **kwargs):
        from . import ops
        if eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS (plms.py:27-28)")
        self.make_schedule(S, eta)
        dev = x_T.device if x_T is not None and x_T.is_cuda else next(self.model.parameters()).device
        C, H, W = shape
        img = torch.randn((batch_size, C, H, W), device=dev) if x_T is None else x_T.to(dev).clone()
        cfg = conditioning is not None and unconditional_conditioning is not None and unconditional_guidance_scale != 1.0
        ctx = None
        if conditioning is not None:
            ctx = conditioning.to(dev).float()
            if cfg:
                ctx = torch.cat([unconditional_conditioning.to(dev).float(), ctx], dim=0)
        ts = [float(t) for t in np.flip(self.ddim_timesteps)]
        self._tables = self.fsc_tables(ts)
        eps_fn = self._eps_fn(batch_size, ctx, cfg, float(unconditional_guidance_scale))
        if self._tables is not None and hasattr(self.model, "build_engine"):
            self._eng.set_schedule(ts, self._tables)
        rows = torch.tensor(self.coefficient_rows(), dtype=torch.float32, device=dev)   # [S, 5] in sampling order
        old = []                                     # newest first
        x0 = torch.empty_like(img)
        e_comb = torch.empty_like(img)
        n_run = S if not untill_fake_t else min(S, untill_fake_t - 1)
        for k in range(n_run):
            e_t = eps_fn(img, ts[k], k).contiguous()
            if not old:
                # pseudo improved Euler: x_prev from e_t, second evaluation there, average
                x_prev = torch.empty_like(img)
                ops.ddim_update(img, e_t, rows[k], x_prev)
                e_next = eps_fn(x_prev, ts[min(k + 1, S - 1)], min(k + 1, S - 1)).contiguous()
                ops.plms_eps(e_t, ["euler", e_next], e_comb)
            else:
                ops.plms_eps(e_t, old[:3], e_comb)
            ops.ddim_update(img, e_comb, rows[k], img, x0)
            old = [e_t] + old[:2]
        return img, {"x_inter": [x_T, img], "pred_x0": [x_T, x0]}
