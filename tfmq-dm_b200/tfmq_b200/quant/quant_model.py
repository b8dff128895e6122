"""QuantModel -- mirror of the reference's `quant/quant_model.py:10-160`: module surgery that turns
the FP UNet's Conv2d / Linear leaves into QuantLayers and its residual / attention blocks into
quantised blocks, quant-state toggles and the first / last layer exemptions.

New here: once calibrated parameters are in place (`load_cali_model` / `cali_model`), `forward`
runs the fused sm_100a step engine (engine.py) instead of the module graph.
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.nn as nn

from .quant_block import (BaseQuantBlock, QuantAttentionBlock, QuantAttnBlock, QuantBasicTransformerBlock,
                          QuantQKMatMul, QuantResBlock, QuantResnetBlock, QuantSMVMatMul,
                          QuantTemporalInformationBlock, QuantTemporalInformationBlockDDIM, b2qb)
from .quant_layer import QMODE, QuantLayer, StraightThrough


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, wq_params: dict = {}, aq_params: dict = {}, cali: bool = True,
                 **kwargs) -> None:
        super().__init__()
        self.model = model
        self.softmax_a_bit = kwargs.get("softmax_a_bit", 8)
        self.in_channels = model.in_channels
        if hasattr(model, "image_size"):
            self.image_size = model.image_size
        self.B = b2qb(aq_params["leaf_param"])
        self.quant_module(self.model, wq_params, aq_params, aq_mode=kwargs.get("aq_mode", [QMODE.NORMAL.value]),
                          prev_name=None)
        self.quant_block(self.model, wq_params, aq_params)
        if cali:
            self.get_tib(self.model, wq_params, aq_params)
        self._engine = None   # fused step engine, built on demand by build_engine() / forward()
        self._cali_depth = 0  # > 0 inside `with qnn.calibrating():` -- forward runs the torch module graph
        self._init_ok = False # every quantiser in use has been initialised (checked lazily, reset with the engine)

    # ------------------------------------------------------------------ surgery
    def get_tib(self, module: nn.Module, wq_params: dict = {}, aq_params: dict = {}):
        for name, child in module.named_children():
            if name == "temb":
                self.tib = QuantTemporalInformationBlockDDIM(child, aq_params, self.model.ch)
            elif name == "time_embed":
                self.tib = QuantTemporalInformationBlock(child, aq_params, self.model.model_channels, None)
            elif isinstance(child, QuantResBlock):
                self.tib.add_emb_layer(child.emb_layers)
            elif isinstance(child, QuantResnetBlock):
                self.tib.add_temb_proj(child.temb_proj)
            else:
                self.get_tib(child, wq_params, aq_params)

    def quant_module(self, module: nn.Module, wq_params: dict = {}, aq_params: dict = {},
                     aq_mode: List[int] = [QMODE.NORMAL.value], prev_name: str = None) -> None:
        """Wrap Conv2d / Linear leaves unless the attribute name marks a skip / shortcut / `op`
        (down-sampling) conv or DDIM's `downsample.conv`; `emb_layers.1` and `temb_proj` become
        quant_emb layers (they belong to the TIB).  Same predicate, operator precedence included,
        as reference :56-66."""
        for name, child in module.named_children():
            wrap = isinstance(child, tuple(QuantLayer.QMAP.keys())) and "skip" not in name and "op" not in name \
                and not (prev_name == "downsample" and name == "conv") and "shortcut" not in name
            if wrap:
                is_emb = (prev_name is not None and "emb_layers" in prev_name and "1" in name) or "temb_proj" in name
                setattr(module, name, QuantLayer(child, wq_params, aq_params, aq_mode=aq_mode, quant_emb=bool(is_emb)))
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module(child, wq_params, aq_params, aq_mode=aq_mode, prev_name=name)

    def quant_block(self, module: nn.Module, wq_params: dict = {}, aq_params: dict = {}) -> None:
        for name, child in module.named_children():
            cls = self.B.get(child.__class__.__name__)
            if cls is None:
                self.quant_block(child, wq_params, aq_params)
            elif cls in (QuantBasicTransformerBlock, QuantAttnBlock):
                setattr(module, name, cls(child, aq_params, softmax_a_bit=self.softmax_a_bit))
            elif cls in (QuantResnetBlock, QuantAttentionBlock, QuantResBlock):
                setattr(module, name, cls(child, aq_params))
            elif cls is QuantSMVMatMul:
                setattr(module, name, cls(aq_params, softmax_a_bit=self.softmax_a_bit))
            elif cls is QuantQKMatMul:
                setattr(module, name, cls(aq_params))

    # ------------------------------------------------------------------ state
    def invalidate_engine(self) -> None:
        """The step engine freezes weights, alpha and the quant state of every layer at trace time: anything that changes
        them drops it, and the next sampling forward re-traces."""
        self._engine = None
        self._init_ok = False

    @contextlib.contextmanager
    def calibrating(self):
        """Inside this context `forward` executes the torch module graph (hooks fire, autograd works, quantisers
        initialise lazily and track running statistics): the calibration-time graph of quant/calibration.py,
        quant/reconstruction.py and quant/data_utill.py.  Leaving it drops the step engine (the state has changed)."""
        self._cali_depth += 1
        # the calibration-time graph borrows torch's conv / matmul kernels: keep them in true fp32 (torch enables TF32 for
        # cuDNN convolutions by default), so that weight-quantiser initialisation, cached block inputs / outputs and the
        # FSC activation ranges are those of the fp32 reference path
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            yield self
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
            self._cali_depth -= 1
            self.invalidate_engine()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """FSC swaps `act_k` (aqtizer delta / zero_point only) per timestep (ddim/functions/denoising.py:26-29): the engine
        just re-reads the activation parameters; any other key changes what was frozen at trace time."""
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        if self._engine is not None:
            if all(".aqtizer." in k for k in state_dict):
                self._engine._load_current_aq()
            else:
                self.invalidate_engine()
        return out

    def _apply(self, fn, *a, **kw):
        self.invalidate_engine()
        return super()._apply(fn, *a, **kw)

    def _uninitialised(self) -> bool:
        """True while a quantiser that the current state uses has not seen data yet: that forward IS the reference's lazy
        initialisation (quant/quant_layer.py:213-216) and runs through the modules."""
        for m in self.model.modules():
            if isinstance(m, QuantLayer):
                if m.use_wq and not getattr(m.wqtizer, "init", True):
                    return True
                if m.use_aq and not m.disable_aq and not getattr(m.aqtizer, "init", True):
                    return True
        return False

    def set_quant_state(self, use_wq: bool = False, use_aq: bool = False) -> None:
        self.invalidate_engine()
        for m in self.model.modules():
            if isinstance(m, (BaseQuantBlock, QuantLayer)):
                m.set_quant_state(use_wq=use_wq, use_aq=use_aq)

    def quant_layers(self) -> List[QuantLayer]:
        return [m for m in self.model.modules() if isinstance(m, QuantLayer)]

    def disable_out_quantization(self) -> None:
        """First / last layer exemptions in module-enumeration order (reference :103-120):
        layers #0, #2 and the last stay fp for good; #1 and #3 keep weight quant but no act quant."""
        self.invalidate_engine()
        ql = self.quant_layers()
        for i in (0, 2, -1):
            ql[i].use_wq = False
            ql[i].disable_aq = True
            ql[i].ignore_recon = True
        ql[1].disable_aq = True
        ql[3].disable_aq = True

    def set_grad_ckpt(self, grad_ckpt: bool) -> None:
        for _, module in self.model.named_modules():
            if isinstance(module, QuantBasicTransformerBlock) or module.__class__.__name__ == "BasicTransformerBlock":
                module.checkpoint = grad_ckpt

    def synchorize_activation_statistics(self) -> None:
        """Average every initialised activation delta over the ranks: one all-reduce of a
        [n_layers] vector instead of the reference's one call per layer (reference :127-132)."""
        from ..dist_utils import allaverage_
        allaverage_([m.aqtizer.delta for m in self.modules()
                     if isinstance(m, QuantLayer) and m.aqtizer.delta is not None])

    def set_running_stat(self, running_stat: bool = False) -> None:
        self.invalidate_engine()
        for m in self.model.modules():
            if isinstance(m, QuantBasicTransformerBlock):
                for attn in (m.attn1, m.attn2):
                    for q in (attn.aqtizer_q, attn.aqtizer_k, attn.aqtizer_v, attn.aqtizer_w):
                        q.running_stat = running_stat
            elif isinstance(m, QuantQKMatMul):
                m.aqtizer_q.running_stat = m.aqtizer_k.running_stat = running_stat
            elif isinstance(m, QuantSMVMatMul):
                m.aqtizer_v.running_stat = m.aqtizer_w.running_stat = running_stat
            elif isinstance(m, QuantAttnBlock):
                for q in (m.aqtizer_q, m.aqtizer_k, m.aqtizer_v, m.aqtizer_w):
                    q.running_stat = running_stat
            elif isinstance(m, QuantLayer):
                m.set_running_stat(running_stat)

    # ------------------------------------------------------------------ forward
    def build_engine(self, batch: int, act_tables=None, timesteps=None, fp_passes: int = 3, context_shape=None,
                     fp_mode: str = "h16"):
        """Freeze the calibrated model into the fused sm_100a step program (engine.StepEngine).
        context_shape = (tokens, context_dim) for SpatialTransformer UNets."""
        from ..engine import StepEngine
        self._engine = StepEngine(self, batch=batch, act_tables=act_tables, timesteps=timesteps, fp_passes=fp_passes,
                                  context_shape=context_shape, fp_mode=fp_mode)
        return self._engine

    def forward(self, x: torch.Tensor, timestep=None, context: torch.Tensor = None) -> torch.Tensor:
        """Sampling forward = the fused sm_100a step engine, (re)built for this batch / context shape when needed.
        The torch module graph runs only where the reference's semantics need modules: inside `calibrating()`, under
        autograd (reconstruction), and for the lazy-initialisation forward of un-initialised quantisers.  There is no
        other route: a CPU tensor outside those cases raises (no CPU path, no eager fallback)."""
        module_graph = self._cali_depth > 0 or torch.is_grad_enabled()
        if not module_graph and not self._init_ok:
            self._init_ok = not self._uninitialised()
            module_graph = not self._init_ok
        if module_graph:
            self.invalidate_engine()
            if context is None:
                return self.model(x, timestep)
            return self.model(x, timestep, context)
        if not x.is_cuda:
            raise RuntimeError("QuantModel.forward outside calibration runs the sm_100a step engine: the input must be a "
                               "CUDA tensor (no CPU path; use `with qnn.calibrating():` for the torch module graph)")
        eng = self._engine
        want_ctx = tuple(context.shape[1:]) if context is not None else None
        if eng is not None:
            have_ctx = tuple(eng.ctx_in.shape[1:]) if eng.ctx_in is not None else None
            if eng.batch != x.shape[0] or have_ctx != want_ctx:
                eng = None
        if eng is None:
            eng = self.build_engine(batch=x.shape[0], context_shape=want_ctx)
        return eng.forward(x, timestep, context)
