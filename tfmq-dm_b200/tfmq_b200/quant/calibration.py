"""PTQ orchestration -- mirror of the reference's `quant/calibration.py` (cali_model,
load_cali_model, cali_model_multi, uaq2adar) with the same arguments, phases and checkpoint
format:  {'weight': qnn.state_dict(), 'act_0': {...}, ..., 'act_{T-1}': {...}}  where each act_k maps
'model.<layer>.aqtizer.delta' / '.zero_point' to 0-dim tensors (reference :99-106,147-154).
"""
from __future__ import annotations

import logging
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn

from .adaptive_rounding import RMODE, AdaRoundQuantizer
from .quant_block import BaseQuantBlock
from .quant_layer import QuantLayer, UniformAffineQuantizer
from .quant_model import QuantModel
from .reconstruction import block_reconstruction, layer_reconstruction, tib_reconstruction

logger = logging.getLogger(__name__)


def _dev(qnn: nn.Module) -> torch.device:
    return next(qnn.parameters()).device


def uaq2adar(model: nn.Module) -> None:
    """Swap every reconstructable layer's weight quantiser for an AdaRoundQuantizer (reference :19-42)."""
    for _, child in model.named_children():
        if isinstance(child, QuantLayer):
            if not child.ignore_recon:
                child.wqtizer = AdaRoundQuantizer(child.wqtizer, rmode=RMODE.LEARNED_HARD_SIGMOID,
                                                  w=child.original_w.data)
        elif isinstance(child, BaseQuantBlock):
            if not child.ignore_recon:
                for _, sub in child.named_modules():
                    if isinstance(sub, QuantLayer):
                        sub.wqtizer = AdaRoundQuantizer(sub.wqtizer, rmode=RMODE.LEARNED_HARD_SIGMOID,
                                                        w=sub.original_w.data)
        else:
            uaq2adar(child)


def _promote(module, with_delta: bool) -> None:
    zp = module.zero_point
    module.zero_point = nn.Parameter(zp if torch.is_tensor(zp) else torch.tensor(float(zp)))
    if with_delta:
        module.delta = nn.Parameter(module.delta)


def _demote(module, with_delta: bool) -> None:
    z = module.zero_point.data
    delattr(module, "zero_point")
    module.zero_point = z
    if with_delta:
        d = module.delta.data
        delattr(module, "delta")
        module.delta = d


def _reset_aqtizers(qnn: QuantModel) -> None:
    qnn.invalidate_engine()
    for name, module in qnn.model.named_modules():
        if "aqtizer" in name and isinstance(module, UniformAffineQuantizer):
            if module.delta is not None:
                del module.delta
                del module.zero_point
            module.delta = None
            module.zero_point = None
            module.init = False
            module._range_state = None


def _collect_act(qnn: QuantModel) -> dict:
    out = {}
    for name, module in qnn.model.named_modules():
        if "aqtizer" in name and isinstance(module, UniformAffineQuantizer) and module.delta is not None:
            out["model." + name + ".delta"] = module.delta.detach().reshape(()).cpu().clone()
            out["model." + name + ".zero_point"] = torch.as_tensor(module.zero_point).detach().reshape(()).cpu().float()
    return out


def cali_model(qnn: QuantModel, w_cali_data: Tuple[torch.Tensor], a_cali_data: Tuple[torch.Tensor],
               use_aq: bool = False, path: str = None, running_stat: bool = False, interval: int = 128,
               **kwargs) -> dict:
    """reference quant/calibration.py:45-155.  Runs inside `qnn.calibrating()`: every forward here is the torch module
    graph (hooks, autograd, lazy quantiser initialisation); the step engine is re-traced afterwards."""
    with qnn.calibrating():
        return _cali_model(qnn, w_cali_data, a_cali_data, use_aq=use_aq, path=path, running_stat=running_stat,
                           interval=interval, **kwargs)


def _cali_model(qnn: QuantModel, w_cali_data: Tuple[torch.Tensor], a_cali_data: Tuple[torch.Tensor],
                use_aq: bool = False, path: str = None, running_stat: bool = False, interval: int = 128,
                **kwargs) -> dict:
    """Phases of reference :45-155: W0 weight-quantiser init, W1 reconstruction (TIB / layers /
    blocks in definition order), W2 parameter promotion, A per-timestep activation ranges (FSC).
    Returns (and, if `path`, saves) the checkpoint dict."""
    logger.info("Calibrating...")
    dev = _dev(qnn)

    def recon_model(model: nn.Module, tag: bool = False) -> bool:
        for name, module in model.named_children():
            logger.info(f"block name: {name} quant: {isinstance(module, BaseQuantBlock)}")
            if name == "output_blocks":
                tag = True
            if name == "tib":
                continue
            if name in ("time_embed", "temb"):
                logger.info("Reconstruction for time embedding")
                tib_reconstruction(qnn.tib, cali_data=cali_data, **kwargs)
                continue
            if isinstance(module, QuantLayer):
                if not module.ignore_recon:
                    logger.info(f"Reconstruction for layer {name}")
                    layer_reconstruction(qnn, module, cali_data=cali_data, **kwargs)
            elif isinstance(module, BaseQuantBlock):
                if not module.ignore_recon:
                    logger.info(f"Reconstruction for block {name}")
                    block_reconstruction(qnn, module, cali_data=cali_data, **kwargs)
            else:
                tag = recon_model(module, tag=tag)
        return tag

    # ---- W0: weight quantiser initialisation (one forward with use_wq)
    cali_data = w_cali_data
    qnn.set_quant_state(use_wq=True, use_aq=False)
    bs = min(8, cali_data[0].shape[0])
    with torch.no_grad():
        qnn(*(x[:bs].to(dev) for x in cali_data))
    qnn.disable_out_quantization()

    # ---- W1: reconstruction
    recon_model(qnn)
    qnn.set_quant_state(use_wq=True, use_aq=False)
    if hasattr(qnn, "tib"):
        delattr(qnn, "tib")

    # ---- W2: promote (delta, zero_point) to parameters so they land in the state_dict
    for name, module in qnn.model.named_modules():
        if "wqtizer" in name and isinstance(module, (UniformAffineQuantizer, AdaRoundQuantizer)):
            _promote(module, with_delta=True)
    model_dict = {"weight": {k: v.detach().cpu().clone() for k, v in qnn.state_dict().items()}}

    # ---- A: Finite-Set Calibration of the activation ranges, one table per timestep interval
    if use_aq:
        qnn.eval()
        cali_data = a_cali_data
        for time in range(cali_data[0].shape[0] // interval):
            t_cali = tuple(x[time * interval:(time + 1) * interval] for x in cali_data)
            qnn.set_quant_state(use_wq=True, use_aq=True)
            _reset_aqtizers(qnn)
            bs = min(16, t_cali[0].shape[0])
            with torch.no_grad():
                inds = np.random.choice(t_cali[0].shape[0], bs, replace=False)
                qnn(*(x[inds].to(dev) for x in t_cali))
                if running_stat:
                    inds = np.arange(t_cali[0].shape[0])
                    np.random.shuffle(inds)
                    qnn.set_running_stat(True)
                    for i in range(0, t_cali[0].shape[0], bs):
                        qnn(*(x[inds[i:i + bs]].to(dev) for x in t_cali))
                    qnn.set_running_stat(False)
            model_dict[f"act_{time}"] = _collect_act(qnn)
        if path:
            torch.save(model_dict, path)
    logger.info("Calibration done.")
    return model_dict


def load_cali_model(qnn: QuantModel, init_data: Tuple[torch.Tensor], use_aq: bool = False, path: str = None,
                    ckpt: dict = None) -> None:
    """reference quant/calibration.py:158-224, inside `qnn.calibrating()` (its dummy forwards run the module graph)."""
    with qnn.calibrating():
        _load_cali_model(qnn, init_data, use_aq=use_aq, path=path, ckpt=ckpt)


def _load_cali_model(qnn: QuantModel, init_data: Tuple[torch.Tensor], use_aq: bool = False, path: str = None,
                     ckpt: dict = None) -> None:
    """reference :158-224: dummy forward to materialise the lazily created quantiser tensors,
    first / last layer exemptions, AdaRound detection by 'alpha' in the key, strict=False load of the
    'weight' part, optional second forward that creates the activation quantiser parameters."""
    logger.info("Loading calibration model...")
    dev = _dev(qnn)
    if ckpt is None:
        ckpt = torch.load(path, map_location="cpu")
    weight = dict(ckpt["weight"])
    init = tuple(x.to(dev) for x in init_data)
    qnn.set_quant_state(use_wq=True, use_aq=False)
    with torch.no_grad():
        qnn(*init)
    qnn.disable_out_quantization()
    if any("alpha" in k for k in weight):
        uaq2adar(qnn)
    for name, module in qnn.model.named_modules():
        if "wqtizer" in name and isinstance(module, (UniformAffineQuantizer, AdaRoundQuantizer)):
            _promote(module, with_delta=True)
    for key in [k for k in weight if "aqtizer" in k]:
        del weight[key]
    qnn.load_state_dict(weight, strict=False)
    qnn.set_quant_state(use_wq=True, use_aq=False)
    for name, module in qnn.model.named_modules():
        if "wqtizer" in name:
            if isinstance(module, AdaRoundQuantizer):
                _demote(module, with_delta=True)
            elif isinstance(module, UniformAffineQuantizer):
                _demote(module, with_delta=False)
    if use_aq:
        qnn.set_quant_state(use_wq=True, use_aq=True)
        with torch.no_grad():
            qnn(*init)
        for module in qnn.model.modules():
            if isinstance(module, UniformAffineQuantizer) and module.delta is not None \
                    and not isinstance(module.zero_point, nn.Parameter):
                _promote(module, with_delta=False)
    logger.info("Loading calibration model done.")


def act_tables_from_ckpt(ckpt: dict) -> list:
    """[act_0, act_1, ...] in step order from a calibrated checkpoint dict."""
    out, k = [], 0
    while f"act_{k}" in ckpt:
        out.append(ckpt[f"act_{k}"])
        k += 1
    return out


def cali_model_multi(gpu: int, dist_backend: str, world_size: int, dist_url: str, rank: int, ngpus_per_node: int,
                     model, use_aq: bool, path: str, w_cali_data, a_cali_data, interval: int, running_stat: bool,
                     kwargs: dict):
    """Data-parallel calibration, one process per GPU (reference :228-389): every rank takes a
    1/world slice of each per-timestep interval; AdaRound alpha gradients are SUM-all-reduced each
    iteration (one flat bucket per unit, reconstruction.py), activation deltas are averaged; rank 0 saves."""
    import torch.distributed as dist
    rank = rank * ngpus_per_node + gpu
    if not dist.is_initialized():
        dist.init_process_group(backend=dist_backend, init_method=dist_url, world_size=world_size, rank=rank)
    if dist_backend == "nccl":
        torch.cuda.set_device(gpu)
    dev = torch.device("cuda", gpu) if torch.cuda.is_available() else torch.device("cpu")
    qnn = QuantModel(model.to(dev), kwargs.pop("wq_params"), kwargs.pop("aq_params"), cali=True,
                     softmax_a_bit=kwargs.pop("softmax_a_bit", 8), aq_mode=kwargs.pop("aq_mode", [2])).to(dev)
    qnn.eval()

    def shard(data, per):
        from ..dist_utils import shard_interval_indices
        idx = shard_interval_indices(data[0].shape[0], per, rank, world_size)
        return tuple(x[idx] for x in data)

    w_shard = shard(w_cali_data, interval) if w_cali_data[0].shape[0] >= interval else w_cali_data
    a_shard = shard(a_cali_data, interval)
    kwargs = dict(kwargs, multi_gpu=True)
    ckpt = cali_model(qnn, w_shard, a_shard, use_aq=False, path=None, running_stat=running_stat,
                      interval=interval // world_size, **kwargs)
    if use_aq:
        with qnn.calibrating():
            ckpt = _fsc_multi(qnn, a_shard, interval // world_size, running_stat, ckpt)
    if rank == 0 and path:
        torch.save(ckpt, path)
    return ckpt          # (the reference returns None; every rank's dict is returned so callers can check rank consistency)


def _fsc_multi(qnn, a_cali_data, interval, running_stat, ckpt):
    dev = _dev(qnn)
    for time in range(a_cali_data[0].shape[0] // interval):
        t_cali = tuple(x[time * interval:(time + 1) * interval] for x in a_cali_data)
        qnn.set_quant_state(True, True)
        _reset_aqtizers(qnn)
        bs = min(16, t_cali[0].shape[0])
        with torch.no_grad():
            qnn(*(x[:bs].to(dev) for x in t_cali))
            if running_stat:
                qnn.set_running_stat(True)
                for i in range(0, t_cali[0].shape[0], bs):
                    qnn(*(x[i:i + bs].to(dev) for x in t_cali))
                qnn.set_running_stat(False)
        qnn.synchorize_activation_statistics()
        ckpt[f"act_{time}"] = _collect_act(qnn)
    return ckpt
