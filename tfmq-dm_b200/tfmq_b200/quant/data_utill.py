"""Input / output caching for reconstruction units -- mirror of the reference's
`quant/data_utill.py` (save_inout, GetLayerInpOut, DataSaverHook, StopForwardException)."""
from __future__ import annotations

import contextlib
import logging
from typing import Dict, Tuple, Union

import torch
import torch.nn as nn

from .quant_block import BaseQuantBlock, QuantBasicTransformerBlock
from .quant_layer import QuantLayer

logger = logging.getLogger(__name__)


class StopForwardException(Exception):
    pass


class DataSaverHook:
    """Forward hook (with kwargs) that keeps a unit's input / output and can abort the forward."""

    def __init__(self, store_input=False, store_output=False, stop_forward=False) -> None:
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = self.output_store = None

    def __call__(self, module: nn.Module, input_batch: Tuple, kwargs: Dict, output_batch) -> None:
        if self.store_input:
            self.input_store = input_batch[:-1] if isinstance(input_batch[-1], int) else input_batch
            if isinstance(module, QuantBasicTransformerBlock):
                self.input_store = (input_batch[0], kwargs["context"])
        if self.store_output:
            self.output_store = output_batch
        if self.stop_forward:
            raise StopForwardException


class GetLayerInpOut:
    """FP forward up to the unit (targets); with asym=True a second forward with the network
    quantised so far gives the inputs the unit will really see (reference :109-169)."""

    def __init__(self, model, layer: Union[QuantLayer, BaseQuantBlock], device, asym=False, use_aq=False) -> None:
        self.model, self.layer, self.device, self.asym, self.use_aq = model, layer, device, asym, use_aq
        self.data_saver = DataSaverHook(store_input=True, store_output=True, stop_forward=True)

    def _forward(self, args):
        # the hooked forward is a calibration forward: torch module graph (QuantModel.calibrating), never the step engine
        ctx = self.model.calibrating() if hasattr(self.model, "calibrating") else contextlib.nullcontext()
        with ctx:
            try:
                self.model(*(a.to(self.device) for a in args))
            except StopForwardException:
                pass

    def __call__(self, xs, ts, cs=None):
        args = (xs, ts) if cs is None else (xs, ts, cs)
        self.model.eval()
        self.model.set_quant_state(False, False)
        handle = self.layer.register_forward_hook(self.data_saver, with_kwargs=True)
        with torch.no_grad():
            self._forward(args)
            if self.asym:
                self.data_saver.store_output = False
                self.model.set_quant_state(use_wq=True, use_aq=self.use_aq)
                self._forward(args)
                self.data_saver.store_output = True
        handle.remove()
        self.model.set_quant_state(False, False)
        self.layer.set_quant_state(True, self.use_aq)
        self.model.train()
        ins = tuple(x.detach() for x in self.data_saver.input_store)
        out = self.data_saver.output_store
        outs = (out.detach(),) if isinstance(out, torch.Tensor) else tuple(x.detach() for x in out)
        return ins, outs


def save_inout(model, layer, cali_data: Tuple[torch.Tensor], asym=False, use_act=False, batch_size=128,
               keep_gpu=True):
    """Cache (inputs, FP outputs) of `layer` over the calibration set, kept on the GPU (180 GB of HBM
    make the reference's CPU staging, :62-67 `keep_gpu` heuristics, unnecessary; keep_gpu=False still
    stages on the host)."""
    device = next(model.parameters()).device
    get = GetLayerInpOut(model, layer, device, asym, use_act)
    ins, outs = None, None
    for i in range(0, cali_data[0].size(0), batch_size):
        a, b = get(*(d[i:i + batch_size] for d in cali_data))
        if ins is None:
            ins, outs = tuple([] for _ in a), tuple([] for _ in b)
        for j, v in enumerate(a):
            ins[j].append(v if keep_gpu else v.cpu())
        for j, v in enumerate(b):
            outs[j].append(v if keep_gpu else v.cpu())
    ins = tuple(torch.cat(x) for x in ins)
    outs = tuple(torch.cat(x) for x in outs)
    return ins, (outs[0] if len(outs) == 1 else outs)
