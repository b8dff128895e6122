"""AdaRound / TIAR reconstruction -- mirror of the reference's `quant/reconstruction.py`
(layer_reconstruction, block_reconstruction, tib_reconstruction: same signatures and semantics).

Inner loop (reference :182-198), re-designed for the GPU: per iteration
  1. soft-rounded weights of every layer of the unit are materialised by one kernel each
     (ops.adaround_soft) and injected as autograd leaves,
  2. the unit's forward / backward gives dL/dW_soft for the batch (torch autograd over the block graph),
  3. one fused kernel per layer (ops.adaround_step) applies the chain rule to alpha, adds the rounding
     regulariser's gradient, performs the Adam update in place and reduces the regulariser value
     -- no optimiser object, no per-iteration host sync; the loss is read back only when it is logged.
Multi-GPU: dL/dW_soft of all layers is SUM-all-reduced as ONE flat bucket per iteration (the reference
all-reduces alpha.grad per tensor, `linklink.allreduce` = SUM; summing alpha.grad also multiplies the
regulariser by world_size, which is reproduced through lambda * world_size).
"""
from __future__ import annotations

import contextlib
import logging
import os
from typing import List

import torch

from .. import ops
from .adaptive_rounding import RMODE, AdaRoundQuantizer
from .data_utill import save_inout
from .quant_block import BaseQuantBlock
from .quant_layer import QuantLayer, StraightThrough, lp_loss
from .reconstruction_util import RLOSS, LinearTempDecay, print_freq, unit_layers

logger = logging.getLogger(__name__)

# Test hook: when set to a list, every iteration appends its total loss (reconstruction + rounding regulariser,
# what the reference's LossFunc returns) -- one host sync per iteration, so it stays None outside the parity tests.
LOSS_TRACE = None

# Bench hook: called with the iteration number at the end of every iteration (tools/bench_calibration.py stamps two of them to
# time the steady state without the caching before the loop and the graph capture of its first iteration).
ITER_HOOK = None

# The iteration body (soft weights, unit forward, loss, backward) as one captured CUDA graph.  TFMQ_RECON_GRAPH=0 runs it eagerly
# (debugging; same kernels, same order, same results).
RECON_GRAPH = os.environ.get("TFMQ_RECON_GRAPH", "1") != "0"


class _LayerState:
    def __init__(self, layer: QuantLayer):
        self.layer = layer
        layer.wqtizer = AdaRoundQuantizer(uaqtizer=layer.wqtizer, rmode=RMODE.LEARNED_HARD_SIGMOID,
                                          w=layer.original_w.data)
        layer.wqtizer.soft_tgt = True
        q = layer.wqtizer
        dev = layer.w.device
        self.cout = layer.w.shape[0]
        self.w2d = layer.w.detach().reshape(self.cout, -1).contiguous().float()
        self.delta = q.delta.detach().reshape(-1).float().to(dev).contiguous()
        zp = q.zero_point
        zp = zp.detach().reshape(-1).float().to(dev) if torch.is_tensor(zp) else torch.full_like(self.delta, float(zp))
        self.zp = zp.expand_as(self.delta).contiguous()
        self.alpha = q.alpha                        # nn.Parameter, updated in place by the kernel
        assert self.alpha.is_contiguous()
        self.m = torch.zeros_like(self.w2d)
        self.v = torch.zeros_like(self.w2d)
        self.level = q.level
        self.w_soft = torch.empty_like(self.w2d)

    def materialise(self):
        ops.adaround_soft(self.w2d, self.delta, self.zp, self.alpha.data, self.level, self.w_soft)
        leaf = self.w_soft.view_as(self.layer.w).detach().requires_grad_(True)
        self.layer.w_override = leaf
        return leaf

    def finish(self):
        self.layer.w_override = None
        self.layer.wqtizer.soft_tgt = False


def _adaround_loop(forward, layers: List[QuantLayer], cached_inputs, cached_outputs, batch_size, iters, w, b_range,
                   warmup, multi_gpu, p):
    if p != 2.0:
        raise NotImplementedError("reconstruction: only the p=2 loss of the entry points is implemented")
    states = [_LayerState(m) for m in layers]
    if not states:
        return
    from .tc_autograd import trim_buffers
    trim_buffers()
    dev = states[0].w2d.device
    decay = LinearTempDecay(iters, warmup, b_range[0], b_range[1])
    loss_start = iters * warmup
    world = 1
    if multi_gpu:
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_initialized() else 1
    round_acc = torch.zeros(1, device=dev)
    tuple_out = isinstance(cached_outputs, (tuple, list))
    n = cached_inputs[0].size(0)
    tgt_src = list(cached_outputs) if tuple_out else [cached_outputs]
    bs = min(batch_size, n)
    # resident batch buffers: the sampled rows of the cached inputs / targets are gathered into them every iteration, so the
    # forward / loss / backward of the unit is one fixed sequence of launches on fixed addresses -- captured ONCE as a CUDA graph
    # and replayed (the module graph + autograd engine cost ~3x the GPU time of an iteration in host time when run eagerly)
    cur_in = tuple(torch.empty((bs,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev) for x in cached_inputs)
    tgts = [torch.empty((bs,) + tuple(t_.shape[1:]), dtype=t_.dtype, device=dev) for t_ in tgt_src]
    rec = torch.zeros(1, device=dev)

    def gather(it_idx):
        for buf, x in zip(cur_in, cached_inputs):
            buf.copy_(x[it_idx.to(x.device)], non_blocking=True)
        for buf, t_ in zip(tgts, tgt_src):
            buf.copy_(t_[it_idx.to(t_.device)], non_blocking=True)

    def body():
        """soft weights -> unit forward -> reconstruction loss and its output gradient -> dL/dW_soft of every layer"""
        leaves = [s.materialise() for s in states]
        out = forward(*cur_in)
        # reconstruction loss and its gradient w.r.t. the unit's outputs in one kernel per output (tfmq_rec_loss: lp_loss with
        # p = 2, `.sum(1).mean()`, quant/quant_layer.py:146-156; LossFuncTimeEmbedding sums it over the TIB's outputs)
        outs = list(out) if tuple_out else [out]
        rec.zero_()
        gouts = []
        for o_, t_ in zip(outs, tgts):
            o_c, t_c = o_.detach().contiguous(), t_.contiguous()
            g_ = torch.empty_like(o_c)
            ops.rec_loss(o_c, t_c, o_c.numel() // o_c.shape[1], rec, g_)
            gouts.append(g_)
        grads = torch.autograd.grad(outs, leaves, grad_outputs=gouts, allow_unused=True)
        return [g.contiguous() if g is not None else torch.zeros_like(l_) for g, l_ in zip(grads, leaves)]

    # (random ops of the unit -- the reference leaves the model in train() mode after caching, so CIFAR's ResnetBlocks draw
    # dropout masks, data_utill.py:72 -- are captured with torch's graph-safe generator state: every replay draws fresh masks)
    graph, graph_grads = None, None
    for it in range(1, iters + 1):
        idx = torch.randperm(n)[:batch_size]
        gather(idx)
        if not RECON_GRAPH or dev.type != "cuda":
            grads = body()
        elif graph is None:
            # warm-up on a side stream (lazy module / attribute initialisation, cached plane buffers), then capture; the body
            # does not change the optimiser state, so running it more than once for the first batch is harmless
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_grads = body()
            graph.replay()
            grads = graph_grads
        else:
            graph.replay()
            grads = graph_grads
        if world > 1:
            from ..dist_utils import allreduce_flat_
            grads = allreduce_flat_(grads)
        b = decay(it) if it >= loss_start else 0.0
        log = it % print_freq == 0
        if log or LOSS_TRACE is not None:
            round_acc.zero_()
        for s, g in zip(states, grads):
            ops.adaround_step(s.w2d, s.delta, s.zp, s.alpha.data, g.contiguous(), s.m, s.v, s.level, it, 1e-3,
                              float(b), float(w) * world, round_acc)
        if LOSS_TRACE is not None:
            LOSS_TRACE.append(float(rec) + float(round_acc) / world)
        if ITER_HOOK is not None:
            ITER_HOOK(it)
        if log:
            logger.info("Total loss:\t{:.8f} (rec:{:.8f}, round:{:.8f})\tb={:.2f}\tcount={}".format(
                float(rec) + float(round_acc) / world, float(rec), float(round_acc) / world, b, it))
    del graph, graph_grads
    for s in states:
        s.finish()


@contextlib.contextmanager
def _plain_convs_on_own_kernels(unit):
    """The convolutions of a unit that are NOT QuantLayers (a ResBlock's 1x1 skip_connection: shortcuts are never quantised,
    quant/quant_model.py:57-58) run in fp32 on cuDNN's SIMT kernel in the reference.  Inside the iteration loop they go through
    the same fp32-accurate tcgen05 kernel as the reconstructed layers (forward only: neither their weights nor the unit's input
    receive a gradient) -- 0.95 ms of a 9 ms iteration on the 64x64 blocks otherwise.  Shapes the kernel does not take keep
    torch's op."""
    from .quant_layer import TC_RECONSTRUCTION
    from .tc_autograd import tc_conv
    patched = []
    if TC_RECONSTRUCTION:
        for m in unit.modules():
            if type(m) is torch.nn.Conv2d and m.weight.is_cuda and m.padding_mode == "zeros":
                kw = dict(stride=m.stride, padding=m.padding, dilation=m.dilation, groups=m.groups)

                def fwd(x, m=m, kw=kw):
                    y = tc_conv(x, m.weight.detach(), m.bias.detach() if m.bias is not None else None, kw)
                    return y if y is not None else torch.nn.Conv2d.forward(m, x)
                m.forward = fwd
                patched.append(m)
    try:
        yield
    finally:
        for m in patched:
            del m.forward            # back to the class's forward


def _common(model, unit, cali_data, batch_size, iters, w, opt_mode, asym, include_act_func, b_range, warmup, use_aq,
            p, multi_gpu, keep_gpu, layers, forward, cache_model):
    if use_aq:
        raise NotImplementedError("learning activation deltas is unreachable from the reference entry points "
                                  "(cali_model never forwards use_aq to the reconstruction calls)")
    if opt_mode != RLOSS.MSE:
        raise NotImplementedError("only RLOSS.MSE is used by the entry points")
    org = None
    if not include_act_func:
        org, unit.act_func = unit.act_func, StraightThrough()
    if layers:
        ins, outs = save_inout(cache_model, unit, cali_data, asym, use_aq, batch_size, keep_gpu)
        with _plain_convs_on_own_kernels(unit):
            _adaround_loop(forward, layers, ins, outs, batch_size, iters, w, b_range, warmup, multi_gpu, p)
    if org is not None:
        unit.act_func = org


def layer_reconstruction(model, layer: QuantLayer, cali_data, batch_size: int = 128, iters: int = 20000,
                         w: float = 0.001, opt_mode: RLOSS = RLOSS.MSE, asym: bool = False,
                         include_act_func: bool = True, b_range: tuple = (20, 2), warmup: float = 0.0,
                         use_aq: bool = False, lr: float = 4e-5, p: float = 2.0, multi_gpu: bool = False,
                         keep_gpu=True) -> None:
    model.set_quant_state(use_wq=False, use_aq=False)
    layer.set_quant_state(use_wq=True, use_aq=use_aq)
    _common(model, layer, cali_data, batch_size, iters, w, opt_mode, asym, include_act_func, b_range, warmup, use_aq,
            p, multi_gpu, keep_gpu, [layer], layer, model)


def block_reconstruction(model, block: BaseQuantBlock, cali_data, batch_size: int = 32, iters: int = 20000,
                         w: float = 0.01, opt_mode: RLOSS = RLOSS.MSE, asym: bool = False,
                         include_act_func: bool = True, b_range: tuple = (20, 2), warmup: float = 0.0,
                         use_aq: bool = False, lr: float = 4e-5, p: float = 2.0, multi_gpu: bool = True,
                         keep_gpu=True) -> None:
    model.set_quant_state(use_wq=False, use_aq=False)
    block.set_quant_state(use_wq=True, use_aq=use_aq)
    # quant_emb layers belong to the TIB; QK / SMV matmul blocks have no layers and return early
    layers = [m for m in block.modules() if isinstance(m, QuantLayer) and m.quant_emb is False]
    _common(model, block, cali_data, batch_size, iters, w, opt_mode, asym, include_act_func, b_range, warmup, use_aq,
            p, multi_gpu, keep_gpu, layers, block, model)


def tib_reconstruction(block: BaseQuantBlock, cali_data, batch_size: int = 32, iters: int = 20000, w: float = 0.01,
                       opt_mode: RLOSS = RLOSS.MSE, asym: bool = False, include_act_func: bool = True,
                       b_range: tuple = (20, 2), warmup: float = 0.0, use_aq: bool = False, lr: float = 4e-5,
                       p: float = 2.0, multi_gpu: bool = True, keep_gpu=True) -> None:
    """TIAR: all time-embedding projections reconstructed jointly against the tuple of their FP outputs."""
    block.set_quant_state(use_wq=True, use_aq=use_aq)
    every = list(dict.fromkeys(unit_layers_all(block)))
    layers = [m for m in every if not m.ignore_recon]
    for m in every:
        # the reference swaps EVERY TIB layer to AdaRound (:233-253), including the fp first layer whose alpha
        # never receives a gradient; keep the same checkpoint schema without optimising it
        if m.ignore_recon and not isinstance(m.wqtizer, AdaRoundQuantizer) and m.wqtizer.delta is not None:
            m.wqtizer = AdaRoundQuantizer(uaqtizer=m.wqtizer, rmode=RMODE.LEARNED_HARD_SIGMOID, w=m.original_w.data)
    _common(block, block, cali_data, batch_size, iters, w, opt_mode, asym, include_act_func, b_range, warmup, use_aq,
            p, multi_gpu, keep_gpu, layers, block, block)


def unit_layers_all(tib) -> List[QuantLayer]:
    """Every QuantLayer of a TIB, including the ignore_recon first layer (the reference swaps all of
    them to AdaRound, :233-253)."""
    out = [m for m in tib.modules() if isinstance(m, QuantLayer)]
    for seq in getattr(tib, "emb_layers", []):
        out += [m for m in seq.modules() if isinstance(m, QuantLayer)]
    out += list(getattr(tib, "temb_projs", []))
    return out
