#!/usr/bin/env python
"""Benchmark of the w4a8 DDIM denoising hot path (BASELINE.json: LDM-4 CelebA-HQ latent UNet, w4a8,
200 DDIM steps, batch 16 per GPU; synthetic latents, seeded random-init weights).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle port of the reference's
                                                           fake-quant path on the host cores (rank 0 only)

A "step" is one denoising step of one batch: FSC parameter switch, UNet forward, DDIM update.
value   = images/s with the latents resident in HBM  (batch * gpus / (ddim_steps * step time))
e2e     = the same through the public API with HOST buffers: per step pinned-host -> device copy of the
          latent, QuantModel.forward(x, t), device -> host read of eps (what the reference's sampler
          loop does every step, ddim/functions/denoising.py:23,38)
roofline= int8 tensor-core ops of the w4a8 conv kernel / its device time, against the int8 peak
cpu_baseline = the oracle port timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tfmq-dm_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

DDIM_STEPS = 200
BATCH = 16
WORKLOAD = ("LDM-4 CelebA-HQ latent UNet 3x64x64 (model_channels 224, mult 1-2-3-4, attn ds 2/4/8, 32-ch heads), "
            "w4a8 QuantModel, 200 DDIM steps eta=0, batch 16 per GPU (BASELINE.json configs[1])")
METRIC = "w4a8 DDIM images/sec"
SEED = 1234
# algorithmic work of one UNet forward per sample (FlopCounterMode on the reference graph)
GFLOP_PER_SAMPLE = 202.4
INT8_GFLOP_PER_SAMPLE = 166.6
# algorithmic HBM bytes of the 46 w4a8 conv launches of one batch-16 step: u8 activations in (with halo), fp32 out,
# fp32 residual in (23 launches), packed int4 weights  
W4A8_ALGO_BYTES_PER_STEP = 2.21e9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------------------- model
def ldm_timesteps(S: int):
    import numpy as np
    return [float(t) for t in np.flip(np.asarray(list(range(0, 1000, 1000 // S))) + 1)]


def build_quantised(dev, batch):
    """Seeded random-init LDM-4 UNet -> QuantModel (w4 channel-wise asym, a8 per-tensor asym, first/last layer
    exemptions) -> synthetic Phase-A: activation ranges from calibration forwards at 8 timesteps, spread over
    the 200 sampling steps as FSC tables."""
    from helpers import fp_model, synth
    from tfmq_b200.quant.calibration import _collect_act, _reset_aqtizers, load_cali_model
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    fp = fp_model("ldm", SEED).to(dev)
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    x = synth.latents((2, 3, 64, 64), 21).to(dev)
    ts = ldm_timesteps(DDIM_STEPS)
    load_cali_model(qnn, (x, torch.full((2,), ts[0], device=dev)), use_aq=True, ckpt={"weight": {}})
    anchors = []
    with torch.no_grad(), qnn.calibrating():        # synthetic Phase A: torch module graph, lazily initialised quantisers
        for i in range(8):
            k = i * (DDIM_STEPS - 1) // 7
            _reset_aqtizers(qnn)
            qnn.set_quant_state(True, True)
            qnn(x, torch.full((2,), ts[k], device=dev))
            anchors.append((k, _collect_act(qnn)))
    tables = []
    for k in range(DDIM_STEPS):
        tables.append(min(anchors, key=lambda a: abs(a[0] - k))[1])
    eng = qnn.build_engine(batch=batch)
    from tfmq_b200.samplers import DDIMSampler
    smp = DDIMSampler(qnn)
    smp.make_schedule(DDIM_STEPS)
    eng.set_schedule(ts, tables, smp.coefficient_rows())
    return qnn, eng, ts


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0:
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = int(float(self.rows[0][1])) if self.rows else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_port_step_time(batch: int, steps: int, warmup: int):
    """One UNet step + DDIM update of the oracle port (fp32 fake-quant, torch CPU ops), seconds per step."""
    from helpers import LDM4_CFG, fp_model, synth
    from oracle import unet_ref as U
    torch.set_flush_denormal(True)
    sd = fp_model("ldm", SEED).state_dict()
    spec = U.build_spec(sd)
    names = sorted(n for n, s in spec.items() if s["aq"])
    x = synth.latents((batch, 3, 64, 64), 21)
    t = torch.full((batch,), 996.0)
    # activation parameters: one calibration pass of the oracle itself (lazy MINMAX init), then frozen
    act = U.CalibratingActParams()
    with torch.no_grad():
        U.ldm_unet_forward(sd, LDM4_CFG, x, t, spec, act)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            e = U.ldm_unet_forward(sd, LDM4_CFG, x, t, spec, act)
            x0 = (x - e * 0.9) / 0.4
            x = 0.5 * x0 + 0.8 * e
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            x = synth.latents((batch, 3, 64, 64), 22 + i)
    return sum(times) / len(times), len(names)


def _reference_shims():
    """Environment gaps only: stub pytorch_lightning / omegaconf (type-only imports), a stub for the
    ddpm module quant/calibration.py imports for a type annotation, and the hard-coded .cuda() calls mapped to identity."""
    import types
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    pl.seed_everything = lambda s: torch.manual_seed(s)
    plu = types.ModuleType("pytorch_lightning.utilities")
    plud = types.ModuleType("pytorch_lightning.utilities.distributed")
    plud.rank_zero_only = lambda f: f
    oc = types.ModuleType("omegaconf")
    ocl = types.ModuleType("omegaconf.listconfig")
    ocl.ListConfig = type("ListConfig", (list,), {})
    ddpm_stub = types.ModuleType("ldm.models.diffusion.ddpm")
    ddpm_stub.LatentDiffusion = torch.nn.Module
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": plu,
                        "pytorch_lightning.utilities.distributed": plud, "omegaconf": oc, "omegaconf.listconfig": ocl,
                        "ldm.models.diffusion.ddpm": ddpm_stub})
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def reference_step_time(batch: int, steps: int, warmup: int, ftz: bool):
    """One denoising step of the UNMODIFIED reference (baseline/_ref: quant.quant_model.QuantModel over
    ldm.modules.diffusionmodules.openaimodel.UNetModel, w4a8 fake-quant, fp32, torch CPU) as the reference's sampler runs it:
    FSC `load_state_dict(act_k)` (ldm/models/diffusion/ddpm.py:1402-1405), UNet forward, the p_sample_ddim update
    (ldm/models/diffusion/ddim.py:196-211).  Same seeded random-init weights and latents as the GPU arm."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    for p_ in (ref, os.path.join(ref, "stable-diffusion")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    _reference_shims()
    torch.set_flush_denormal(bool(ftz))
    from helpers import synth
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from quant.quant_layer import QMODE, Scaler, UniformAffineQuantizer
    from quant.quant_model import QuantModel
    from tfmq_b200.host.ldm_unet import celebahq_ldm4_config
    fp = UNetModel(**celebahq_ldm4_config()).eval()
    synth.fill_state_dict(fp, SEED)
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    qnn = QuantModel(fp, wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value])
    qnn.eval()
    x = synth.latents((batch, 3, 64, 64), 21)
    ts = ldm_timesteps(DDIM_STEPS)
    with torch.no_grad():
        qnn.set_quant_state(True, True)
        qnn(x, torch.full((batch,), ts[0]))          # lazy quantiser initialisation (weights: channel-wise MINMAX)
        qnn.disable_out_quantization()
        act = {}
        for name, m in qnn.model.named_modules():
            if "aqtizer" in name and isinstance(m, UniformAffineQuantizer) and m.delta is not None:
                m.zero_point = torch.nn.Parameter(torch.as_tensor(m.zero_point).float())     # as load_cali_model leaves them
                act["model." + name + ".delta"] = m.delta.detach().clone()
                act["model." + name + ".zero_point"] = m.zero_point.detach().clone()
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            qnn.load_state_dict(act, strict=False)                       # the FSC switch of DiffusionWrapper.forward
            e = qnn(x, torch.full((batch,), ts[i % DDIM_STEPS]))
            x0 = (x - 0.9 * e) / 0.4                                       # p_sample_ddim's update with fixed coefficients
            xn = 0.5 * x0 + 0.8 * e
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            x = synth.latents((batch, 3, 64, 64), 22 + i)
            del xn
    return min(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = 2
    warm = max(0, min(args.warmup, 2))
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "quant"))
    if have_ref:
        best, mean = reference_step_time(sample_batch, max(1, args.steps), warm, ftz=True)
        t_step = mean
        kind = "reference"
        # the as-is timing (the reference never sets flush-denormal; random-init weights push fake-quant into denormals)
        asis_best, asis_mean = reference_step_time(sample_batch, 2, 1, ftz=False)
        how = (f"{args.steps} steps (+{warm} warm-up) at batch {sample_batch} of the 200-step batch-16 workload: the unmodified "
               f"reference (baseline/_ref) QuantModel w4a8 fake-quant path, torch CPU fp32, flush-denormal ON (mean "
               f"{mean * 1e3:.0f} ms, best {best * 1e3:.0f} ms per step); as-is (denormals, 2 steps): {asis_mean * 1e3:.0f} ms per step")
    else:
        t_step, _ = cpu_port_step_time(sample_batch, max(1, args.steps), warm)
        kind = "port"
        how = (f"{args.steps} steps (+{warm} warm-up) at batch {sample_batch}: oracle port of the reference fake-quant path "
               "(baseline/_ref absent: run `python baseline/populate_ref.py`), flush-denormal on")
    value = sample_batch / (DDIM_STEPS * t_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 fake-quant (CPU)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ddim_steps": DDIM_STEPS, "sample_batch": sample_batch},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind, "sample": how},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def time_w4a8_kernels(eng):
    """Device time of the dominant kernel (tcgen05 w4a8 implicit-GEMM conv) inside one step: the program is
    run eagerly with a CUDA-event pair around every conv_w4a8 launch on the launching stream.  Also returns what exactly
    those launches compute: int8 ops (2 M Cout K) and algorithmic HBM bytes (u8 activations with halo in, fp32 out,
    fp32 residual in, packed int4 weights + per-channel constants), summed over the SAME launches."""
    from tfmq_b200 import ops
    ev = []
    work = {"ops": 0.0, "bytes": 0.0}
    orig = ops.conv_w4a8

    def timed(act, ksize, packed, wzp_u8, wdelta, wsum, bias, aq, out, emb=None, res=None, stats=None):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(act, ksize, packed, wzp_u8, wdelta, wsum, bias, aq, out, emb=emb, res=res, stats=stats)
        e.record()
        ev.append((s, e))
        n, h, w, cout = out.shape
        cin = act.shape[3]
        if len(ev) <= work.get("n", 1 << 30):
            work["ops"] += 2.0 * n * h * w * cout * cin * ksize * ksize
            work["bytes"] += act.numel() + out.numel() * 4 + (out.numel() * 4 if res is not None else 0) + packed.numel() + cout * 13

    ops.conv_w4a8 = timed
    try:
        saved = eng.x_in.clone()
        for rep in range(2):
            ev.clear()
            work["ops"] = work["bytes"] = 0.0
            eng._run_program(True)
            torch.cuda.synchronize()
        eng.x_in.copy_(saved)
    finally:
        ops.conv_w4a8 = orig
    return sum(s.elapsed_time(e) for s, e in ev) * 1e-3, len(ev), work["ops"], work["bytes"]


def ncu_traffic():
    """DRAM bytes per launch of the w4a8 conv kernel from the committed ncu capture (profiles/r2_ncu_full_igemm.csv: every conv
    launch of one step; dram__bytes_read.sum + dram__bytes_write.sum, mean over the 46 w4a8 launches)."""
    import csv
    name = next((n for n in ("r2_ncu_full_igemm.csv", "r1j_ncu_full_igemm.csv", "r1f_ncu_full_igemm.csv")
                 if os.path.exists(os.path.join(ROOT, "profiles", n))), None)
    if name is None:
        return None, "no ncu capture committed"
    path = os.path.join(ROOT, "profiles", name)
    rows = list(csv.reader(open(path)))[1:]
    hdr = rows[0]
    k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    vals = [(float(x[r]) + float(x[w])) * 1e6 for x in rows[2:] if "igemm_kernel<0" in x[k]]
    if not vals:
        return None, "capture holds no w4a8 launch"
    note = f"mean over {len(vals)} captured w4a8 launches (profiles/{name})"
    if name.startswith("r2_"):
        note += ("; all 46 launches of one step, the same set algorithmic_bytes_per_launch averages over; DRAM traffic is BELOW the "
                 "algorithmic bytes because a layer's fp32 output (<= 59 MB) and its u8 input stay in the 126 MB L2 between "
                 "producer and consumer -- the kernel's memory-side load is its L2 traffic, 293 MB per launch (lts__t_bytes: every "
                 "A tile is fetched once per tap and per N tile), see DESIGN 7.1")
    return sum(vals) / len(vals), note


def int8_cublas_tops(dev):
    """cuBLASLt int8 GEMM (torch._int_mm) 8192^3, best of 10 -- a library reference point for the int8 peak."""
    try:
        a = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        best = 1e9
        for _ in range(12):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch._int_mm(a, b)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def int8_peak_protocol():
    """int8 dense peak measured with the protocol of MEASURED_PEAKS.json (8192^3 s8, best of 10 / 4 s back to back) by
    tools/int8_peak.py on this pool's B200 (profiles/r2a_int8_peak.json)."""
    path = os.path.join(ROOT, "profiles", "r2a_int8_peak.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d["int8_tops"], d["int8_tops_sustained"]
    return None, None


def int8_pipeline_tops(dev):
    """The conv kernel's own TMA -> UMMA pipeline on a dense u8 x s8 GEMM (8192^3, no int4 unpack, no halo):
    what the tcgen05 structure reaches when nothing but the main loop is in the way."""
    try:
        from tfmq_b200 import ops
        a = torch.randint(0, 255, (8192, 8192), dtype=torch.uint8, device=dev)
        b = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        out = torch.empty((8192, 8192), dtype=torch.int32, device=dev)
        best = 1e9
        for _ in range(6):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.gemm_i8_peak(a, b, out)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception as exc:      # noqa: BLE001
        return f"failed: {exc}"


def time_first_stage(dev, batch):
    """The step after the path: vq-f4 first-stage decode of one batch of latents (latent -> 256x256 image),
    so that images/s can also be read as "including the decode".  Reported next to the headline, not folded into it."""
    from helpers import synth
    from tfmq_b200.first_stage import FirstStageModel, vq_f4_config
    cfg = vq_f4_config()
    fs = FirstStageModel(**cfg).eval()
    synth.fill_state_dict(fs, SEED)
    fs.quantize.embedding.weight.data.copy_(synth.latents((cfg["n_embed"], cfg["embed_dim"]), 91))
    fs = fs.to(dev)
    z = synth.latents((batch, 3, 64, 64), 77).to(dev)
    for _ in range(3):
        fs.decode_first_stage(z)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fs.decode_first_stage(z)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 5, fs.engine(batch, 64, 64, dev).launches_per_decode


def extra_config_step(dev, name: str, nb: int, steps: int):
    """Device time per denoising step of another BASELINE config on this GPU (graph replays, CUDA events):
    "cifar": configs[0], DDIM CIFAR-10 batch 1; "sd_v14" / "cin256": the per-GPU share of configs[2] / [4] (guidance halves
    included in nb), synthetic conditioning, lazily initialised MINMAX quantisers."""
    from helpers import fp_model, synth
    from tfmq_b200.quant.quant_layer import QMODE, Scaler
    from tfmq_b200.quant.quant_model import QuantModel
    wq = dict(bits=4, channel_wise=True, scaler=Scaler.MINMAX)
    aq = dict(bits=8, channel_wise=False, scaler=Scaler.MINMAX, leaf_param=True)
    if name == "cifar":
        fp, res, cin, ctx, scale = fp_model("cifar", SEED), 32, 3, None, None
    else:
        from tfmq_b200.host import ldm_unet as H
        cfg = dict(sd_v14=H.sd_v14_config, cin256=H.cin256_config)[name]()
        fp = H.UNetModel(**cfg).eval()
        synth.fill_state_dict(fp, 7)
        tk = 77 if name == "sd_v14" else 1
        res, cin, scale = 64, cfg["in_channels"], (7.5 if name == "sd_v14" else 3.0)
        ctx = synth.latents((nb, tk, cfg["context_dim"]), 22).to(dev)
    qnn = QuantModel(fp.to(dev), wq, aq, cali=False, softmax_a_bit=8, aq_mode=[QMODE.NORMAL.value, QMODE.QDIFF.value]).eval()
    x = synth.latents((nb, cin, res, res), 21).to(dev)
    t = torch.full((nb,), 601.0, device=dev)
    qnn.set_quant_state(True, True)
    with torch.no_grad(), qnn.calibrating():
        n0 = min(nb, 2)
        qnn(*((x[:n0], t[:n0]) + ((ctx[:n0],) if ctx is not None else ())))
    qnn.disable_out_quantization()
    eng = qnn.build_engine(batch=nb, context_shape=tuple(ctx.shape[1:]) if ctx is not None else None)
    eng.set_schedule([601.0] * 8, None, [[0.9, 0.4, 0.92, 0.39, 0.0]] * 8)
    if scale is not None:
        eng.set_guidance(scale)
        eng.ctx_in.copy_(ctx)
    eng.x_in.copy_(x)
    for k in range(4):
        eng.step(k)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(steps):
        eng.step(k % 8)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    launches = eng.launches_per_step
    del eng, qnn, fp
    torch.cuda.empty_cache()
    return ms, launches


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: there is no CPU fallback for the product path "
                           "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_log = None
    if world > 1:
        # one JSON line on stdout: NCCL's INFO output goes to a per-rank file (its communicator lines prove the rank count)
        nccl_log = os.path.join(tempfile.gettempdir(), f"tfmq_bench_nccl_{os.getpid()}_rank%r.log".replace("%r", str(rank)))
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
        os.environ["NCCL_DEBUG_FILE"] = nccl_log
        dist.init_process_group("nccl", device_id=dev)
    from tfmq_b200 import _lib
    from helpers import synth

    qnn, eng, ts = build_quantised(dev, BATCH)
    if world > 1:
        # the one collective of the sampling path: rank 0's packed int4 weights / scales / FSC table -> all ranks
        from tfmq_b200.dist_utils import broadcast_tensors, engine_constants
        broadcast_tensors(engine_constants(eng), 0)
    x_T = synth.latents((BATCH, 3, 64, 64), 100 + rank).to(dev)
    ctx = _lib.context(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: latents resident in HBM
    eng.x_in.copy_(x_T)
    for i in range(args.warmup):
        eng.step(i % DDIM_STEPS)
    launches_per_step = eng.launches_per_step
    eng.x_in.copy_(x_T)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        eng.step(i % DDIM_STEPS)
    e.record()
    barrier()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps

    # ---- e2e: host buffers through QuantModel.forward, H2D + D2H inside the timed region
    x_host = x_T.cpu().pin_memory()
    eps_host = torch.empty_like(x_host).pin_memory()
    t_dev = [torch.full((BATCH,), t, device=dev) for t in ts]
    e2e_steps = max(50, min(args.steps, 200))
    step_ms = []
    with torch.no_grad():
        for i in range(3):
            eng.select_step(i)
            qnn(x_host.to(dev, non_blocking=True), t_dev[i])
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for i in range(e2e_steps):
            t0 = time.perf_counter()
            eng.select_step(i % DDIM_STEPS)
            out = qnn(x_host.to(dev, non_blocking=True), t_dev[i % DDIM_STEPS])
            eps_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the sampler needs eps on the host to continue
            step_ms.append((time.perf_counter() - t0) * 1e3)
        e2.record()
        barrier()
    step_ms.sort()
    ms2 = torch.tensor([s2.elapsed_time(e2) / e2e_steps, step_ms[len(step_ms) // 2], step_ms[0]], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e, ms_e2e_median, ms_e2e_min = (float(v) for v in ms2.tolist())
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    shard = {}
    if (world == 8 or os.environ.get("TFMQ_BENCH_SHARDS") == "1") and not args.no_extra:     # the env switch: dry runs at other N
        # BASELINE configs[2] (SD v1.4 batch 8 over 8 GPUs: 1 prompt x 2 guidance halves per GPU, 50 steps) and configs[4]
        # (cin256 batch 64 over 8 GPUs: 8 classes x 2 per GPU, 250 steps): every rank times its share, max over ranks
        for nm, nb, nsteps in (("sd_v14", 2, 50), ("cin256", 16, 250)):
            try:
                ms_x, l_x = extra_config_step(dev, nm, nb, 20)
            except Exception as exc:      # noqa: BLE001
                ms_x, l_x = float("nan"), -1
                print(f"[rank {rank}] {nm} shard failed: {exc}", file=sys.stderr)
            tt = torch.tensor([ms_x], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            shard[nm] = {"ms_per_step_max_over_ranks": tt.item(), "launches_per_step": l_x, "batch_per_gpu": nb,
                         "images_per_s_8gpu": 8 * (nb // 2) / (nsteps * tt.item() * 1e-3), "steps_per_image": nsteps}
    if rank == 0:
        pk = peaks()
        # ops and algorithmic bytes of exactly the launches that are timed (the weight-only-quantised layers run on the fp
        # path and are not among them)
        conv_s, conv_n, int8_ops, algo_bytes = time_w4a8_kernels(eng)
        achieved = int8_ops / conv_s / 1e12
        cub = int8_cublas_tops(dev)
        # int8 dense peak: measured, protocol of MEASURED_PEAKS.json (8192^3 s8: best of 10 = burst, 4 s back to back =
        # sustained).  The kernel is event-timed launch by launch in an eager pass at un-capped clocks: the BURST figure applies.
        burst, sustained = int8_peak_protocol()
        int8_peak = burst if burst else (cub if cub else 2.0 * pk["bf16_burst"])
        own = int8_pipeline_tops(dev)
        traffic, traffic_note = ncu_traffic()
        images_s = BATCH * world / (DDIM_STEPS * ms_per_step * 1e-3)
        # e2e value from the MEDIAN step (host wall clock around copy-in, step, copy-out and the sync; max over ranks): a single
        # host hiccup (the nvidia-smi sampler thread forks a process during the timed region) moved the 50-step mean by 15 %
        # between otherwise identical runs; the mean is reported next to it
        e2e_images_s = BATCH * world / (DDIM_STEPS * ms_e2e_median * 1e-3)
        cpu_line = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            # the reference arm in a child process (its environment shims patch torch.Tensor.cuda: not in this process)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4",
                                      "--warmup", "1"], capture_output=True, text=True, timeout=900)
                cpu_line = json.loads([l_ for l_ in out.stdout.splitlines() if l_.startswith("{")][-1])["cpu_baseline"]
            except Exception as exc:      # noqa: BLE001
                cpu_line = {"value": None, "unit": "images/s", "cores": cores, "kind": "unavailable", "sample": repr(exc)[:200]}
        line = {
            "metric": METRIC, "value": images_s, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 x s4->s8 (int32 accumulate); fp layers fp16 hi/lo split x3 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "ddim_steps": DDIM_STEPS, "batch_per_gpu": BATCH,
                       "l2": "per-step activation working set (several GB) >> 126 MB L2, no explicit flush",
                       "step_gflop": GFLOP_PER_SAMPLE * BATCH, "step_tflops": GFLOP_PER_SAMPLE * BATCH / ms_per_step,
                       "images_per_s_at_50_steps": images_s * DDIM_STEPS / 50},
            "e2e": {"value": e2e_images_s, "unit": "images/s", "value_from": "median step of the e2e loop",
                    "ms_per_step": ms_e2e_median, "ms_per_step_mean": ms_e2e, "ms_per_step_median": ms_e2e_median,
                    "ms_per_step_min": ms_e2e_min, "steps": e2e_steps,
                    "h2d_bytes_per_step": x_host.numel() * 4 + BATCH * 4, "d2h_bytes_per_step": eps_host.numel() * 4},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s",
                         "frac": achieved / int8_peak, "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_launch": algo_bytes / max(conv_n, 1),
                         "int8_ops_per_step": int8_ops,
                         "kernel": "igemm_kernel<MODE_W4A8, CG=2> (tcgen05 kind::i8, cta_group::2)",
                         "launches_per_step": conv_n,
                         "kernel_ms_per_step": conv_s * 1e3,
                         "peak_note": f"int8 dense BURST peak measured with the MEASURED_PEAKS.json protocol (8192^3 s8 cuBLASLt, best "
                                      f"of 10: {burst} TOP/s; 4 s sustained: {sustained}; profiles/r2a_int8_peak.json); the same GEMM "
                                      f"measured in this run: {cub}; this kernel's pipeline on a dense u8 x s8 8192^3 GEMM "
                                      f"(MODE_I8): {own}",
                         "step_frac": GFLOP_PER_SAMPLE * BATCH / ms_per_step / int8_peak},
            "clocks": sampler.summary() if sampler else None,
        }
        if world == 1:
            try:
                fs_ms, fs_launches = time_first_stage(dev, BATCH)
                line["first_stage"] = {
                    "workload": "vq-f4 decode_first_stage, 16 latents 3x64x64 -> 3x256x256 (SURVEY f3; not part of `value`)",
                    "ms_per_decode": fs_ms, "launches": fs_launches,
                    "images_per_s_with_decode": BATCH / (DDIM_STEPS * ms_per_step * 1e-3 + fs_ms * 1e-3)}
            except Exception as exc:      # noqa: BLE001
                line["first_stage"] = f"failed: {exc}"
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        if world == 1 and not args.no_extra:
            try:      # BASELINE configs[0]: "absolute img/s only" (launch-latency-bound at batch 1)
                ms_c, l_c = extra_config_step(dev, "cifar", 1, 100)
                line["cifar10_batch1"] = {"workload": "DDIM CIFAR-10 32x32 UNet, w4a8, 50 steps, batch 1 (BASELINE.json configs[0])",
                                          "ms_per_step": ms_c, "images_per_s": 1.0 / (50 * ms_c * 1e-3), "launches_per_step": l_c}
            except Exception as exc:      # noqa: BLE001
                line["cifar10_batch1"] = f"failed: {exc}"
        if shard:
            line["sharded_configs_8gpu"] = shard
        if nccl_log is not None:
            try:
                txt = open(nccl_log).read()
                import re
                m = re.findall(r"nranks (\d+)", txt)
                line["nccl"] = {"init_lines": sum(1 for l_ in txt.splitlines() if "comm 0x" in l_ or "Init COMPLETE" in l_),
                                "nranks": sorted({int(v) for v in m}), "log": nccl_log}
            except OSError as exc:
                line["nccl"] = f"log unreadable: {exc}"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the CIFAR batch-1 / 8-GPU shard timings of the other configs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 6:
            args.steps = 6          # bounded sample: each CPU step is seconds
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
