"""ctypes binding of include/tfmq_b200.h.

This is the only place Python touches the C ABI.  There is no fallback: if the shared
library is missing, or no sm_100 device is present when a context is requested, the
caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libtfmq_b200.so"

c_f32p = C.c_void_p  # device pointers travel as integers
i64 = C.c_int64


class GnTarget(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("cpg", C.c_int), ("ch_off", C.c_int), ("groups", C.c_int),
                ("reserved", C.c_int)]


class ActDesc(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_ld", i64),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("upsample", C.c_int),
        ("gn_stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("groups", C.c_int), ("eps", C.c_float), ("silu", C.c_int),
        ("aq", C.c_void_p), ("dst_u8", C.c_void_p),
        ("halo", C.c_int), ("dst_c", C.c_int), ("dst_c_off", C.c_int),
        ("dst_f32", C.c_void_p), ("dst_ld", i64),
        ("dst_hi", C.c_void_p), ("dst_lo", C.c_void_p), ("dst_h_ld", i64),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("geglu", C.c_int),
    ]


class ConvW4A8Desc(C.Structure):
    _fields_ = [
        ("act", C.c_void_p),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
        ("ksize", C.c_int),
        ("packed", C.c_void_p), ("wzp", C.c_void_p), ("wdelta", C.c_void_p), ("wsum", C.c_void_p),
        ("bias", C.c_void_p), ("aq", C.c_void_p),
        ("emb", C.c_void_p), ("emb_ld", i64),
        ("res", C.c_void_p), ("res_ld", i64),
        ("out", C.c_void_p), ("out_ld", i64),
        ("n_stat", C.c_int), ("stat", GnTarget * 2),
    ]


class ConvFpDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ld", i64),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
        ("ksize", C.c_int), ("stride", C.c_int), ("pad_lo", C.c_int),
        ("out_h", C.c_int), ("out_w", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("wscale", C.c_void_p), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("res_ld", i64),
        ("out", C.c_void_p), ("out_ld", i64),
        ("passes", C.c_int),
        ("emb", C.c_void_p), ("emb_ld", i64),
        ("n_stat", C.c_int), ("stat", GnTarget * 2),
    ]


class ConvH16Desc(C.Structure):
    _fields_ = [
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p), ("x_ld", i64),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
        ("ksize", C.c_int), ("stride", C.c_int), ("pad_lo", C.c_int),
        ("out_h", C.c_int), ("out_w", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("wscale", C.c_void_p), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("res_ld", i64),
        ("out", C.c_void_p), ("out_ld", i64),
        ("emb", C.c_void_p), ("emb_ld", i64),
        ("n_stat", C.c_int), ("stat", GnTarget * 2),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_h_ld", i64),
        ("ksplit", C.c_int),
    ]


class LinearDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ld", i64),
        ("m", C.c_int), ("in_f", C.c_int), ("out_f", C.c_int),
        ("silu_in", C.c_int),
        ("aq", C.c_void_p), ("w_f32", C.c_void_p), ("codes", C.c_void_p),
        ("wzp_f", C.c_void_p), ("wdelta", C.c_void_p), ("bias", C.c_void_p),
        ("out", C.c_void_p), ("out_ld", i64),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_sb", i64), ("q_sh", i64), ("q_st", i64),
        ("k", C.c_void_p), ("k_sb", i64), ("k_sh", i64), ("k_st", i64),
        ("v", C.c_void_p), ("v_sb", i64), ("v_sh", i64), ("v_st", i64),
        ("o", C.c_void_p), ("o_sb", i64), ("o_sh", i64), ("o_st", i64),
        ("b", C.c_int), ("heads", C.c_int), ("tq", C.c_int), ("tk", C.c_int), ("d", C.c_int),
        ("scale", C.c_float),
        ("o_hi", C.c_void_p), ("o_lo", C.c_void_p),
    ]


class AttnH16Desc(C.Structure):
    _fields_ = [
        ("q_hi", C.c_void_p), ("q_lo", C.c_void_p), ("q_sb", i64), ("q_sh", i64), ("q_st", i64),
        ("k_hi", C.c_void_p), ("k_lo", C.c_void_p), ("k_sb", i64), ("k_sh", i64), ("k_st", i64),
        ("v_hi", C.c_void_p), ("v_lo", C.c_void_p), ("v_sb", i64), ("v_sh", i64), ("v_st", i64),
        ("o", C.c_void_p), ("o_hi", C.c_void_p), ("o_lo", C.c_void_p), ("o_sb", i64), ("o_sh", i64), ("o_st", i64),
        ("b", C.c_int), ("heads", C.c_int), ("tq", C.c_int), ("tk", C.c_int), ("d", C.c_int),
        ("scale", C.c_float),
    ]


P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "tfmq_create": (C.c_int, [C.POINTER(P), C.c_int]),
    "tfmq_destroy": (C.c_int, [P]),
    "tfmq_last_error": (C.c_char_p, [P]),
    "tfmq_abi_version": (C.c_int, []),
    "tfmq_launch_count": (i64, [P]),
    "tfmq_pack_w4": (C.c_int, [P, P, P, P, P, C.c_int, C.c_int, P, P, P, P]),
    "tfmq_gn_stats": (C.c_int, [P, P, i64, C.c_int, C.c_int, C.c_int, C.c_int, P, P]),
    "tfmq_gn_stats_part": (C.c_int, [P, P, i64, C.c_int, C.c_int, C.c_int, C.POINTER(GnTarget), P]),
    "tfmq_fill_zero": (C.c_int, [P, P, C.c_size_t, P]),
    "tfmq_act_prepare": (C.c_int, [P, C.POINTER(ActDesc), P]),
    "tfmq_conv_w4a8": (C.c_int, [P, C.POINTER(ConvW4A8Desc), P]),
    "tfmq_conv_fp": (C.c_int, [P, C.POINTER(ConvFpDesc), P]),
    "tfmq_conv_h16": (C.c_int, [P, C.POINTER(ConvH16Desc), P]),
    "tfmq_conv_in": (C.c_int, [P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, i64, P]),
    "tfmq_conv_out": (C.c_int, [P, P, i64, P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P]),
    "tfmq_linear_small": (C.c_int, [P, C.POINTER(LinearDesc), P]),
    "tfmq_linear_grouped_plan": (C.c_int, [P, C.POINTER(LinearDesc), C.c_int, C.POINTER(C.c_int)]),
    "tfmq_linear_grouped": (C.c_int, [P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]),
    "tfmq_timestep_embedding": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, P, P]),
    "tfmq_attention": (C.c_int, [P, C.POINTER(AttnDesc), P]),
    "tfmq_attention_h16": (C.c_int, [P, C.POINTER(AttnH16Desc), P]),
    "tfmq_ddim_update": (C.c_int, [P, P, P, P, P, i64, P, P, P]),
    "tfmq_cfg_combine": (C.c_int, [P, P, P, C.c_float, i64, P, P]),
    "tfmq_plms_eps": (C.c_int, [P, P, P, P, P, C.c_int, i64, P, P]),
    "tfmq_first_stage_input": (C.c_int, [P, P, C.c_float, P, C.c_int, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, P, P]),
    "tfmq_minmax_rows": (C.c_int, [P, P, i64, i64, P, P]),
    "tfmq_mse_scale_search": (C.c_int, [P, P, i64, i64, C.c_int, P, P, P]),
    "tfmq_act_range_update": (C.c_int, [P, P, i64, i64, C.c_int, C.c_float, C.c_int, P, P, P]),
    "tfmq_adaround_soft": (C.c_int, [P, P, P, P, P, C.c_int, i64, C.c_int, P, P]),
    "tfmq_adaround_step": (C.c_int, [P, P, P, P, P, P, P, P, C.c_int, i64, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_float, P, P]),
    "tfmq_rec_loss": (C.c_int, [P, P, P, i64, C.c_int, P, P, P]),
    "tfmq_gemm_i8_peak": (C.c_int, [P, P, P, C.c_int, C.c_int, C.c_int, P, P]),
}

EXPORTS = tuple(_SIGS)

_lib = None


def load() -> C.CDLL:
    """dlopen the library and type every entry point (works without a GPU)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError => the library does not match the header
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class Context:
    """One tfmq_ctx per device; raises if the device is not an sm_100 GPU."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.handle = P()
        rc = self.lib.tfmq_create(C.byref(self.handle), int(device))
        if rc != 0:
            msg = self.lib.tfmq_last_error(self.handle).decode() if self.handle else "allocation failed"
            if self.handle:
                self.lib.tfmq_destroy(self.handle)
            self.handle = None
            raise RuntimeError(f"tfmq_create failed ({rc}): {msg}")
        self.device = device

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.lib.tfmq_last_error(self.handle).decode()}")

    @property
    def launches(self) -> int:
        return int(self.lib.tfmq_launch_count(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.tfmq_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts: dict = {}


def context(device: int = 0) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
