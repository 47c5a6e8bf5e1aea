"""ctypes binding of the C-ABI library (include/chimera_st_b200.h).

The library is the ONLY compute path: if it is missing or fails to load this module raises --
there is no CPU / PyTorch fallback.  torch is used for device memory and streams only.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libchimera_st_b200.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_GLU = 0, 1, 2, 3
DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class GemmParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p), ("C", C.c_void_p),
        ("ab_dtype", C.c_int), ("c_dtype", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_longlong), ("ldc", C.c_longlong), ("ldr", C.c_longlong),
        ("a_rows", C.c_longlong),
        ("act", C.c_int), ("alpha", C.c_float),
        ("nb_outer", C.c_int), ("nb_inner", C.c_int),
        ("a_bs_outer", C.c_longlong), ("a_bs_inner", C.c_longlong), ("w_bs_inner", C.c_longlong),
        ("c_bs_outer", C.c_longlong), ("c_bs_inner", C.c_longlong),
        ("r_bs_outer", C.c_longlong), ("r_bs_inner", C.c_longlong),
        ("bias_bs_inner", C.c_int),
        ("rows_per_seg", C.c_int), ("seg_rows_valid", C.c_int),
        ("out_rows_per_seg", C.c_longlong), ("out_row_off", C.c_int),
        ("seg_len", C.c_void_p), ("segs_per_outer", C.c_int),
        ("ln_in_stats", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_in_slots", C.c_int),
        ("res_stats", C.c_void_p), ("res_slots", C.c_int), ("res_gamma", C.c_void_p), ("res_beta", C.c_void_p),
        ("C2", C.c_void_p), ("c2_dtype", C.c_int), ("ldc2", C.c_longlong),
        ("out_stats", C.c_void_p), ("ln_dim", C.c_int), ("exact_act", C.c_int), ("acc_scale", C.c_float),
    ]


class DecLinearParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("residual", C.c_void_p), ("out", C.c_void_p * 3),
        ("lda", C.c_longlong), ("ldr", C.c_longlong), ("ldo", C.c_longlong * 3), ("step_stride", C.c_longlong * 3),
        ("step", C.c_void_p),
        ("a_dtype", C.c_int), ("w_dtype", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("n_seg", C.c_int), ("act", C.c_int), ("out_dtype", C.c_int * 3),
    ]


class DecBeamParams(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("tok_in", C.c_void_p), ("tok_out", C.c_void_p), ("sc_in", C.c_void_p), ("sc_out", C.c_void_p),
        ("hist_in", C.c_void_p), ("hist_out", C.c_void_p), ("ignore", C.c_void_p),
        ("fin_tokens", C.c_void_p), ("fin_pos", C.c_void_p), ("fin_score", C.c_void_p), ("fin_len", C.c_void_p),
        ("n_final", C.c_void_p), ("finished", C.c_void_p), ("counters", C.c_void_p),
        ("B", C.c_int), ("K", C.c_int), ("V", C.c_int), ("T", C.c_int), ("max_len", C.c_int), ("min_len", C.c_int),
        ("pad", C.c_int), ("eos", C.c_int), ("len_penalty", C.c_float),
    ]


_SIGS = {
    "cst_abi_version": (C.c_int, []),
    "cst_last_error": (C.c_char_p, []),
    "cst_device_info": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cst_frame_lengths": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "cst_wave_i16_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cst_conv0_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "cst_conv0_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int, C.c_void_p]),
    "cst_conv0_apply_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_void_p]),
    "cst_gemm": (C.c_int, [C.POINTER(GemmParams), C.c_void_p]),
    "cst_layernorm": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                C.c_void_p]),
    "cst_split_f16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "cst_layernorm_ab": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p,
                                   C.c_int, C.c_int, C.c_void_p]),
    "cst_posconv_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cst_broadcast_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cst_posconv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_void_p]),
    "cst_posconv_stacked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_void_p]),
    "cst_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "cst_attention_segs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_void_p, C.c_void_p]),
    "cst_text_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cst_contrastive_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "cst_label_smoothed_ce": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_longlong, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]),
    "cst_sum": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "cst_transpose": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_longlong, C.c_void_p]),
    "cst_cast": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]),
    "cst_colsum": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "cst_act_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_float,
                              C.c_void_p]),
    "cst_act_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                              C.c_int, C.c_longlong, C.c_float, C.c_void_p]),
    "cst_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cst_attention_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong] * 6 + [C.c_int] * 6 +
                          [C.c_void_p, C.c_void_p]),
    "cst_attention_bwd_tc_ws_bytes": (C.c_longlong, [C.c_int] * 4),
    "cst_attention_bwd_tc": (C.c_int, [C.c_void_p] * 7 + [C.c_longlong] * 5 + [C.c_int] * 6 + [C.c_void_p] * 3),
    "cst_attention_dropout_fwd_ws_bytes": (C.c_longlong, [C.c_int] * 4),
    "cst_attention_dropout_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_int] + [C.c_longlong] * 3 + [C.c_int] * 6
                                  + [C.c_void_p, C.c_float, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]),
    "cst_attention_bwd_tc_dropout": (C.c_int, [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong] * 5 + [C.c_int] * 6
                                     + [C.c_void_p, C.c_float, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]),
    "cst_col2im": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "cst_rows_remap": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "cst_dropout": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p,
                              C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_uint, C.c_void_p]),
    "cst_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_float] * 7 + [C.c_void_p, C.c_void_p]),
    "cst_conv0_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "cst_embed_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "cst_add_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cst_dec_embed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p]),
    "cst_dec_linear": (C.c_int, [C.POINTER(DecLinearParams), C.c_void_p]),
    "cst_dec_attention": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                    C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cst_dec_attention_grouped": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                    C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "cst_dec_attention_beam": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                         C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p]),
    "cst_dec_beam_select": (C.c_int, [C.POINTER(DecBeamParams), C.c_void_p]),
    "cst_dec_select": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


class CstError(RuntimeError):
    pass


def load():
    """Load (once) and return the shared library; raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CstError("%s not built: run `python chimera-st_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.cst_abi_version() != 1:
            raise CstError("ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise CstError("cst error %d: %s" % (rc, load().cst_last_error().decode()))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def device_info():
    name = C.create_string_buffer(256)
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(load().cst_device_info(name, 256, C.byref(sm), C.byref(ma), C.byref(mi)))
    return name.value.decode(), sm.value, ma.value, mi.value
