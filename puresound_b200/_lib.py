"""ctypes binding of ``libpuresound_b200.so`` (the C ABI in include/puresound_b200.h).

There is no fallback: if the shared library is missing or the device is not a
B200-class GPU the ops raise, they never route to PyTorch eager or to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PS_B200_LIB: path of another build of the same library (profiling runs load the -DPS_EXPERIMENTS build this way)
LIB_PATH = os.environ.get("PS_B200_LIB") or os.path.join(HERE, "libpuresound_b200.so")

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F32 = C.c_float

ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
PRO_NONE, PRO_AFFINE, PRO_ROWNORM, PRO_MASK = 0, 1, 2, 3
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2


class GemmDesc(C.Structure):
    _fields_ = [
        ("batch", I64), ("rows", I64), ("M", I64), ("K", I64),
        ("X", P), ("x_batch_stride", I64), ("x_row_stride", I64),
        ("W", P), ("w_row_stride", I64),
        ("Y", P), ("y_batch_stride", I64), ("y_row_stride", I64),
        ("pro_mode", I32), ("pro_act", I32),
        ("pro_a", P), ("pro_b", P), ("pro_batch_stride", I64),
        ("pro_rowstats", P), ("pro_slope", P), ("X2", P),
        ("bias", P), ("bias_batch", P),
        ("epi_act", I32), ("backend", I32), ("epi_slope", P),
        ("residual", P), ("res_batch_stride", I64), ("res_row_stride", I64),
        ("stats_partials", P), ("W_packed", P),
        ("fin_gamma", P), ("fin_beta", P), ("fin_eps", F32), ("fin_scale", P), ("fin_shift", P), ("fin_counter", P),
        ("ln_gamma", P), ("ln_beta", P), ("ln_eps", F32),
    ]


class DwconvDesc(C.Structure):
    _fields_ = [
        ("batch", I64), ("T", I64), ("C", I64), ("P", I32), ("dilation", I32), ("causal", I32),
        ("x", P), ("y", P), ("w", P), ("bias", P),
        ("pro_mode", I32), ("pro_act", I32),
        ("pro_a", P), ("pro_b", P), ("pro_batch_stride", I64),
        ("pro_rowstats", P), ("pro_slope", P),
        ("stats_partials", P), ("stats_slots", I64),
        ("fin_gamma", P), ("fin_beta", P), ("fin_eps", F32), ("fin_scale", P), ("fin_shift", P), ("fin_counter", P),
    ]


class LstmDesc(C.Structure):
    _fields_ = [
        ("n_seq", I64), ("L", I64), ("H", I64), ("D", I32),
        ("inner", I64), ("outer_stride", I64), ("inner_stride", I64), ("step_stride", I64),
        ("gx", P), ("w_hh_t", P), ("h0", P), ("c0", P),
        ("out", P), ("hn", P), ("cn", P),
        ("w_packed", P), ("gx_interleaved", I32),
    ]


class StreamDwDesc(C.Structure):
    _fields_ = [
        ("streams", I64), ("C", I64), ("P", I32), ("dilation", I32),
        ("u", P), ("y", P), ("ring", P), ("step", P), ("w", P), ("bias", P),
        ("norm_kind", I32), ("eps", F32),
        ("n1_a", P), ("n1_b", P), ("slope1", P),
        ("n2_a", P), ("n2_b", P), ("slope2", P),
    ]


class GatedDesc(C.Structure):
    _fields_ = [
        ("batch", I64), ("rows", I64), ("C", I64),
        ("a", P), ("a_batch_stride", I64), ("a_row_stride", I64),
        ("b", P), ("b_batch_stride", I64), ("b_row_stride", I64),
        ("y", P), ("y_batch_stride", I64), ("y_row_stride", I64),
        ("mid", I64), ("a_mid_stride", I64), ("b_mid_stride", I64), ("y_mid_stride", I64),
        ("a_mode", I32), ("a_act", I32), ("a_pa", P), ("a_pb", P), ("a_pro_batch_stride", I64), ("a_rowstats", P), ("a_slope", P),
        ("b_mode", I32), ("b_act", I32), ("b_pa", P), ("b_pb", P), ("b_pro_batch_stride", I64), ("b_rowstats", P), ("b_slope", P),
    ]


class StreamHopBlock(C.Structure):
    _fields_ = [
        ("w_in", P), ("w_in_ld", I64), ("ebias", P),
        ("n1_a", P), ("n1_b", P), ("slope1", P),
        ("dw_w", P), ("dw_b", P),
        ("n2_a", P), ("n2_b", P), ("slope2", P),
        ("w_pw", P), ("b_pw", P),
        ("n3_a", P), ("n3_b", P), ("slope3", P),
        ("w_out", P), ("b_out", P),
        ("ring", P), ("P", I32), ("dilation", I32),
        ("w_in_p", P), ("w_pw_p", P), ("w_out_p", P),
    ]


class StreamHopDesc(C.Structure):
    _fields_ = [
        ("streams", I64),
        ("C", I32), ("H", I32), ("win", I32), ("hop", I32), ("n_blocks", I32), ("norm_kind", I32), ("enc_relu", I32), ("mask_act", I32),
        ("constraint", I32), ("eps", F32),
        ("w_enc", P), ("w_dec_t", P), ("blocks", P), ("chunk", P),
        ("hist", P), ("frame", P), ("frame_out", P), ("acc", P), ("out", P),
        ("step", P),
        ("feats", P), ("x", P), ("u1", P), ("u2", P), ("u3", P),
        ("barrier", P),
        ("w_enc_p", P), ("w_dec_p", P),
    ]


STRUCTS = {0: GemmDesc, 1: DwconvDesc, 2: LstmDesc, 3: StreamDwDesc, 4: GatedDesc, 5: StreamHopBlock, 6: StreamHopDesc}

# name -> (restype, argtypes); must list every symbol include/puresound_b200.h declares
SIGNATURES = {
    "ps_error_string": (C.c_char_p, [C.c_int]),
    "ps_last_cuda_error": (C.c_char_p, []),
    "ps_version": (C.c_int, []),
    "ps_device_ok": (C.c_int, []),
    "ps_struct_size": (I64, [C.c_int]),
    "ps_gemm": (C.c_int, [C.POINTER(GemmDesc), P]),
    "ps_gemm_stats_slots": (I64, [I64, I64]),
    "ps_gemm_path": (I32, [C.POINTER(GemmDesc)]),
    "ps_gemm_packed_bytes": (I64, [I64, I64]),
    "ps_gemm_pack_weights": (C.c_int, [P, I64, I64, I64, P, P]),
    "ps_stats_finalize": (C.c_int, [P, I64, I64, P, P, F32, I64, P, P, P, P]),
    "ps_stats_region": (C.c_int, [P, I64, I64, I64, I64, I64, I64, I64, I64, P, P]),
    "ps_bn_fold": (C.c_int, [P, P, P, P, F32, I64, P, P, P]),
    "ps_rowstats": (C.c_int, [P, I64, I64, I64, F32, P, P]),
    "ps_dwconv": (C.c_int, [C.POINTER(DwconvDesc), P]),
    "ps_dwconv_stats_slots": (I64, [I64, I64]),
    "ps_rownorm": (C.c_int, [P, P, P, I64, I64, P, P, F32, I32, P, P]),
    "ps_ola": (C.c_int, [P, I64, I64, I64, I64, P, I32, P, P]),
    "ps_mask_apply": (C.c_int, [P, P, P, I64, I64, I32, I32, P]),
    "ps_magnitude": (C.c_int, [P, P, I64, I64, I32, I32, P]),
    "ps_band_fill": (C.c_int, [P, I64, I64, I64, P, F32, P]),
    "ps_asp_pool": (C.c_int, [P, P, I64, I64, I64, P, P]),
    "ps_l2normalize": (C.c_int, [P, P, I64, I64, P]),
    "ps_segment": (C.c_int, [P, P, I64, I64, I64, I64, I64, I32, P]),
    "ps_merge": (C.c_int, [P, P, I64, I64, I64, I64, I64, I32, P]),
    "ps_lstm": (C.c_int, [C.POINTER(LstmDesc), P]),
    "ps_lstm_packed_bytes": (I64, [I64, I32]),
    "ps_lstm_pack_weights": (C.c_int, [P, I64, I32, P, P]),
    "ps_film_combine": (C.c_int, [P, P, P, I64, I64, P]),
    "ps_gated": (C.c_int, [C.POINTER(GatedDesc), P]),
    "ps_attention": (C.c_int, [P, P, I64, I64, I64, I32, I32, P]),
    "ps_sdr": (C.c_int, [P, P, I64, I64, I64, I64, I32, I32, I32, F32, F32, P, P]),
    "ps_transpose": (C.c_int, [P, P, I64, I64, I64, P]),
    "ps_stream_dwconv_step": (C.c_int, [C.POINTER(StreamDwDesc), P]),
    "ps_stream_push": (C.c_int, [P, P, P, I64, I64, I64, P]),
    "ps_stream_ola": (C.c_int, [P, P, P, I64, I64, I64, I32, P]),
    "ps_stream_advance": (C.c_int, [P, P]),
    "ps_stream_hop": (C.c_int, [C.POINTER(StreamHopDesc), P]),
}

_lib = None


class EngineMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises EngineMissing if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineMissing(
            f"{LIB_PATH} is not built: run `python -m puresound_b200.build` (nvcc, sm_100a). "
            "puresound_b200 has no CPU or PyTorch-eager fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ps_version() != 1:
        raise EngineMissing(f"ABI version mismatch: library {lib.ps_version()}, host 1")
    for which, st in STRUCTS.items():
        if lib.ps_struct_size(which) != C.sizeof(st):
            raise EngineMissing(f"descriptor {st.__name__} size mismatch: C {lib.ps_struct_size(which)} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    """Map a negative ps_status to the exception type the reference raises for the
    same situation: NotImplementedError for unsupported configurations
    (base_nn.py:94-95), ValueError for bad arguments, RuntimeError for CUDA faults."""
    if status == 0:
        return
    lib = load()
    msg = lib.ps_error_string(status).decode()
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    detail = lib.ps_last_cuda_error().decode()
    raise EngineError(f"{what}: {msg}: {detail}")
