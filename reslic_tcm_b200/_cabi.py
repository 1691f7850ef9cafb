"""ctypes binding of the C ABI in ``include/reslic_b200.h``.

There is deliberately NO fallback: if the shared library is missing or a kernel call
fails, the call raises.  Tensors cross the boundary as raw device pointers
(``tensor.data_ptr()``) plus sizes/strides; the current torch CUDA stream is passed
explicitly, so the kernels are ordered with the surrounding torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

from . import _build

ABI_VERSION = 17
PEER_HANDLE_BYTES = 64
RATE_DEFERRED = 2
RATE_COLLECT = 3
EB_LUT_STRIDE = 130
Q_DEQUANTIZE, Q_NOISE, Q_IDENTITY = 0, 1, 2
MATH_FAST, MATH_MIRROR = 0, 1

c_f32p = C.c_void_p
c_i32p = C.c_void_p


def new(cls):
    """A zeroed descriptor with ``struct_size`` filled in — the only way descriptors are made here."""
    d = cls()
    d.struct_size = C.sizeof(cls)
    return d


class RateExchangeDesc(C.Structure):
    """struct reslic_rate_exchange."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("world", C.c_int32), ("rank", C.c_int32), ("ring", C.c_int32), ("reserved", C.c_int32),
        ("step", C.c_int64),
        ("peer_base", C.c_void_p), ("cursor", C.c_void_p), ("extra", C.c_void_p),
        ("pixels", C.c_double), ("images", C.c_double),
    ]


class GcDesc(C.Structure):
    """struct reslic_gc_desc (field order must match the header)."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("y", C.c_void_p), ("y_bs", C.c_int64),
        ("mu", C.c_void_p), ("mu_bs", C.c_int64),
        ("sigma", C.c_void_p), ("sigma_bs", C.c_int64),
        ("noise", C.c_void_p), ("noise_bs", C.c_int64),
        ("B", C.c_int64), ("n", C.c_int64),
        ("mode", C.c_int32), ("scale_bound", C.c_float), ("likelihood_bound", C.c_float),
        ("scale_table", C.c_void_p), ("table_len", C.c_int32),
        ("yhat", C.c_void_p), ("yhat_bs", C.c_int64),
        ("ste", C.c_void_p), ("ste_bs", C.c_int64),
        ("lik", C.c_void_p), ("lik_bs", C.c_int64),
        ("sym", C.c_void_p), ("sym_bs", C.c_int64),
        ("idx", C.c_void_p), ("idx_bs", C.c_int64),
        ("bits", C.c_void_p), ("bits_accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
        ("next_y", C.c_void_p), ("next_y_bs", C.c_int64),
        ("exchange", C.POINTER(RateExchangeDesc)),
    ]


class GcBwdDesc(C.Structure):
    """struct reslic_gc_bwd_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("y", C.c_void_p), ("y_bs", C.c_int64),
        ("mu", C.c_void_p), ("mu_bs", C.c_int64),
        ("sigma", C.c_void_p), ("sigma_bs", C.c_int64),
        ("noise", C.c_void_p), ("noise_bs", C.c_int64),
        ("B", C.c_int64), ("n", C.c_int64),
        ("mode", C.c_int32), ("scale_bound", C.c_float), ("likelihood_bound", C.c_float),
        ("g_yhat", C.c_void_p), ("g_yhat_bs", C.c_int64),
        ("g_ste", C.c_void_p), ("g_ste_bs", C.c_int64),
        ("g_lik", C.c_void_p), ("g_lik_bs", C.c_int64),
        ("g_y", C.c_void_p), ("g_y_bs", C.c_int64),
        ("g_mu", C.c_void_p), ("g_mu_bs", C.c_int64),
        ("g_sigma", C.c_void_p), ("g_sigma_bs", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
    ]


class EbDesc(C.Structure):
    """struct reslic_eb_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("z", C.c_void_p), ("z_bs", C.c_int64),
        ("noise", C.c_void_p), ("noise_bs", C.c_int64),
        ("B", C.c_int64), ("C", C.c_int64), ("hw", C.c_int64),
        ("mode", C.c_int32), ("likelihood_bound", C.c_float),
        ("matrix", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("factor", C.c_void_p * 4),
        ("medians", C.c_void_p),
        ("zhat", C.c_void_p), ("zhat_bs", C.c_int64),
        ("ste", C.c_void_p), ("ste_bs", C.c_int64),
        ("lik", C.c_void_p), ("lik_bs", C.c_int64),
        ("sym", C.c_void_p), ("sym_bs", C.c_int64),
        ("bits", C.c_void_p), ("bits_accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
        ("lut", C.c_void_p),
        ("next_y", C.c_void_p), ("next_y_bs", C.c_int64), ("next_y_n", C.c_int64),
    ]


class EbBwdDesc(C.Structure):
    """struct reslic_eb_bwd_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("z", C.c_void_p), ("z_bs", C.c_int64),
        ("noise", C.c_void_p), ("noise_bs", C.c_int64),
        ("B", C.c_int64), ("C", C.c_int64), ("hw", C.c_int64),
        ("mode", C.c_int32), ("likelihood_bound", C.c_float),
        ("matrix", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("factor", C.c_void_p * 4),
        ("medians", C.c_void_p),
        ("g_zhat", C.c_void_p), ("g_zhat_bs", C.c_int64),
        ("g_lik", C.c_void_p), ("g_lik_bs", C.c_int64),
        ("g_z", C.c_void_p), ("g_z_bs", C.c_int64),
        ("g_matrix", C.c_void_p * 5), ("g_bias", C.c_void_p * 5), ("g_factor", C.c_void_p * 4),
        ("g_medians", C.c_void_p),
        ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("half_lo", C.c_void_p), ("half_lo_bs", C.c_int64),
        ("half_up", C.c_void_p), ("half_up_bs", C.c_int64),
        ("cell", C.c_void_p), ("cell_bs", C.c_int64),
        ("g_dist", C.c_void_p), ("n_dist", C.c_int64),
    ]


class StanhTables(C.Structure):
    """struct reslic_stanh_tables."""

    _fields_ = [
        ("b", C.c_void_p), ("w", C.c_void_p), ("cum_w", C.c_void_p), ("average_points", C.c_void_p),
        ("distance_points", C.c_void_p), ("K", C.c_int32), ("symmetric", C.c_int32), ("beta", C.c_float),
    ]


class StanhGcDesc(C.Structure):
    """struct reslic_stanh_gc_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("y", C.c_void_p), ("y_bs", C.c_int64),
        ("mu", C.c_void_p), ("mu_bs", C.c_int64),
        ("sigma", C.c_void_p), ("sigma_bs", C.c_int64),
        ("B", C.c_int64), ("n", C.c_int64),
        ("training", C.c_int32), ("removing_mean", C.c_int32),
        ("scale_bound", C.c_float), ("likelihood_bound", C.c_float),
        ("tables", StanhTables),
        ("yhat", C.c_void_p), ("yhat_bs", C.c_int64),
        ("lik", C.c_void_p), ("lik_bs", C.c_int64),
        ("sym", C.c_void_p), ("sym_bs", C.c_int64),
        ("bits", C.c_void_p), ("bits_accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("ste", C.c_void_p), ("ste_bs", C.c_int64),
    ]


class StanhGcBwdDesc(C.Structure):
    """struct reslic_stanh_gc_bwd_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("y", C.c_void_p), ("y_bs", C.c_int64),
        ("mu", C.c_void_p), ("mu_bs", C.c_int64),
        ("sigma", C.c_void_p), ("sigma_bs", C.c_int64),
        ("B", C.c_int64), ("n", C.c_int64),
        ("training", C.c_int32), ("removing_mean", C.c_int32),
        ("scale_bound", C.c_float), ("likelihood_bound", C.c_float),
        ("tables", StanhTables),
        ("g_yhat", C.c_void_p), ("g_yhat_bs", C.c_int64),
        ("g_lik", C.c_void_p), ("g_lik_bs", C.c_int64),
        ("g_y", C.c_void_p), ("g_y_bs", C.c_int64),
        ("g_mu", C.c_void_p), ("g_mu_bs", C.c_int64),
        ("g_sigma", C.c_void_p), ("g_sigma_bs", C.c_int64),
        ("g_params", C.c_void_p), ("g_params_len", C.c_int64),
    ]


class EbStanhDesc(C.Structure):
    """struct reslic_eb_stanh_desc."""

    _fields_ = [
        ("struct_size", C.c_uint64),
        ("z", C.c_void_p), ("z_bs", C.c_int64),
        ("B", C.c_int64), ("C", C.c_int64), ("hw", C.c_int64),
        ("training", C.c_int32), ("likelihood_bound", C.c_float),
        ("matrix", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("factor", C.c_void_p * 4),
        ("tables", StanhTables),
        ("zhat", C.c_void_p), ("zhat_bs", C.c_int64),
        ("lik", C.c_void_p), ("lik_bs", C.c_int64),
        ("sym", C.c_void_p), ("sym_bs", C.c_int64),
        ("bits", C.c_void_p), ("bits_accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("half_lo", C.c_void_p), ("half_lo_bs", C.c_int64),
        ("half_up", C.c_void_p), ("half_up_bs", C.c_int64),
        ("cell", C.c_void_p), ("cell_bs", C.c_int64),
    ]


# every symbol include/reslic_b200.h declares: name -> (restype, argtypes)
EXPORTS = {
    "reslic_abi_version": (C.c_int, []),
    "reslic_last_error": (C.c_char_p, []),
    "reslic_device_sm_count": (C.c_int, []),
    "reslic_set_math_mode": (C.c_int, [C.c_int]),
    "reslic_get_math_mode": (C.c_int, []),
    "reslic_workspace_bytes": (C.c_int64, [C.c_int64]),
    "reslic_sizeof_gc_desc": (C.c_int64, []),
    "reslic_sizeof_gc_bwd_desc": (C.c_int64, []),
    "reslic_sizeof_eb_desc": (C.c_int64, []),
    "reslic_sizeof_eb_bwd_desc": (C.c_int64, []),
    "reslic_sizeof_stanh_tables": (C.c_int64, []),
    "reslic_sizeof_stanh_gc_desc": (C.c_int64, []),
    "reslic_sizeof_stanh_gc_bwd_desc": (C.c_int64, []),
    "reslic_sizeof_eb_stanh_desc": (C.c_int64, []),
    "reslic_sizeof_rate_exchange": (C.c_int64, []),
    "reslic_rate_exchange_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "reslic_rate_exchange_read_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "reslic_rate_exchange_publish_f64": (C.c_int, [C.POINTER(RateExchangeDesc), C.c_void_p, C.c_int64, C.c_void_p]),
    "reslic_peer_buffer_create": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "reslic_peer_buffer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "reslic_peer_buffer_close": (C.c_int, [C.c_void_p]),
    "reslic_peer_buffer_destroy": (C.c_int, [C.c_void_p]),
    "reslic_rate_finalize_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "reslic_gc_fwd_f32": (C.c_int, [C.POINTER(GcDesc), C.c_void_p]),
    "reslic_gc_bwd_f32": (C.c_int, [C.POINTER(GcBwdDesc), C.c_void_p]),
    "reslic_build_indexes_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p]),
    "reslic_dequantize_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "reslic_eb_bwd_f32": (C.c_int, [C.POINTER(EbBwdDesc), C.c_void_p]),
    "reslic_eb_bwd_workspace_bytes": (C.c_int64, [C.c_int64]),
    "reslic_eb_fwd_f32": (C.c_int, [C.POINTER(EbDesc), C.c_void_p]),
    "reslic_eb_build_lut_f32": (C.c_int, [C.POINTER(EbDesc), C.c_void_p, C.c_void_p]),
    "reslic_stanh_gc_fwd_f32": (C.c_int, [C.POINTER(StanhGcDesc), C.c_void_p]),
    "reslic_stanh_gc_bwd_f32": (C.c_int, [C.POINTER(StanhGcBwdDesc), C.c_void_p]),
    "reslic_lrp_tail_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "reslic_eb_stanh_fwd_f32": (C.c_int, [C.POINTER(EbStanhDesc), C.c_void_p]),
    "reslic_stanh_gap_workspace_bytes": (C.c_int64, []),
    "reslic_stanh_act_f32": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(StanhTables), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "reslic_rans_encoder_create": (C.c_void_p, []),
    "reslic_rans_check_reciprocals": (C.c_int64, [C.c_int64, C.c_uint64]),
    "reslic_rans_encoder_destroy": (None, [C.c_void_p]),
    "reslic_rans_encoder_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p]),
    "reslic_rate_from_likelihood_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32,
                                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "reslic_rans_slots_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "reslic_rans_encoder_push_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]),
    "reslic_rans_encoder_flush": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "reslic_rans_decoder_create": (C.c_void_p, [C.c_void_p, C.c_int64]),
    "reslic_rans_decoder_destroy": (None, [C.c_void_p]),
    "reslic_rans_decoder_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "reslic_pmf_to_quantized_cdf": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
}

# struct of this binding -> the library function that reports its sizeof (checked at load: a stale binding fails
# loudly here instead of making the library read past a short struct)
STRUCT_SIZES = {
    GcDesc: "reslic_sizeof_gc_desc", GcBwdDesc: "reslic_sizeof_gc_bwd_desc", EbDesc: "reslic_sizeof_eb_desc",
    EbBwdDesc: "reslic_sizeof_eb_bwd_desc", StanhTables: "reslic_sizeof_stanh_tables",
    StanhGcDesc: "reslic_sizeof_stanh_gc_desc", StanhGcBwdDesc: "reslic_sizeof_stanh_gc_bwd_desc",
    EbStanhDesc: "reslic_sizeof_eb_stanh_desc", RateExchangeDesc: "reslic_sizeof_rate_exchange",
}

_lib = None
_lock = threading.Lock()


class ReslicError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("RESLIC_B200_LIB", _build.LIB_PATH)


def load():
    """Load (once) and type the shared library.  Raises if it is absent — there is no
    CPU or PyTorch fallback for this path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            raise ReslicError(
                f"CUDA extension not built: {path} is missing. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (needs nvcc). There is no fallback path."
            )
        lib = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.reslic_abi_version() != ABI_VERSION:
            raise ReslicError(f"ABI mismatch: library {lib.reslic_abi_version()} != binding {ABI_VERSION}")
        for cls, fn in STRUCT_SIZES.items():
            if C.sizeof(cls) != getattr(lib, fn)():
                raise ReslicError(f"struct layout mismatch: {cls.__name__} is {C.sizeof(cls)} bytes here, {getattr(lib, fn)()} in the library")
        mode = os.environ.get("RESLIC_MATH_MODE")
        if mode:
            check_code = lib.reslic_set_math_mode({"fast": MATH_FAST, "mirror": MATH_MIRROR}[mode.lower()])
            if check_code != 0:
                raise ReslicError("bad RESLIC_MATH_MODE")
        _lib = lib
    return _lib


def set_math_mode(mode: int) -> None:
    check(load().reslic_set_math_mode(int(mode)), "reslic_set_math_mode")


def check(code: int, what: str):
    if code != 0:
        msg = load().reslic_last_error()
        raise ReslicError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------ workspace cache
_ws = {}


def workspace(device: torch.device, B: int) -> torch.Tensor:
    """Zero-initialised scratch for the per-image rate reduction, one per (device, stream).
    Kernels leave it zeroed, so it is allocated/zeroed once and reused."""
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    need = int(load().reslic_workspace_bytes(B))
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
