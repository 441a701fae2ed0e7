"""ctypes binding of libtinysplat_b200.so (the C ABI declared in include/tinysplat_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
product raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported
from here.)"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TINYSPLAT_B200_LIB") or os.path.join(_HERE, "libtinysplat_b200.so")

_p = C.c_void_p
_i = C.c_int
_f = C.c_float

_SIGNATURES = {
    "ts_version": ([], C.c_int),
    "ts_last_error": ([], C.c_char_p),
    "ts_rec_floats": ([], C.c_int),
    "ts_grad_floats": ([], C.c_int),
    "ts_launch_count": ([], C.c_int64),
    "ts_project_fwd": ([_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _i, _i, _i, _i, _f, _i,
                        _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p], C.c_int),
    "ts_project_bwd": ([_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _i, _i, _i,
                        _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "ts_sh_fwd": ([_i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _i, _p], C.c_int),
    "ts_sh_bwd": ([_i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _p], C.c_int),
    "ts_bin_count": ([_i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p], C.c_int),
    "ts_bin_scan": ([_i, _p, _p, _p, _i, _p], C.c_int),
    "ts_bin_emit": ([_i, _p, _p, _p, _i, _i, _i, _p, _p, _p, _i, _p], C.c_int),
    "ts_bin_reset_cursors": ([_i, _p, _p, _p], C.c_int),
    "ts_bin_sort": ([_i, _p, _p, _p, _i, _i, _p, _p, _i, _p], C.c_int),
    "ts_bin_smem_sort_cap": ([], C.c_int),
    "ts_bin_counter_stride": ([], C.c_int),
    "ts_bin_scan_work_ints": ([], C.c_int),
    "ts_bin_tile_order": ([_i, _p, _p, _p], C.c_int),
    "ts_blend_fwd": ([_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p], C.c_int),
    "ts_blend_bwd": ([_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p], C.c_int),
    "ts_ssim_fwd": ([_i, _i, _i, _i, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p], C.c_int),
    "ts_ssim_bwd": ([_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "ts_knn_points": ([_i, _i, _i, _p, _p, _p, _p, _p], C.c_int),
    "ts_adam_max_tensors": ([], C.c_int),
    "ts_adam_step": ([_i, _p, _p, _p, _p, _p, _p, _p, C.c_double, C.c_double, C.c_double, _p], C.c_int),
    "ts_blend_unpack_grads": ([_i, _i, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "ts_dp_prepare": ([_i, _p, _p, _p, _p, _p, _p], C.c_int),
    "ts_project_bwd_views": ([_i, _i, _p, _p, _f, _p, _p, _i, _i, _i, _p, C.c_int64, _p, _f, _p, _p, _p, _p, _p],
                             C.c_int),
    "ts_sh_bwd_views": ([_i, _i, _i, _i, _p, _p, _p, C.c_int64, _f, _p, _p, _p], C.c_int),
    "ts_sh_bwd_views_rgb": ([_i, _i, _i, _i, _p, _p, _p, C.c_int64, _f, _p, _p, _p], C.c_int),
    "ts_project_bwd_views_peer": ([_i, _i, _p, _p, _f, _p, _p, _i, _i, _i, _p, C.c_int64, _p, _f, _i, _i,
                                   _p, _p, _p, _p, _p], C.c_int),
    "ts_dp_push": ([_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p], C.c_int),
    "ts_dp_exchange_split": ([_i], C.c_int),
    "ts_peer_barrier": ([_i, _i, _p, _i, C.c_uint32, _p, C.c_double, _i, _p], C.c_int),
    "ts_peer_barrier_slots": ([], C.c_int),
    "ts_dp_exchange_timeline": ([_i], C.c_int),
    "ts_dp_exchange_timeline_read": ([C.c_char_p, _i], C.c_int),
    "ts_dp_exchange_peer": ([_i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f,
                             C.c_uint32, C.c_double, _p, _p, _p, _p], C.c_int),
    "ts_peer_max_ranks": ([], C.c_int),
    "ts_peer_ipc_handle_bytes": ([], C.c_int),
    "ts_peer_flag_bytes": ([], C.c_int),
    "ts_peer_alloc": ([C.c_int64, _p], C.c_int),
    "ts_peer_free": ([_p], C.c_int),
    "ts_peer_ipc_get": ([_p, _p], C.c_int),
    "ts_peer_ipc_open": ([_p, _p], C.c_int),
    "ts_peer_ipc_close": ([_p], C.c_int),
    "ts_set_blend_mode": ([_i], C.c_int),
    "ts_get_blend_mode": ([], C.c_int),
    "ts_project_sh_bwd": ([_i, _i, _i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _i, _i, _i, _p, _p, _p, _p,
                           _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "ts_l1_loss_work_floats": ([], C.c_int),
    "ts_l1_loss": ([C.c_int64, _p, _p, _i, _f, _f, _p, _p, _p, _p], C.c_int),
    "ts_debug_rowmask": ([_p, _p, _i, _i], C.c_uint32),
    "ts_debug_approx": ([_i, _p, _p, _p, _p], C.c_int),
}

# flags (include/tinysplat_b200.h enum ts_flags)
PROJ_LOG_SCALES, PROJ_RAW_QUATS, PROJ_DEPTH_CH3, PROJ_OPACITY_LOGIT = 1, 2, 4, 8
SH_DIRS_FROM_MEANS, SH_OFFSET_CLAMP = 1, 2
BIN_OPACITY_LOGIT = 1
BIN_PACK_ONLY = 2
BLEND_GRADS_ZEROED = 2

_STATUS = {0: "TS_OK", -1: "TS_ERR_INVALID", -2: "TS_ERR_ALIGN", -3: "TS_ERR_CUDA",
           -4: "TS_ERR_CAPACITY"}

_lib: Optional[C.CDLL] = None


class TinysplatError(RuntimeError):
    pass


def exported_symbols():
    return list(_SIGNATURES)


def load() -> C.CDLL:
    """dlopen the in-tree library; raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TinysplatError(
            f"{LIB_PATH} is missing: build it with `python -m tinysplat_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        detail = load().ts_last_error().decode() if status == -3 else ""
        raise TinysplatError(f"{what} failed: {_STATUS.get(status, status)} {detail}")


# ---- optional per-call CUDA-event timing (bench.py's live per-kernel breakdown) ---------------
_profile = None
_profile_only = None


def profile_start(only=None) -> None:
    """only: entry-point names to time (None = every call).  Timing a call costs two event records on
    its stream; a bench that wants an undisturbed step times only the kernel it reports."""
    global _profile, _profile_only
    _profile = []
    _profile_only = None if only is None else frozenset(only)


def profile_stop():
    """Returns {entry point: [ms, ...]} for every C-ABI call since profile_start()."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    out = {}
    for name, s, e in rec:
        out.setdefault(name, []).append(s.elapsed_time(e))
    return out


def call(name: str, *args) -> None:
    """Invoke C-ABI entry `name`; raise on a non-zero status.  Events go on the current stream,
    which is the stream every kernel is launched on."""
    fn = getattr(load(), name)
    if _profile is None or (_profile_only is not None and name not in _profile_only):
        check(fn(*args), name)
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    status = fn(*args)
    e.record()
    _profile.append((name, s, e))
    check(status, name)


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise TinysplatError(
                "tinysplat_b200 ops run only on CUDA tensors (sm_100a); got a "
                f"{t.device} tensor.  There is no CPU fallback.")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, 16-byte aligned view/copy of t."""
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def launch_count() -> int:
    return int(load().ts_launch_count())
