"""Builds (if stale) and loads tests/emu/_build/libkernels_emu.so: every hot-path kernel compiled as
host code on the fiber SIMT emulator (tests/emu/ts_emu.h).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "_build", "libkernels_emu.so")
CSRC = os.path.join(HERE, "..", "tinysplat_b200", "csrc")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(EMU, f) for f in ("kernels_emu.cpp", "ts_emu.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, "..", "include", "tinysplat_b200.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I" + EMU,
                        "-o", LIB, srcs[0]], check=True)
    lib = C.CDLL(LIB)
    p, i, f = C.c_void_p, C.c_int, C.c_float
    lib.emu_blend_fwd.argtypes = [i, i, i, i, i, p, p, p, p, p, p, p, p, i]
    lib.emu_blend_bwd.argtypes = [i, i, i, i, i, i, p, p, p, p, p, p, p, p, i, p, p, i]
    lib.emu_bin_count.argtypes = [i, i, p, p, p, p, p, i, i, i, i, p, p]
    lib.emu_bin_scan.argtypes = [i, p, p, p, i]
    lib.emu_bin_emit.argtypes = [i, p, p, p, i, i, i, p, p]
    lib.emu_bin_sort.argtypes = [i, p, p, p, i, i, p, p]
    lib.emu_set_tile_order.argtypes = [p]
    lib.emu_set_tile_order.restype = None
    lib.emu_bin_tile_order.argtypes = [i, p, p]
    lib.emu_set_key_capacity.argtypes = [i]
    lib.emu_set_key_capacity.restype = None
    lib.emu_bin_reset_cursors.argtypes = [i, p, p]
    lib.emu_render_fused.argtypes = [i, i, i, i, i, p, p, p, p, p, p, p, p, f, f, p, i, i, p, p, i,
                                     p, p, p, p, p, p, p, p, p, p, p, p, p]
    lib.emu_shard_bwd_views.argtypes = [i, i, i, i, i, i, p, p, p, p, p, p, C.c_int64, f, p, p, p, p, p, p]
    lib.emu_dp_prepare.argtypes = [i, p, p, p, p, p]
    lib.emu_peer_exchange.argtypes = [i, i, i, i, i, i, i, p, p, p, p, p, p, p, p, p, f, p, p, p, p, p, p, p]
    lib.emu_project_fwd.argtypes = [i, p, p, p, p, p, f, f, i, i, i, p, p, p, p, p, p]
    lib.emu_project_bwd.argtypes = [i, p, p, p, p, p, f, f, i, i, i, p, p, p, p, p, p, p, p, p, p, p]
    lib.emu_adam_step.argtypes = [i, p, p, p, p, p, p, p, C.c_double, C.c_double, C.c_double]
    lib.emu_ssim_fwd.argtypes = [i, i, i, i, p, p, p, p, p, f, f, p, p, p, p]
    lib.emu_ssim_bwd.argtypes = [i, i, i, i, p, p, p, p, p, p, p, p, p, p]
    lib.emu_knn_points.argtypes = [i, i, i, p, p, p, p]
    lib.emu_l1_loss.argtypes = [C.c_longlong, p, p, i, f, f, p, p, p, i]
    _lib = lib
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
