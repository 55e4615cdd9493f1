"""ctypes binding of libwepp_b200.so (include/wepp_b200.h).  Fails loudly if the CUDA
library has not been built — there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwepp_b200.so")

i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
VP = C.c_void_p


class WeppStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_nodes", "n_events", "n_euler_entries", "n_reads", "n_buckets", "n_lists", "n_tiles",
        "list_entries_total", "scanned_entries", "scanned_read_entries", "algorithmic_bytes",
        "kernel_launches")] + [("ms_place_total", C.c_float), ("ms_scan_kernel", C.c_float),
                               ("ms_node_kernels", C.c_float), ("reads_per_tile", C.c_int32),
                               ("stripe_width", C.c_int32), ("place_path", C.c_int32), ("n_states", C.c_int32),
                               ("n_window_groups", C.c_int32), ("ms_exchange", C.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/wepp_b200.h declares
SIGNATURES = {
    "wepp_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "wepp_destroy": (None, [VP]),
    "wepp_last_error": (C.c_char_p, []),
    "wepp_abi_version": (C.c_int, []),
    "wepp_set_stream": (C.c_int, [VP, VP]),
    "wepp_sync": (C.c_int, [VP]),
    "wepp_set_options": (C.c_int, [VP, C.c_int32, C.c_int32]),
    "wepp_set_arena": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, VP, C.c_int32]),
    "wepp_set_reads": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP]),
    "wepp_set_mapped": (C.c_int, [VP, VP]),
    "wepp_set_allreduce": (C.c_int, [VP, VP, VP]),
    "wepp_place": (C.c_int, [VP, C.c_int32, C.c_int64]),
    "wepp_place_subset": (C.c_int, [VP, C.c_int64, VP, C.c_int32, C.c_int64]),
    "wepp_get_read_results": (C.c_int, [VP, VP, VP]),
    "wepp_get_node_results": (C.c_int, [VP, VP, VP]),
    "wepp_get_node_summary": (C.c_int, [VP, VP, VP]),
    "wepp_get_epp": (C.c_int, [VP, VP, VP, C.c_int64, C.POINTER(C.c_int64)]),
    "wepp_cartesian_map": (C.c_int, [VP, C.c_int64] + [VP] * 11),
    "wepp_filter_peaks": (C.c_int, [VP, VP, VP, VP, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "wepp_rescore": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, VP, C.c_int64]),
    "wepp_rescore_reads": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, C.c_int32, VP, VP, VP, VP, VP, C.c_int64]),
    "wepp_peer_export": (C.c_int, [VP, VP]),
    "wepp_peer_open": (C.c_int, [VP, C.c_int32, C.c_int32, VP]),
    "wepp_peer_merge": (C.c_int, [VP]),
    "wepp_peer_close": (C.c_int, [VP]),
    "wepp_group_create": (C.c_int, [C.c_int32, VP, C.POINTER(VP)]),
    "wepp_group_destroy": (None, [VP]),
    "wepp_group_size": (C.c_int32, [VP]),
    "wepp_group_handle": (VP, [VP, C.c_int32]),
    "wepp_group_take": (VP, [VP, C.c_int32]),
    "wepp_group_run": (C.c_int, [VP, VP, VP]),
    "wepp_group_set_arena": (C.c_int, [VP, C.c_int32, VP, VP, VP, VP, VP, C.c_int32]),
    "wepp_group_set_reads": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP]),
    "wepp_group_place": (C.c_int, [VP]),
    "wepp_group_get_read_results": (C.c_int, [VP, VP, VP]),
    "wepp_group_filter_peaks": (C.c_int, [VP, VP, VP, VP, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "wepp_cli_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "wepp_device_buffer": (C.c_int, [VP, C.c_int32, C.POINTER(VP), C.POINTER(C.c_int64)]),
    "wepp_get_stats": (C.c_int, [VP, C.POINTER(WeppStats)]),
    "wepp_arena_build": (C.c_int, [C.c_int32, VP, VP, VP, VP, VP, C.c_int32, C.c_int32, VP, C.c_int64, VP, VP, VP, VP, VP, C.POINTER(VP)]),
    "wepp_arena_free": (None, [VP]),
    "wepp_arena_dims": (C.c_int, [VP, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "wepp_arena_get": (C.c_int, [VP] * 10),
    "wepp_arena_get_reads": (C.c_int, [VP] * 4),
    "wepp_set_arena_from": (C.c_int, [VP, VP]),
    "wepp_mat_load": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(VP)]),
    "wepp_mat_parse": (C.c_int, [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(VP)]),
    "wepp_mat_free": (None, [VP]),
    "wepp_mat_dims": (C.c_int, [VP, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "wepp_mat_get": (C.c_int, [VP] * 9),
    "wepp_mat_get_clades": (C.c_int, [VP] * 3),
    "wepp_mat_serialize": (C.c_int64, [C.c_int32] + [VP] * 9 + [C.c_int64]),
    "wepp_reads_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, C.c_int32, C.POINTER(VP)]),
    "wepp_reads_parse": (C.c_int, [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int32, C.POINTER(VP)]),
    "wepp_reads_free": (None, [VP]),
    "wepp_reads_dims": (C.c_int, [VP] + [C.POINTER(C.c_int64)] * 7),
    "wepp_reads_get": (C.c_int, [VP] * 9),
    "wepp_reads_get_reverse": (C.c_int, [VP] * 6),
    "wepp_host_euler_stripes": (C.c_int64, [C.c_int32, VP, VP, VP, VP, VP, C.c_int32, C.c_int32, VP, C.c_int64, VP, C.c_int32]),
    "wepp_host_read_plan": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64] + [VP] * 10 + [C.POINTER(C.c_int32)]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m wepp_b200.build` "
                "(nvcc, sm_100a).  wepp_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    """Pointer to a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(VP)


class WeppError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"wepp_b200 error {code}: {msg}")
        self.code = code


def check(rc: int) -> int:
    if rc < 0:
        raise WeppError(int(rc), load().wepp_last_error().decode("utf-8", "replace"))
    return rc
