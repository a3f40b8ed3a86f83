"""ctypes binding of libnb200.so -- exactly the symbols include/nb200.h declares.

There is deliberately NO fallback here: if the shared library is missing or a call fails the
error is raised (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
# NB200_LIB selects an experiment build of the same sources (csrc/Makefile `variant`); never a fallback
LIB_PATH = os.environ.get("NB200_LIB") or os.path.join(PKG_DIR, "lib", "libnb200.so")
HEADER = os.path.join(ROOT, "include", "nb200.h")

NB200_FP64, NB200_FP32 = 64, 32
UNIQUE_ID_BYTES = 128
IPC_BYTES = 256
COMPARE_STATS = 22

_c = ctypes
_ctx_p = _c.c_void_p
_dp = _c.POINTER(_c.c_double)

#: name -> (restype, argtypes); must list every function in include/nb200.h
SIGNATURES = {
    "nb200_create": (_c.c_int, [_c.POINTER(_ctx_p), _c.c_int, _c.c_size_t, _c.c_int, _c.c_int]),
    "nb200_create_rank": (_c.c_int, [_c.POINTER(_ctx_p), _c.c_int, _c.c_size_t, _c.c_int, _c.c_int,
                                     _c.c_int, _c.c_int, _c.c_void_p]),
    "nb200_get_unique_id": (_c.c_int, [_c.c_void_p]),
    "nb200_ipc_export": (_c.c_int, [_ctx_p, _c.c_void_p]),
    "nb200_ipc_attach": (_c.c_int, [_ctx_p, _c.c_void_p, _c.c_int]),
    "nb200_destroy": (None, [_ctx_p]),
    "nb200_upload_aos": (_c.c_int, [_ctx_p, _c.c_void_p, _c.c_size_t]),
    "nb200_download_aos": (_c.c_int, [_ctx_p, _c.c_void_p, _c.c_size_t]),
    "nb200_generate": (_c.c_int, [_ctx_p, _c.c_int, _c.c_ulonglong, _c.c_double]),
    "nb200_shard_range": (_c.c_int, [_ctx_p, _c.POINTER(_c.c_size_t), _c.POINTER(_c.c_size_t)]),
    "nb200_forces": (_c.c_int, [_ctx_p, _c.c_double, _c.c_double, _dp]),
    "nb200_step": (_c.c_int, [_ctx_p, _c.c_double, _c.c_double, _c.c_double, _c.c_int]),
    "nb200_energy": (_c.c_int, [_ctx_p, _c.c_double, _c.c_double, _dp, _dp]),
    "nb200_accuracy_pct": (_c.c_int, [_ctx_p, _dp, _dp, _dp]),
    "nb200_compare_forces": (_c.c_int, [_ctx_p, _ctx_p, _dp]),
    "nb200_p2p_leaves": (_c.c_int, [_c.c_int, _c.c_int, _c.c_size_t, _c.c_void_p, _c.c_size_t, _c.c_size_t,
                                    _c.POINTER(_c.c_longlong), _c.POINTER(_c.c_longlong), _c.POINTER(_c.c_longlong),
                                    _c.POINTER(_c.c_longlong), _c.c_double, _c.c_double, _c.c_double, _c.c_int, _c.c_int,
                                    _dp, _dp]),
    "nb200_validation_forces": (_c.c_int, [_ctx_p, _dp, _c.POINTER(_c.c_longlong), _c.c_int]),
    "nb200_measure_fp32_peak": (_c.c_int, [_c.c_int, _dp]),
    "nb200_debug_sym_exchange": (_c.c_int, [_c.c_int, _c.c_int, _c.POINTER(_c.c_int), _c.c_int]),
    "nb200_debug_sym_rows": (_c.c_int, [_c.c_size_t, _c.c_int, _c.c_int, _c.POINTER(_c.c_int), _c.c_int]),
    "nb200_last_elapsed_ms": (_c.c_int, [_ctx_p, _dp]),
    "nb200_launch_count": (_c.c_longlong, [_ctx_p]),
    "nb200_set_option": (_c.c_int, [_ctx_p, _c.c_char_p, _c.c_long]),
    "nb200_plan": (_c.c_char_p, [_ctx_p]),
    "nb200_last_error": (_c.c_char_p, [_ctx_p]),
    "nb200_version": (_c.c_char_p, []),
}


def header_symbols() -> list[str]:
    """Function names declared in include/nb200.h (used by the ABI test)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", text)))


def build(verbose: bool = False) -> str:
    """Compile libnb200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc")], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libnb200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
