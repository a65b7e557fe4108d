"""ctypes binding of ``libtetris_piclim_sm100.so`` (the C ABI declared in ``include/tetris_piclim.h``).

There is deliberately no fallback: if the CUDA library is missing, importing the binding raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_longlong, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "libtetris_piclim_sm100.so")      # short in-tree path, see build.py

P = c_void_p      # every array argument is passed as a raw address
ABI_VERSION = 2   # TPL_ABI_VERSION of include/tetris_piclim.h

# name -> (restype, argtypes); the trailing `stream` argument of the device API is listed explicitly
DEVICE_API = {
    "pack": (c_int, [P, c_int64, c_int, c_int, P, P, c_int, P, P, P, P, P, P]),
    "unpack": (c_int, [P, c_int64, c_int, P, P, P, P, P, P, P, P, P, P]),
    "reset_from_pool": (c_int, [P, c_int64, c_int, P, c_int, P, P, c_int, P, P, c_uint64, c_uint64, c_int, P]),
    "step": (c_int, [P, c_int64, c_int, P, P, P, P, P, P, c_int, c_int, P]),
    "afterstates": (c_int, [P, c_int64, c_int, P, P, P, c_int, c_int, P]),
    "step_observe": (c_int, [P, c_int64, c_int, P, P, P, P, P, P, P, c_int, P, P, c_uint64, c_uint64, c_int, P, P, P, c_int, c_int, P]),
    "afterstates_distinct": (c_int, [P, c_int64, c_int, P, c_int64, P, c_uint32, P, c_int, c_int, c_int, P]),
    "step_observe_distinct": (c_int, [P, c_int64, c_int, P, P, P, P, P, P, P, c_int, P, P, c_uint64, c_uint64, c_int, P, c_int64, P, c_uint32,
                                      P, c_int, c_int, c_int, P]),
    "expand_distinct": (c_int, [P, P, c_int, P, P]),
    "value_pack": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P]),
    "value_rows": (c_int, [P, P, c_int64, P, P, P]),
    "replay_push": (c_int, [P, P, P, P, P, P, c_int, c_int64, c_int64, c_float, c_float, c_float, P, P, P, P, P, P, P]),
    "select_action": (c_int, [P, P, P, c_int, c_float, c_float, c_float, c_float, c_uint64, c_uint64, c_uint32, P, P, P, P, P]),
    "gen_pieces": (c_int, [P, c_int, c_int, c_uint64, c_uint64, P, c_uint32, P]),
    "rollout_random": (c_int, [P, c_int64, c_int, P, c_int, P, P, P, c_int, c_uint64, c_uint64, c_int, c_int, c_int, P]),
    "rollout_greedy": (c_int, [P, c_int64, c_int, P, c_int, P, P, P, c_int, P, c_uint64, c_uint64, c_int, c_int, c_int, P]),
}

HOST_API = {
    "env_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_uint64, c_uint64]),
    "env_destroy": (None, [c_void_p]),
    "env_set_limits": (c_int, [c_void_p, c_int, c_int]),
    "env_set_pool": (c_int, [c_void_p, c_int, P, P, c_int, P]),
    "env_reset": (c_int, [c_void_p, P, P, c_int, c_int]),
    "env_load": (c_int, [c_void_p, P, P, c_int, P, P, P, P, P]),
    "env_move": (c_int, [c_void_p, P, P, P, P, P]),
    "env_get_state": (c_int, [c_void_p, P, P, P, P, P, P, P, P, P]),
    "env_afterstates": (c_int, [c_void_p, P, P]),
    "env_step_observe": (c_int, [c_void_p, P, P, P, P, P, P, P]),
    "env_step_observe_distinct": (c_int, [c_void_p, P, P, P, P, P, P, c_int64, P, ctypes.POINTER(c_int64)]),
    "env_distinct_capacity": (c_int64, [c_void_p]),
    "env_chunks": (c_int, [c_void_p]),
    "env_feats_ptr": (c_void_p, [c_void_p]),
    "env_state_ptr": (c_void_p, [c_void_p, ctypes.POINTER(c_int64)]),
    "env_stream": (c_void_p, [c_void_p]),
    "host_alloc": (c_void_p, [c_int64]),
    "host_free": (None, [c_void_p]),
}

MISC_API = {
    "abi_version": (c_int, []),
    "last_error": (c_char_p, []),
    "launch_count": (c_longlong, []),
    "distinct_tables": (None, [P, P, P]),
}

ALL_SYMBOLS = ["tpl_" + k for k in list(DEVICE_API) + list(HOST_API) + list(MISC_API)]


class TplError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for table in (DEVICE_API, HOST_API, MISC_API):
            for name, (res, args) in table.items():
                fn = getattr(L, "tpl_" + name)
                fn.restype = res
                fn.argtypes = args
        if L.tpl_abi_version() != ABI_VERSION:
            raise ImportError("libtetris_piclim_sm100.so: ABI version mismatch")
        _lib = L
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().tpl_last_error()
        raise TplError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().tpl_launch_count())
