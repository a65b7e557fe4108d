"""ctypes binding of the C oracle (oracle/piclim_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpiclim_oracle.so")
_lib = None

P_MAX = 42      # piece-queue capacity of the 64-byte env record (128 bits / 3 bits per piece)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "piclim_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpiclim_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


class BatchState:
    """Flat-array batch of oracle envs (bitrows)."""

    def __init__(self, n: int, P: int = P_MAX):
        self.n, self.P = n, P
        self.rows = np.zeros((n, 20), np.uint16)
        self.pieces = np.zeros((n, P), np.uint8)
        self.npieces = np.zeros(n, np.uint8)
        self.head = np.zeros(n, np.uint8)
        self.lines = np.zeros(n, np.int32)
        self.moves = np.zeros(n, np.int32)
        self.state = np.zeros(n, np.int8)

    def load(self, rows, pieces, npieces):
        self.rows[:] = rows
        self.pieces[:] = 0
        pieces = np.asarray(pieces, np.uint8)
        self.pieces[:, :pieces.shape[1]] = pieces
        self.npieces[:] = npieces
        self.head[:] = 0
        self.lines[:] = 0
        self.moves[:] = 0
        self.state[:] = 0
        return self

    def copy(self):
        o = BatchState(self.n, self.P)
        for k in ("rows", "pieces", "npieces", "head", "lines", "moves", "state"):
            getattr(o, k)[...] = getattr(self, k)
        return o


def step_batch(st: BatchState, rot, loc, L: int, M: int):
    rot = np.ascontiguousarray(rot, np.int32)
    loc = np.ascontiguousarray(loc, np.int32)
    dlines = np.zeros(st.n, np.int8)
    flags = np.zeros(st.n, np.uint8)
    lib().orc_step_batch(ctypes.c_int(st.n), _p(st.rows, ctypes.c_uint16), _p(st.pieces, ctypes.c_uint8),
                         ctypes.c_int(st.P), _p(st.npieces, ctypes.c_uint8), _p(st.head, ctypes.c_uint8),
                         _p(st.lines, ctypes.c_int32), _p(st.moves, ctypes.c_int32), _p(st.state, ctypes.c_int8),
                         _p(rot, ctypes.c_int32), _p(loc, ctypes.c_int32), _p(dlines, ctypes.c_int8),
                         _p(flags, ctypes.c_uint8), ctypes.c_int(L), ctypes.c_int(M))
    return dlines, flags


def afterstates_batch(st: BatchState, L: int, M: int, want_boards: bool = False, nthreads: int = 1):
    feats = np.zeros((st.n, 40, 4), np.uint8)
    flags = np.zeros((st.n, 40), np.uint8)
    boards = np.zeros((st.n, 40, 20), np.uint16) if want_boards else None
    lib().orc_afterstates_batch(ctypes.c_int(st.n), _p(st.rows, ctypes.c_uint16), _p(st.pieces, ctypes.c_uint8),
                                ctypes.c_int(st.P), _p(st.npieces, ctypes.c_uint8), _p(st.head, ctypes.c_uint8),
                                _p(st.lines, ctypes.c_int32), _p(st.moves, ctypes.c_int32),
                                ctypes.c_int(L), ctypes.c_int(M), _p(feats, ctypes.c_uint8), _p(flags, ctypes.c_uint8),
                                _p(boards, ctypes.c_uint16), ctypes.c_int(nthreads))
    return feats, flags, boards


def features_batch(rows):
    rows = np.ascontiguousarray(rows, np.uint16).reshape(-1, 20)
    out = np.zeros((rows.shape[0], 3), np.uint8)
    lib().orc_features_batch(ctypes.c_int(rows.shape[0]), _p(rows, ctypes.c_uint16), _p(out, ctypes.c_uint8))
    return out


def gen_pieces(seed: int, env_base: int, n: int, episode: int, count: int):
    out = np.zeros((n, count), np.uint8)
    lib().orc_gen_pieces(ctypes.c_uint64(seed), ctypes.c_uint64(env_base), ctypes.c_int(n),
                         ctypes.c_uint32(episode), ctypes.c_int(count), _p(out, ctypes.c_uint8))
    return out


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return tuple(int(x) for x in o)


def rollout(st: BatchState, env_base: int, seed: int, L: int, M: int, pool_rows, pool_pieces, pool_np,
            steps: int, first_reset: bool, episode=None, tstep=None, nthreads: int = 1, weights=None):
    """Random-agent (weights None) or greedy (int32 weights[6]) rollout with auto-reset."""
    pool_rows = np.ascontiguousarray(pool_rows, np.uint16)
    K = pool_rows.shape[0]
    pp = np.zeros((K, st.P), np.uint8)
    pool_pieces = np.asarray(pool_pieces, np.uint8)
    pp[:, :pool_pieces.shape[1]] = pool_pieces
    pool_np = np.ascontiguousarray(pool_np, np.uint8)
    if episode is None:
        episode = np.zeros(st.n, np.uint32)
    if tstep is None:
        tstep = np.zeros(st.n, np.uint32)
    stats = np.zeros(8, np.int64)
    wts = np.ascontiguousarray(weights, np.int32) if weights is not None else None
    lib().orc_rollout(ctypes.c_int(st.n), ctypes.c_uint64(env_base), ctypes.c_uint64(seed),
                             ctypes.c_int(L), ctypes.c_int(M), _p(pool_rows, ctypes.c_uint16),
                             _p(pp, ctypes.c_uint8), ctypes.c_int(st.P), _p(pool_np, ctypes.c_uint8), ctypes.c_int(K),
                             ctypes.c_int(steps), ctypes.c_int(1 if first_reset else 0),
                             _p(st.rows, ctypes.c_uint16), _p(st.pieces, ctypes.c_uint8), _p(st.npieces, ctypes.c_uint8),
                             _p(st.head, ctypes.c_uint8), _p(st.lines, ctypes.c_int32), _p(st.moves, ctypes.c_int32),
                             _p(st.state, ctypes.c_int8), _p(episode, ctypes.c_uint32), _p(tstep, ctypes.c_uint32),
                             _p(stats, ctypes.c_int64), ctypes.c_int(nthreads), _p(wts, ctypes.c_int32))
    return episode, tstep, stats


def reset_done(st: BatchState, env_base: int, seed: int, pool_rows, pool_pieces, pool_np, episode, tstep=None) -> int:
    """Auto-reset of the envs whose episode has ended (TPL_RESET_DONE): episode bumped, config drawn by the counter RNG."""
    pool_rows = np.ascontiguousarray(pool_rows, np.uint16)
    K = pool_rows.shape[0]
    pp = np.zeros((K, st.P), np.uint8)
    pool_pieces = np.asarray(pool_pieces, np.uint8)
    pp[:, :pool_pieces.shape[1]] = pool_pieces
    pool_np = np.ascontiguousarray(pool_np, np.uint8)
    return int(lib().orc_reset_done(ctypes.c_int(st.n), ctypes.c_uint64(env_base), ctypes.c_uint64(seed), _p(pool_rows, ctypes.c_uint16),
                                    _p(pp, ctypes.c_uint8), ctypes.c_int(st.P), _p(pool_np, ctypes.c_uint8), ctypes.c_int(K),
                                    _p(st.rows, ctypes.c_uint16), _p(st.pieces, ctypes.c_uint8), _p(st.npieces, ctypes.c_uint8),
                                    _p(st.head, ctypes.c_uint8), _p(st.lines, ctypes.c_int32), _p(st.moves, ctypes.c_int32),
                                    _p(st.state, ctypes.c_int8), _p(episode, ctypes.c_uint32), _p(tstep, ctypes.c_uint32)))


rollout_random = rollout
