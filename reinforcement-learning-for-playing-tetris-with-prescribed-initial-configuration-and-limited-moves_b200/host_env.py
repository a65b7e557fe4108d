"""``HostBatchedTetris``: the host-buffer side of the C ABI (``tpl_env_*``).

Every argument and result is a numpy array in host memory; the handle inside the library owns the device
state, a stream and staging buffers.  This is the binding a maintainer would add to the reference's
``game/tetris.py`` and the call path ``bench.py`` times for its end-to-end figure.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .configs import MAX_PIECES, ConfigPool

RESET_ALL, RESET_MASK, RESET_DONE = 0, 1, 2


def _p(a: Optional[np.ndarray]):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class PinnedArray:
    """A numpy view of page-locked host memory from ``tpl_host_alloc`` (freed with the object)."""

    def __init__(self, shape, dtype):
        self._L = _lib.lib()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self._L.tpl_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("tpl_host_alloc failed")
        buf = (ctypes.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self._L.tpl_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class HostBatchedTetris:
    def __init__(self, num_envs: int, L: int, M: int, device: int = 0, seed: int = 0, env_base: int = 0,
                 config_pool: Optional[ConfigPool] = None):
        self._L = _lib.lib()
        self.num_envs, self.L, self.M = int(num_envs), int(L), int(M)
        self._h = ctypes.c_void_p()
        _lib.check(self._L.tpl_env_create(ctypes.byref(self._h), self.num_envs, self.L, self.M, int(device), int(seed),
                                          int(env_base)), "tpl_env_create")
        if config_pool is not None:
            self.set_pool(config_pool)

    def close(self):
        if self._h:
            self._L.tpl_env_destroy(self._h)
            self._h = ctypes.c_void_p()

    terminate = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_limits(self, L: int, M: int) -> None:
        _lib.check(self._L.tpl_env_set_limits(self._h, int(L), int(M)), "tpl_env_set_limits")
        self.L, self.M = int(L), int(M)

    def set_pool(self, pool: ConfigPool) -> None:
        rows = np.ascontiguousarray(pool.rows, np.uint16)
        pieces = np.ascontiguousarray(pool.pieces, np.uint8)
        npieces = np.ascontiguousarray(pool.npieces, np.uint8)
        if int(pieces.max(initial=0)) > 6 or int(npieces.max(initial=0)) > min(MAX_PIECES, pieces.shape[1]):
            raise ValueError("pool: piece ids must be 0..6 and npieces <= min(42, pieces per row)")
        _lib.check(self._L.tpl_env_set_pool(self._h, rows.shape[0], _p(rows), _p(pieces), pieces.shape[1], _p(npieces)),
                   "tpl_env_set_pool")

    def reset(self, idx=None, mask=None, done_only: bool = False, gen_pieces: int = 0) -> None:
        idx = np.ascontiguousarray(idx, np.int32) if idx is not None else None
        mask = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        mode = RESET_DONE if done_only else (RESET_MASK if mask is not None else RESET_ALL)
        _lib.check(self._L.tpl_env_reset(self._h, _p(idx), _p(mask), mode, int(gen_pieces)), "tpl_env_reset")

    def load(self, rows, pieces, npieces, lines=None, moves=None, state=None, head=None) -> None:
        n = self.num_envs
        rows = np.ascontiguousarray(rows, np.uint16).reshape(n, 20)
        pieces = np.ascontiguousarray(pieces, np.uint8).reshape(n, -1)
        npieces = np.ascontiguousarray(npieces, np.uint8).reshape(n)
        if int(npieces.max(initial=0)) > min(MAX_PIECES, pieces.shape[1]):
            raise ValueError("npieces exceeds the 42-piece queue or the pieces array")
        if int(pieces.max(initial=0)) > 6:
            raise ValueError("piece ids must be 0..6")
        c = lambda x, dt: np.ascontiguousarray(x, dt).reshape(n) if x is not None else None   # noqa: E731
        lines, moves, state, head = c(lines, np.int32), c(moves, np.int32), c(state, np.int8), c(head, np.uint8)
        _lib.check(self._L.tpl_env_load(self._h, _p(rows), _p(pieces), pieces.shape[1], _p(npieces), _p(lines), _p(moves),
                                        _p(state), _p(head)), "tpl_env_load")

    def move(self, rot, loc):
        n = self.num_envs
        la = np.asarray(loc, np.int64).reshape(n)
        if (la < 0).any():
            raise ValueError("location must be >= 0")
        r = np.mod(np.asarray(rot, np.int64).reshape(n), 4).astype(np.uint8)
        l = np.minimum(la, 255).astype(np.uint8)
        dl, fl, st = np.empty(n, np.int8), np.empty(n, np.uint8), np.empty(n, np.int8)
        _lib.check(self._L.tpl_env_move(self._h, _p(r), _p(l), _p(dl), _p(fl), _p(st)), "tpl_env_move")
        return dl, fl, st

    def feats_device_ptr(self) -> int:
        """Device address of the uint8[40, n, 4] afterstate words written by the last ``step_observe``."""
        return int(self._L.tpl_env_feats_ptr(self._h) or 0)

    def fields(self, queue: bool = True) -> dict:
        n = self.num_envs
        out = dict(rows=np.empty((n, 20), np.uint16), cur=np.empty(n, np.uint8), next=np.empty(n, np.uint8),
                   lines=np.empty(n, np.int32), moves=np.empty(n, np.int32), state=np.empty(n, np.int8),
                   head=np.empty(n, np.uint8), npieces=np.empty(n, np.uint8))
        q = np.empty((n, MAX_PIECES), np.uint8) if queue else None
        _lib.check(self._L.tpl_env_get_state(self._h, _p(out["rows"]), _p(out["cur"]), _p(out["next"]), _p(out["lines"]),
                                             _p(out["moves"]), _p(out["state"]), _p(out["head"]), _p(out["npieces"]), _p(q)),
                   "tpl_env_get_state")
        if queue:
            out["queue"] = q
        return out

    def get_state(self):
        f = self.fields(queue=False)
        return f["rows"], f["cur"], f["next"], self.L - f["lines"], self.M - f["moves"], f["state"]

    def afterstates(self):
        n = self.num_envs
        feats, flags = np.empty((40, n, 4), np.uint8), np.empty((40, n), np.uint8)
        _lib.check(self._L.tpl_env_afterstates(self._h, _p(feats), _p(flags)), "tpl_env_afterstates")
        return feats.reshape(4, 10, n, 4).transpose(2, 0, 1, 3), flags.reshape(4, 10, n).transpose(2, 0, 1)

    def distinct_capacity(self) -> int:
        """words the ``rows`` buffer of ``step_observe_distinct`` must hold (chunk regions included)"""
        return int(self._L.tpl_env_distinct_capacity(self._h))

    def chunks(self) -> int:
        return int(self._L.tpl_env_chunks(self._h))

    def step_observe_distinct(self, rot: np.ndarray, loc: np.ndarray, dlines: np.ndarray, flags: np.ndarray, state: np.ndarray,
                              rows: np.ndarray, runs: np.ndarray) -> int:
        """``step_observe`` with the afterstates in the distinct-placements form: ``rows`` uint32[distinct_capacity()],
        ``runs`` uint32[N] (see ``distinct.expand``).  Returns the number of ``rows`` words that crossed PCIe."""
        words = ctypes.c_int64(0)
        _lib.check(self._L.tpl_env_step_observe_distinct(self._h, _p(rot), _p(loc), _p(dlines), _p(flags), _p(state), _p(rows),
                                                         rows.size, _p(runs), ctypes.byref(words)), "tpl_env_step_observe_distinct")
        return int(words.value)

    def step_observe(self, rot: np.ndarray, loc: np.ndarray, dlines: np.ndarray, flags: np.ndarray, state: np.ndarray,
                     feats: Optional[np.ndarray], aflags: Optional[np.ndarray]) -> None:
        """One host-facing rollout step into caller-provided (ideally pinned) uint8/int8 buffers:
        H2D actions -> move -> auto-reset finished envs -> afterstates -> D2H.  ``aflags=None`` selects the
        compact form (feats byte 0 = rows cleared | flags << 3); ``feats=None`` leaves the features on the device
        (``feats_device_ptr()``) for a policy that runs there."""
        _lib.check(self._L.tpl_env_step_observe(self._h, _p(rot), _p(loc), _p(dlines), _p(flags), _p(state), _p(feats), _p(aflags)),
                   "tpl_env_step_observe")
