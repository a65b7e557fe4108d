"""The distinct-placements ("alias-free") afterstate form: numbering tables and the expansion back to the 40-slot grid.

The reference's action grid (rotations 0..3 x location 0..9) aliases: ``get_tetromino`` reduces the rotation with
``rot % n_rot`` (``game/tetris.py:61``) and ``move`` clamps the location to ``10 - width`` (``:364``), so only 9 (O),
17 (I, S, Z) or 34 (L, J, T) of the 40 slots are different placements -- 23.1 on average.  ``tpl_afterstates_distinct`` /
``tpl_step_observe_distinct`` write exactly those, one ``uint32`` word each (byte 0 = rows cleared | flags << 3, then
holes, bumpiness, aggregate height), the placements of one env contiguous and ordered by rotation, then column.  Where an
env's run starts -- and which piece it is for, hence its length and the (rot, loc) of every placement -- is reported in a
per-env run descriptor.  The tables come from the library itself (``tpl_distinct_tables``)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

DISTINCT_MAX = 34
FLAG_ALIAS, FLAG_NOPIECE = 8, 16
_tables = None


def capacity(n: int) -> int:
    """``TPL_DISTINCT_CAPACITY(n)``: words a rows array for n envs must hold."""
    return 34 * n + 4 * ((n + 31) // 32)


def tables():
    """(count uint8[8], slot_of uint8[8, 34], canon_of uint8[8, 40]); row 7 = "no piece" (count 0)."""
    global _tables
    if _tables is None:
        count, slot_of, canon_of = np.zeros(8, np.uint8), np.full((8, DISTINCT_MAX), 255, np.uint8), np.zeros((8, 40), np.uint8)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731
        _lib.lib().tpl_distinct_tables(p(count), p(slot_of), p(canon_of))
        _tables = (count, slot_of, canon_of)
    return _tables


def run_offset(runs):
    return runs & 0x1FFFFFFF


def run_piece(runs):
    return runs >> 29


def expand(rows: np.ndarray, runs: np.ndarray) -> np.ndarray:
    """rows uint32[...], runs uint32[n] -> the compact 40-slot form uint8[n, 40, 4] (byte 0 = rows cleared | flags << 3,
    FLAG_ALIAS set on the slots that repeat an earlier one) -- what ``tpl_afterstates`` writes for the same states."""
    count, slot_of, canon_of = tables()
    rows = np.ascontiguousarray(rows).view(np.uint32).reshape(-1)
    runs = np.ascontiguousarray(runs).view(np.uint32)
    piece, off = (runs >> 29).astype(np.int64), (runs & 0x1FFFFFFF).astype(np.int64)
    canon = canon_of[piece].astype(np.int64)                                   # [n, 40]
    has = (piece < 7)[:, None]
    idx = np.where(has, off[:, None] + canon, 0)
    if len(rows) == 0:                                   # no env has a piece left (every queue is empty): nothing was written
        rows = np.zeros(1, np.uint32)
    words = np.where(has, rows[np.minimum(idx, len(rows) - 1)], np.uint32(FLAG_NOPIECE << 3))
    alias = has & (slot_of[piece[:, None], canon] != np.arange(40)[None, :])
    words = words | (alias.astype(np.uint32) * np.uint32(FLAG_ALIAS << 3))
    return np.ascontiguousarray(words.astype(np.uint32)).view(np.uint8).reshape(len(runs), 40, 4)


def gather_index(runs, device=None):
    """torch helper for a value net on the device: runs (uint32 / int32 CUDA tensor [n]) -> (idx int64[n, 34], valid bool[n, 34],
    slot int64[n, 34]): ``rows[idx]`` are the placements of each env (padded), ``slot`` their rot * 10 + loc."""
    import torch
    count, slot_of, _ = tables()
    dev = runs.device if device is None else device
    r = runs.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    piece, off = r >> 29, r & 0x1FFFFFFF
    j = torch.arange(DISTINCT_MAX, device=dev)
    cnt = torch.as_tensor(count.astype(np.int64), device=dev)[piece]
    valid = j[None, :] < cnt[:, None]
    idx = torch.where(valid, off[:, None] + j[None, :], off[:, None].expand(-1, DISTINCT_MAX) * 0)
    slot = torch.as_tensor(slot_of.astype(np.int64), device=dev)[piece]
    return idx, valid, slot
