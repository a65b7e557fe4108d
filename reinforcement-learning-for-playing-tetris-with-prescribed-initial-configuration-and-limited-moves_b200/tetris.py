"""Drop-in single-env facade with the reference's exact ``Tetris`` interface (``game/tetris.py:140-470``).

``Tetris(L, M, warm_reset=True, render=False, framerate=30, debug=False)`` with ``move(rotations, location)``,
``get_state()``, ``reset()``, ``terminate()`` and the public mutable attributes ``board`` (bool[20,10]),
``pieces`` (list), ``state`` (None/True/False), ``lines_cleared``, ``moves_used``, ``L``, ``M`` (and ``solution``
when ``debug``) -- the callers at ``game/main.py:49-57``, ``game/performance_test.py:13-17`` and the intended
``model/train.py:6`` run against it unmodified.  Every ``move`` is executed by the CUDA library through the
host-buffer C ABI (a 1-env batch): the host attributes are uploaded, ``tpl_env_move`` runs, the result is read
back.  This is the parity/drop-in tool; throughput lives in ``BatchedTetris``.

Reset points: by default they come from the native carving generator (``csrc/carve_gen.cpp``) fed with Python's
GLOBAL ``random`` state, so ``random.seed(k); Tetris(L, M, warm_reset=False, debug=True)`` yields the reference's
board, pieces and ``solution`` for that seed and leaves ``random`` where the reference would (``game/tetris.py:226-284``);
``carve``, ``calculate_drop(_deltas)`` and ``RandomPieceGenerator`` are provided with the reference's signatures, so the
reference's own test file (``game/main.py``) runs against this module.  ``config_pool=`` switches to drawing from a
pool with the counter RNG instead.  Differences, all outside the hot path: ``warm_reset=True``
starts no worker processes (configs are generated on demand, in microseconds-to-milliseconds); the reference's other
producer, the forward generator + solver (``game/tetris_algo_main``), is available as ``configs.forward_pool`` /
``forward_games``; ``render=True`` is not supported (pygame UI).
Like the reference, ``reset()`` does NOT zero ``lines_cleared``/``moves_used``/``state`` (``:438-443``; only
the constructor does, ``:149-151``); pass ``fresh=True`` to get an RL-style reset.
"""
from __future__ import annotations

from typing import Optional

import random

import numpy as np

from .configs import (ConfigPool, bool_from_rows, carve_apply, carve_one_from_global_random, rng_words, rows_from_bool,
                      STREAM_CONFIG, MAX_PIECES)
from .host_env import HostBatchedTetris

_ROW_MASKS = (
    ((0xF,), (1, 1, 1, 1)),
    ((4, 7), (3, 2, 2), (7, 1), (1, 1, 3)),
    ((1, 7), (2, 2, 3), (7, 4), (3, 1, 1)),
    ((2, 7), (2, 3, 2), (7, 2), (1, 3, 1)),
    ((6, 3), (1, 3, 2)),
    ((3, 6), (2, 3, 1)),
    ((3, 3),),
)


def _entry(masks):
    w = max(m.bit_length() for m in masks)
    shape = np.array([[bool((m >> j) & 1) for j in range(w)] for m in masks], dtype=bool)
    profile = tuple(max(i for i, m in enumerate(masks) if (m >> j) & 1) for j in range(w))
    return shape, profile


# same structure as the reference's table (``game/tetris.py:23-57``): tetrominos[piece][rot] = (shape, profile)
tetrominos = tuple(tuple(_entry(m) for m in fam) for fam in _ROW_MASKS)


def get_tetromino(piece: int, rotations: int):
    fam = tetrominos[piece]
    return fam[rotations % len(fam)]


def _state_to_code(state) -> int:
    return 0 if state is None else (1 if state else 2)


class RandomPieceGenerator:
    """7-bag piece source with the reference's interface (``game/tetris.py:64-108``), on the global ``random`` stream:
    ``get_random_piece() -> ((piece, index), regenerated)`` draws without removing (``delete_index`` removes),
    ``get_random_sequence(n)`` concatenates shuffled bags, starting with whatever is left of the current one."""

    def __init__(self) -> None:
        self.pieces = []

    def generate_pieces(self) -> None:
        self.pieces = list(range(7))

    def _refill(self) -> bool:
        if self.pieces:
            return False
        self.generate_pieces()
        return True

    def get_random_piece(self):
        regenerated = self._refill()
        index = random.randint(0, len(self.pieces) - 1)
        return (self.pieces[index], index), regenerated

    def delete_index(self, index) -> None:
        del self.pieces[index]

    def get_random_sequence(self, length: int):
        out = []
        while len(out) < length:
            self._refill()
            random.shuffle(self.pieces)
            out.extend(self.pieces[:min(length - len(out), 7)])
            self.pieces = []
        return out

    def reset(self) -> None:
        self.pieces.clear()

    def __len__(self) -> int:
        return len(self.pieces)


class Tetris:
    def __init__(self, L: int, M: int, warm_reset: bool = True, render: bool = False, framerate: int = 30,
                 debug: bool = False, *, config_pool: Optional[ConfigPool] = None, seed: int = 0, device: int = 0):
        if render:
            raise NotImplementedError("render=True (pygame window) is outside the B200 hot path")
        self.L, self.M = L, M
        self.warm_reset = warm_reset
        self.render = False
        self.lines_cleared = 0
        self.moves_used = 0
        self.state = None
        self.debug = debug
        if debug:
            self.solution = []
        self.random_piece_generator = RandomPieceGenerator()
        self.board = np.full((20, 10), False, dtype=bool)
        self.pieces = []
        self._seed = seed
        self._device = device
        self._episode = 0
        self._pool = config_pool
        self._env_handle = None                 # the CUDA handle is created by the first call that needs the GPU
        self.load_warm_reset()

    @property
    def _env(self) -> HostBatchedTetris:
        if self._env_handle is None:
            self._env_handle = HostBatchedTetris(1, self.L, self.M, device=self._device, seed=self._seed)
        return self._env_handle

    # -- reset (game/tetris.py:438-449) ---------------------------------------------------------------
    def reset(self, fresh: bool = False) -> None:
        self.board[:, :] = False
        self.pieces.clear()
        if fresh:
            self.lines_cleared, self.moves_used, self.state = 0, 0, None
        self.load_warm_reset()

    def load_warm_reset(self) -> None:
        if self._pool is None:                  # the reference's own generator, on the global `random` stream
            self._generate_initial_config()
            return
        w0 = rng_words(self._seed, np.array([0], np.uint64), self._episode, STREAM_CONFIG, 0)[0]
        k = int((int(w0[0]) * self._pool.K) >> 32)
        self._episode += 1
        self.board = bool_from_rows(self._pool.rows[k])
        self.pieces = [int(p) for p in self._pool.pieces[k, :int(self._pool.npieces[k])]]
        if self.debug:
            sol = self._pool.solutions
            self.solution = ([(int(r), int(c)) for r, c in sol[k, :int(self._pool.nsol[k])]] if sol is not None else [])

    # -- carving generator (game/tetris.py:226-352), native, bit-identical to the reference -----------------
    def _generate_initial_config(self) -> None:
        rows, pieces, solution = carve_one_from_global_random(self.L, self.M)
        self.board = bool_from_rows(rows)
        self.pieces = pieces
        if self.debug:
            self.solution = solution

    def carve(self, piece: int, rotations: int, location: int, allow_partial: bool) -> bool:
        rows = rows_from_bool(self.board)
        ok = carve_apply(rows, piece, rotations, location, allow_partial)
        if ok:
            self.board[:, :] = bool_from_rows(rows)
        return ok

    def calculate_drop_deltas(self, location, reverse_tetromino_topography, tetromino_width):
        """Column tops (first filled row from the top, else 20) minus the shape's bottom profile (:427-433)."""
        cols = self.board[:, location:location + tetromino_width]
        tops = np.where(cols.any(axis=0), cols.argmax(axis=0), 20)
        return tops - np.asarray(reverse_tetromino_topography)

    def calculate_drop(self, drop_deltas) -> int:
        return int(min(drop_deltas)) - 1

    # -- move (game/tetris.py:354-422) ----------------------------------------------------------------
    def move(self, rotations: int, location: int) -> None:
        if not self.pieces:
            raise IndexError("pop from empty list")
        if location < 0:
            raise ValueError("negative location")
        if not 0 <= int(self.pieces[0]) <= 6:                 # the reference indexes `tetrominos[piece]` (:60-61)
            raise IndexError("tuple index out of range")
        dl, fl = self._run_move(rotations, location)
        f = self._env.fields(queue=False)
        self.pieces.pop(0)
        new_board = bool_from_rows(f["rows"][0])
        if dl > 0:
            self.board = new_board                     # the reference rebinds the array on a clear (:405-407)
        else:
            self.board[:, :] = new_board
        self.lines_cleared += dl
        self.moves_used = int(f["moves"][0])
        if fl & 2:
            self.state = True
        elif fl & (1 | 4):
            self.state = False

    def _sync_limits(self):
        if (self.L, self.M) != (self._env.L, self._env.M):          # L and M are public mutable attributes (:143-144)
            self._env.set_limits(self.L, self.M)

    def _run_move(self, rotations, location):
        self._sync_limits()
        n = min(len(self.pieces), MAX_PIECES)
        p = np.zeros((1, MAX_PIECES), np.uint8)
        # `pieces` is a public mutable list: an id outside 0..6 further down the queue is harmless until it reaches the front
        # (move() raises there like the reference); the upload itself only takes valid ids
        p[0, :n] = [q if 0 <= q <= 6 else 0 for q in self.pieces[:n]]
        self._env.load(rows_from_bool(self.board)[None], p, np.array([n], np.uint8),
                       lines=[min(int(self.lines_cleared), 65535)], moves=[min(int(self.moves_used), 65535)],
                       state=[_state_to_code(self.state)])
        dl, fl, _ = self._env.move([rotations], [min(location, 255)])
        return int(dl[0]), int(fl[0])

    # -- get_state (game/tetris.py:435-436) -----------------------------------------------------------
    def get_state(self):
        return self.board, self.pieces[0], self.pieces[1], self.L - self.lines_cleared, self.M - self.moves_used, self.state

    # -- extensions the north star asks for ------------------------------------------------------------
    def step(self, rotations: int, location: int):
        """``move`` returning (rows cleared by this move, state)."""
        before = self.lines_cleared
        self.move(rotations, location)
        return int(self.lines_cleared - before), self.state

    def afterstates(self):
        """(feats uint8[4,10,4] = (rows cleared, holes, bumpiness, aggregate height), flags uint8[4,10])."""
        self._sync_limits()
        if self.pieces and not 0 <= int(self.pieces[0]) <= 6:
            raise IndexError("tuple index out of range")
        n = min(len(self.pieces), MAX_PIECES)
        p = np.zeros((1, MAX_PIECES), np.uint8)
        p[0, :n] = [q if 0 <= q <= 6 else 0 for q in self.pieces[:n]]
        self._env.load(rows_from_bool(self.board)[None], p, np.array([n], np.uint8),
                       lines=[min(int(self.lines_cleared), 65535)], moves=[min(int(self.moves_used), 65535)],
                       state=[_state_to_code(self.state)])
        feats, flags = self._env.afterstates()
        return feats[0].copy(), flags[0].copy()

    def next_states(self) -> dict:
        """{(rot, loc): (rows cleared, holes, bumpiness, aggregate height)} over the distinct placements."""
        feats, flags = self.afterstates()
        return {(r, c): tuple(int(v) for v in feats[r, c]) for r in range(4) for c in range(10)
                if not flags[r, c] & (8 | 16)}

    def terminate(self):
        if self._env_handle is not None:
            self._env_handle.close()
            self._env_handle = None
