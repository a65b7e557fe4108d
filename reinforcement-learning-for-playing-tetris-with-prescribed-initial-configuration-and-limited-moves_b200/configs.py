"""Prescribed-configuration pools: the ``(board, pieces)`` reset points the envs are reset from.

The reference produces reset points with two slow CPU generators (carving, ``game/tetris.py:226-352``; forward
generator + solver, ``game/tetris_algo_main/``) and hands them over as ``(board bool[20,10], pieces list)``
tuples (``game/tetris.py:476-479``, consumed at ``:447``).  Here a pool is three arrays in the boundary format:

    rows    uint16[K, 20]   bitrows, bit c = column c, row 0 = top
    pieces  uint8[K, P]     piece ids 0..6 (0=I 1=L 2=J 3=T 4=S 5=Z 6=O), P <= 42
    npieces uint8[K]        valid entries per row (the reference supplies M + 1)

Sources: ``load_pool(path)`` reads an ``.npz`` written by the carve generator (the committed fixture
``tests/golden/carve_pool_L10_M30.npz`` was produced by the unmodified reference), and ``synthetic_pool``
generates the synthetic boards of SURVEY.md section 8d, config 3, from a counter-based RNG so a pool is a pure
function of ``(K, seed, M)`` on every machine.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

MAX_PIECES = 42
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

STREAM_PIECES, STREAM_ACTION, STREAM_CONFIG, STREAM_BOARD = 0, 1, 2, 3


class ConfigPool(NamedTuple):
    rows: np.ndarray
    pieces: np.ndarray
    npieces: np.ndarray
    solutions: Optional[np.ndarray] = None     # int8[K, S, 2] (rot, loc), -1 padded; only carve pools have it
    nsol: Optional[np.ndarray] = None

    @property
    def K(self) -> int:
        return int(self.rows.shape[0])


def philox4x32(c0, c1, c2, c3, k0: int, k1: int, rounds: int = 10):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): counters are uint32 arrays, key two ints."""
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) & _MASK for x in (c0, c1, c2, c3))
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> _S32) ^ c1 ^ np.uint64(k0)) & _MASK, p1 & _MASK, \
                         ((p0 >> _S32) ^ c3 ^ np.uint64(k1)) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def rng_words(seed: int, ids, episode, stream: int, index):
    ids = np.asarray(ids, np.uint64)
    episode = np.broadcast_to(np.asarray(episode, np.uint64), ids.shape)
    index = np.broadcast_to(np.asarray(index, np.uint64), ids.shape)
    c3 = (np.uint64(stream & 0xF) << np.uint64(28)) | (index & np.uint64(0x0FFFFFFF))
    return philox4x32(ids & _MASK, ids >> _S32, episode, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def bags_from_words(u) -> np.ndarray:
    """uint32 words -> one permutation of 0..6 each (k = mulhi(u, 5040); Fisher-Yates on k's
    factorial-base digits), shape [..., 7]."""
    u = np.asarray(u, np.uint64)
    k = (u * np.uint64(5040)) >> _S32
    perm = np.broadcast_to(np.arange(7, dtype=np.uint8), u.shape + (7,)).copy()
    flat = perm.reshape(-1, 7)
    kf = k.reshape(-1)
    ar = np.arange(flat.shape[0])
    for i in range(6, 0, -1):
        j = (kf % np.uint64(i + 1)).astype(np.int64)
        kf = kf // np.uint64(i + 1)
        a = flat[ar, i].copy()
        flat[ar, i] = flat[ar, j]
        flat[ar, j] = a
    return flat.reshape(u.shape + (7,))


GEN_MAX = MAX_PIECES * 256       # longest generated sequence (the kernels refill the 42-piece queue block by block)


def gen_pieces(seed: int, env_ids, episode, count: int) -> np.ndarray:
    """Counter-based 7-bag sequences uint8[n, count]: concatenated bags, truncated, the contract of
    RandomPieceGenerator.get_random_sequence (``game/tetris.py:95-102``).  Bag g comes from word g & 3 of Philox call g >> 2."""
    if not 0 <= count <= GEN_MAX:
        raise ValueError(f"count must be in 0..{GEN_MAX}")
    env_ids = np.asarray(env_ids, np.uint64).reshape(-1)
    nbags = max((count + 6) // 7, 1)
    words = []
    for call in range((nbags + 3) // 4):
        w = rng_words(seed, env_ids, episode, STREAM_PIECES, call)
        words += [w[0], w[1], w[2], w[3]]
    words = np.stack(words[:nbags], axis=1)                                     # [n, nbags]
    return bags_from_words(words).reshape(len(env_ids), 7 * nbags)[:, :count].astype(np.uint8)


def synthetic_pool(K: int, seed: int = 0, M: int = 30, max_height: int = 12, fill_byte: int = 154) -> ConfigPool:
    """height H ~ U{0..max_height}; rows 20-H..19 i.i.d. 10-bit with
    P(bit) = fill_byte/256 (154/256 = 0.6016), resampled while the row is empty or full; rows above empty;
    pieces = M+1 pieces of the counter-based 7-bag stream of (seed, config id, episode 0xB0A2D)."""
    if M + 1 > MAX_PIECES:
        raise ValueError("a pool stores explicit piece lists of at most 42 pieces (M <= 41); for longer episodes take the boards from "
                         "any pool and let the kernels generate the pieces: BatchedTetris(..., gen_pieces=M + 1)")
    ids = np.arange(K, dtype=np.uint64)
    hw = rng_words(seed, ids, 0, STREAM_BOARD, 0x0FFFFFFF)[0]
    H = ((hw * np.uint64(max_height + 1)) >> _S32).astype(np.int64)
    rows = np.zeros((K, 20), np.uint16)
    for i in range(max_height):                       # i-th row above the floor
        chosen = np.zeros(K, np.uint16)
        done = np.zeros(K, bool)
        for attempt in range(8):
            a = rng_words(seed, ids, 0, STREAM_BOARD, (i * 8 + attempt) * 2)
            b = rng_words(seed, ids, 0, STREAM_BOARD, (i * 8 + attempt) * 2 + 1)
            by = []
            for w in (a[0], a[1], b[0]):
                for s in range(4):
                    by.append((w >> np.uint64(8 * s)) & np.uint64(0xFF))
            row = np.zeros(K, np.uint16)
            for c in range(10):
                row |= ((by[c] < np.uint64(fill_byte)).astype(np.uint16) << np.uint16(c))
            ok = (row != 0) & (row != 0x3FF) & ~done
            chosen[ok] = row[ok]
            done |= ok
        rows[:, 19 - i] = np.where(i < H, chosen, 0)
    pieces = np.zeros((K, MAX_PIECES), np.uint8)
    pieces[:, :M + 1] = gen_pieces(seed, ids, 0xB0A2D, M + 1)
    return ConfigPool(rows, pieces, np.full(K, M + 1, np.uint8))


_carve_lib = None


def _carve():
    global _carve_lib
    if _carve_lib is None:
        import ctypes
        from . import build as _build
        lib = ctypes.CDLL(_build.build_carve())
        V = ctypes.c_void_p
        lib.carve_generate.restype = ctypes.c_int
        lib.carve_generate.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, V, V, ctypes.c_int, V, V, V, ctypes.c_int]
        lib.carve_generate_from_state.restype = ctypes.c_int
        lib.carve_generate_from_state.argtypes = [V, ctypes.c_int, ctypes.c_int, V, V, ctypes.c_int, V, V, V]
        lib.carve_apply.restype = ctypes.c_int
        lib.carve_apply.argtypes = [V, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.forward_generate.restype = ctypes.c_int
        lib.forward_generate.argtypes = [ctypes.c_uint64] + [ctypes.c_int] * 5 + [V] * 6 + [ctypes.c_int]
        _carve_lib = lib
    return _carve_lib


def carve_one_from_global_random(L: int, M: int):
    """One prescribed config drawn from Python's GLOBAL ``random`` stream exactly as the reference's
    ``Tetris(L, M, warm_reset=False)`` draws it (``game/tetris.py:226-284``): the MT19937 state is handed to the native
    generator and the advanced state is put back, so ``random.seed(k)`` before the call reproduces the reference's
    board, pieces and solution, and later ``random`` calls continue where the reference's would.
    Returns (rows uint16[20], pieces list[int], solution list[(rot, loc)])."""
    import ctypes
    import random
    ver, internal, gauss = random.getstate()
    st = np.array(internal, dtype=np.uint32)                      # 624 words + index
    rows = np.zeros(20, np.uint16)
    pieces = np.zeros(max(M + 1, 1), np.uint8)
    npieces, nsol = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    sol = np.full((max(M, 1), 2), -1, np.int8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)                  # noqa: E731
    if _carve().carve_generate_from_state(p(st), L, M, p(rows), p(pieces), len(pieces), p(npieces), p(sol), p(nsol)) != 0:
        raise ValueError("native carve generator rejected its arguments (1 <= L <= 16, M >= 1)")
    random.setstate((ver, tuple(int(x) for x in st), gauss))
    return rows, [int(x) for x in pieces[:int(npieces[0])]], [(int(r), int(c)) for r, c in sol[:int(nsol[0])]]


def carve_apply(rows: np.ndarray, piece: int, rotations: int, location: int, allow_partial: bool) -> bool:
    """``Tetris.carve`` (``game/tetris.py:286-311``) on uint16[20] bitrows, in place."""
    import ctypes
    rc = _carve().carve_apply(ctypes.c_void_p(rows.ctypes.data), int(piece), int(rotations), int(location), int(bool(allow_partial)))
    if rc < 0:
        raise ValueError("carve: piece/location out of range")
    return rc == 1


def carve_pool(K: int, L: int, M: int, seed0: int = 0, threads: Optional[int] = None, with_solutions: bool = True) -> ConfigPool:
    """K prescribed configs from the native carving generator (csrc/carve_gen.cpp): config k is exactly what the
    reference produces for random.seed(seed0 + k); Tetris(L, M, warm_reset=False, debug=True)
    (game/tetris.py:226-352) -- board, the M+1 pieces and the recorded solution -- at native speed, threaded."""
    import ctypes
    import os
    from . import build as _build
    if M + 1 > MAX_PIECES:
        raise ValueError("a pool stores explicit piece lists of at most 42 pieces (M <= 41); for longer episodes take the boards from "
                         "any pool and let the kernels generate the pieces: BatchedTetris(..., gen_pieces=M + 1)")
    lib = ctypes.CDLL(_build.build_carve())
    lib.carve_generate.restype = ctypes.c_int
    lib.carve_generate.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 2 +         [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int]
    rows = np.zeros((K, 20), np.uint16)
    pieces = np.zeros((K, MAX_PIECES), np.uint8)
    npieces = np.zeros(K, np.uint8)
    sols = np.full((K, M, 2), -1, np.int8) if with_solutions else None
    nsol = np.zeros(K, np.uint8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None       # noqa: E731
    rc = lib.carve_generate(int(seed0), K, L, M, p(rows), p(pieces), MAX_PIECES, p(npieces), p(sols), p(nsol),
                            int(threads or os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("carve_generate rejected its arguments (1 <= L <= 16, M >= 1)")
    return ConfigPool(rows, pieces, npieces, sols, nsol)


def forward_games(goal: int, tetrominoes: int, seed0: int = 0, count: int = 100, initial_height_max: int = 4,
                  max_attempts: int = 1000, threads: Optional[int] = None) -> dict:
    """The reference's FORWARD producer for seeds seed0 .. seed0+count-1, natively (csrc/forward_gen.cpp):
    ``TetrisGameGenerator(seed, goal, tetrominoes, initial_height_max)`` (``game/tetris_algo_main/TetrisGameGenerator.py``)
    followed by ``TetrisSolver(board, sequence, goal, max_attempts).solve()`` (``TetrisSolver.py:112-163``), i.e.
    ``main.py:generate_game`` / ``solve_game``.  Returns arrays: ``rows`` u16[count, 20], ``letters`` u8[count, tetrominoes]
    (the sequence as piece ids of ``game/tetris.py``, ``piece_translations`` ``:8-16``), ``solvable``, ``failed`` (the solver's
    failed_attempts), ``moves`` i8[count, tetrominoes, 3] (the solver's stack: name index in I J L O S T Z, rotation in the
    generator's own table, column) and ``nmoves``.  ``max_attempts < 0`` generates without solving."""
    import ctypes
    import os
    if not 1 <= tetrominoes <= 255:
        raise ValueError("1 <= tetrominoes <= 255")
    rows = np.zeros((count, 20), np.uint16)
    letters = np.zeros((count, tetrominoes), np.uint8)
    solvable, nmoves = np.zeros(count, np.uint8), np.zeros(count, np.uint8)
    failed = np.zeros(count, np.int32)
    moves = np.full((count, tetrominoes, 3), -1, np.int8)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)                  # noqa: E731
    rc = _carve().forward_generate(int(seed0), count, goal, tetrominoes, initial_height_max, max_attempts, p(rows), p(letters),
                                   p(solvable), p(failed), p(moves), p(nmoves), int(threads or os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("forward_generate rejected its arguments")
    return dict(rows=rows, letters=letters, solvable=solvable, failed=failed, moves=moves, nmoves=nmoves)


def forward_pool(L: int, M: int, start: int = 0, end: int = 100, initial_height_max: int = 4, max_attempts: int = 1000,
                 threads: Optional[int] = None) -> ConfigPool:
    """The reset points ``forward_warm_reset_worker`` puts on the queue (``game/tetris.py:482-488``):
    ``translate(main.generate_batch(L, M))`` -- the winnable games among seeds ``start .. end-1`` (``main.py:35-42, 62-73``),
    each as (board, [random.randint(0, 6)] + sequence).  Like ``translate`` (``:19-20``) the leading piece of every reset
    point is drawn from Python's GLOBAL ``random`` stream, one draw per winnable game in seed order."""
    import random
    if M + 1 > MAX_PIECES:
        raise ValueError("a pool stores explicit piece lists of at most 42 pieces (M <= 41); for longer episodes take the boards from "
                         "any pool and let the kernels generate the pieces: BatchedTetris(..., gen_pieces=M + 1)")
    g = forward_games(L, M, start, end - start, initial_height_max, max_attempts, threads)
    keep = np.flatnonzero(g["solvable"])
    pieces = np.zeros((len(keep), MAX_PIECES), np.uint8)
    for k, j in enumerate(keep):
        pieces[k, 0] = random.randint(0, 6)
        pieces[k, 1:M + 1] = g["letters"][j]
    return ConfigPool(g["rows"][keep].copy(), pieces, np.full(len(keep), M + 1, np.uint8))


def load_pool(path: str) -> ConfigPool:
    z = np.load(path)
    pieces = np.zeros((z["rows"].shape[0], MAX_PIECES), np.uint8)
    pieces[:, :z["pieces"].shape[1]] = z["pieces"]
    sol = z["solutions"] if "solutions" in z.files else None
    nsol = z["nsol"] if "nsol" in z.files else None
    return ConfigPool(z["rows"].astype(np.uint16), pieces, z["npieces"].astype(np.uint8), sol, nsol)


def save_pool(path: str, pool: ConfigPool, **extra) -> None:
    d = dict(rows=pool.rows, pieces=pool.pieces, npieces=pool.npieces)
    if pool.solutions is not None:
        d.update(solutions=pool.solutions, nsol=pool.nsol)
    d.update(extra)
    np.savez_compressed(path, **d)


def concat_pools(*pools: ConfigPool) -> ConfigPool:
    return ConfigPool(np.concatenate([p.rows for p in pools]), np.concatenate([p.pieces for p in pools]),
                      np.concatenate([p.npieces for p in pools]))


def rows_from_bool(board) -> np.ndarray:
    """bool[..., 20, 10] -> uint16[..., 20] bitrows."""
    b = np.asarray(board, bool)
    return (b.astype(np.uint16) << np.arange(10, dtype=np.uint16)).sum(axis=-1).astype(np.uint16)


def bool_from_rows(rows) -> np.ndarray:
    """uint16[..., 20] -> bool[..., 20, 10]."""
    r = np.asarray(rows, np.uint16)
    return ((r[..., None] >> np.arange(10, dtype=np.uint16)) & 1).astype(bool)
