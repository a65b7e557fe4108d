"""Generate the committed golden fixtures from the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything written here comes out of /root/reference/game/tetris.py itself (imported in place via
oracle/refshim.py); the oracle is NOT used to produce any expected value.  Features are computed
here with plain numpy on the reference's bool[20,10] boards, following SURVEY.md section 8a-F.

Fixtures (all small, np.savez_compressed):
  kat_carve.npz        seeded carve configs (random.seed(k); Tetris(L, M, warm_reset=False, debug=True)),
                       their recorded solutions, and the lines/moves/state trace of replaying them
  moves_random.npz     adversarial random-move episodes: state after every move
  afterstates.npz      4x10 afterstate grids composed as clone -> move(r, c), with features and flags
  carve_pool_L10_M30.npz  a pool of prescribed (board, pieces) reset points for (L=10, M=30)
"""
from __future__ import annotations

import copy
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P = 42
T_MAX = 44


def pack_rows(board) -> np.ndarray:
    b = np.asarray(board, bool)
    return (b.astype(np.uint16) << np.arange(10, dtype=np.uint16)).sum(axis=1).astype(np.uint16)


def np_features(board):
    """(holes, bumpiness, aggregate height) straight from a bool[20,10] board."""
    b = np.asarray(board, bool)
    filled = b.any(axis=0)
    top = np.where(filled, b.argmax(axis=0), 20)
    h = 20 - top
    agg = int(h.sum())
    bump = int(np.abs(np.diff(h)).sum())
    holes = int(sum(int(h[c]) - int(b[:, c].sum()) for c in range(10)))
    return holes, bump, agg


def sc(state):
    return refshim.state_code(state)


def adversarial_board(rng):
    H = int(rng.integers(0, 21))
    dens = float(rng.uniform(0.3, 0.9))
    b = np.zeros((20, 10), bool)
    if H:
        b[20 - H:] = rng.random((H, 10)) < dens
    kind = int(rng.integers(0, 6))
    if kind == 0 and H:
        b[int(rng.integers(20 - H, 20))] = True
    elif kind == 1:
        b[0, int(rng.integers(0, 10))] = True
    elif kind == 2 and H >= 3:
        b[20 - H + 1:, int(rng.integers(0, 10))] = False
    return b


def make_kat(tetris):
    cases = [(0, 10, 30), (1, 10, 30), (0, 15, 40), (2, 10, 30), (3, 10, 30), (5, 12, 35)]
    out = {}
    for i, (seed, L, M) in enumerate(cases):
        random.seed(seed)
        g = tetris.Tetris(L, M, warm_reset=False, debug=True)
        rows0 = pack_rows(g.board)
        pieces0 = np.array(g.pieces, np.uint8)
        sol = np.array(g.solution, np.int16)
        trace = []
        for (r, c) in g.solution:
            g.move(r, c)
            trace.append((int(g.lines_cleared), int(g.moves_used), sc(g.state)))
        assert g.state is True            # game/main.py:49-57 test_carving_invertability
        out[f"k{i}_meta"] = np.array([seed, L, M], np.int32)
        out[f"k{i}_rows"] = rows0
        out[f"k{i}_pieces"] = pieces0
        out[f"k{i}_solution"] = sol
        out[f"k{i}_trace"] = np.array(trace, np.int32)
        out[f"k{i}_final_rows"] = pack_rows(g.board)
    out["count"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "kat_carve.npz"), **out)


def make_pool(tetris, K=256, L=10, M=30):
    rows = np.zeros((K, 20), np.uint16)
    pieces = np.zeros((K, P), np.uint8)
    npieces = np.zeros(K, np.uint8)
    sols = np.full((K, M, 2), -1, np.int8)
    nsol = np.zeros(K, np.uint8)
    for k in range(K):
        random.seed(1000 + k)
        g = tetris.Tetris(L, M, warm_reset=False, debug=True)
        rows[k] = pack_rows(g.board)
        npieces[k] = len(g.pieces)
        pieces[k, :len(g.pieces)] = g.pieces
        nsol[k] = len(g.solution)
        sols[k, :len(g.solution)] = g.solution
    np.savez_compressed(os.path.join(HERE, "carve_pool_L10_M30.npz"), rows=rows, pieces=pieces, npieces=npieces,
                        solutions=sols, nsol=nsol, L=np.array(L), M=np.array(M))
    return rows, pieces, npieces, sols, nsol


def make_moves(tetris, pool, E=360):
    rng = np.random.default_rng(7)
    prow, ppieces, pnp, psol, pnsol = pool
    rows0 = np.zeros((E, 20), np.uint16)
    pieces = np.zeros((E, P), np.uint8)
    npieces = np.zeros(E, np.uint8)
    LM = np.zeros((E, 2), np.int32)
    acts = np.zeros((E, T_MAX, 2), np.int16)
    nmoves = np.zeros(E, np.int32)
    a_rows = np.zeros((E, T_MAX, 20), np.uint16)
    a_meta = np.zeros((E, T_MAX, 4), np.int32)          # lines_cleared, moves_used, state, pieces left
    snapshots = []                                      # (episode, move index) -> env copy for afterstates
    for e in range(E):
        if e % 3 == 0:                                  # carve config: solution prefix then random moves
            k = int(rng.integers(0, len(prow)))
            L, M = 10, 30
            board = ((prow[k][:, None] >> np.arange(10)) & 1).astype(bool)
            pcs = [int(x) for x in ppieces[k, :pnp[k]]]
            script = [tuple(int(v) for v in s) for s in psol[k, :int(rng.integers(0, pnsol[k] + 1))]]
        else:
            L = int(rng.integers(1, 16))
            M = int(rng.integers(1, 41))
            board = adversarial_board(rng)
            pcs = [int(x) for x in rng.integers(0, 7, M + 1)]
            script = []
        g = refshim.inject(tetris, L, M, board, pcs)
        rows0[e] = pack_rows(board)
        npieces[e] = len(pcs)
        pieces[e, :len(pcs)] = pcs
        LM[e] = (L, M)
        t = 0
        extra = 2
        while g.pieces and t < T_MAX and extra >= 0:
            if len(g.pieces) >= 1 and rng.random() < 0.25:
                snapshots.append((e, t, L, M, g.board.copy(), list(g.pieces), int(g.lines_cleared), g.moves_used))
            if t < len(script):
                rot, loc = script[t]
            else:
                rot, loc = int(rng.integers(-2, 8)), int(rng.integers(0, 13))
            g.move(rot, loc)
            acts[e, t] = (rot, loc)
            a_rows[e, t] = pack_rows(g.board)
            a_meta[e, t] = (int(g.lines_cleared), g.moves_used, sc(g.state), len(g.pieces))
            t += 1
            if g.state is not None:
                extra -= 1
        nmoves[e] = t
    np.savez_compressed(os.path.join(HERE, "moves_random.npz"), rows0=rows0, pieces=pieces, npieces=npieces, LM=LM,
                        actions=acts, nmoves=nmoves, after_rows=a_rows, after_meta=a_meta)
    return snapshots


def make_afterstates(tetris, snapshots, S=320):
    rng = np.random.default_rng(11)
    idx = rng.permutation(len(snapshots))[:S]
    S = len(idx)
    rows = np.zeros((S, 20), np.uint16)
    pieces = np.zeros((S, P), np.uint8)
    npieces = np.zeros(S, np.uint8)
    meta = np.zeros((S, 4), np.int32)                    # L, M, lines_cleared, moves_used
    feats = np.zeros((S, 4, 10, 4), np.uint8)
    flags = np.zeros((S, 4, 10), np.uint8)               # 1 topout, 2 win, 4 lose-by-moves (alias bit added by tests)
    boards = np.zeros((S, 4, 10, 20), np.uint16)
    for s, i in enumerate(idx):
        e, t, L, M, board, pcs, lines, moves = snapshots[i]
        rows[s] = pack_rows(board)
        npieces[s] = len(pcs)
        pieces[s, :len(pcs)] = pcs
        meta[s] = (L, M, lines, moves)
        for r in range(4):
            for c in range(10):
                g = refshim.inject(tetris, L, M, board, pcs)
                g.lines_cleared, g.moves_used = lines, moves
                g.move(r, c)
                dl = int(g.lines_cleared) - lines
                fl = 0
                if g.moves_used == moves:
                    assert g.state is False
                    fl = 1
                elif g.state is True:
                    fl = 2
                elif g.state is False:
                    fl = 4
                holes, bump, agg = np_features(g.board)
                feats[s, r, c] = (dl, holes, bump, agg)
                flags[s, r, c] = fl
                boards[s, r, c] = pack_rows(g.board)
    np.savez_compressed(os.path.join(HERE, "afterstates.npz"), rows=rows, pieces=pieces, npieces=npieces, meta=meta,
                        feats=feats, flags=flags, boards=boards)


def main():
    tetris = refshim.load()
    make_kat(tetris)
    pool = make_pool(tetris)
    snaps = make_moves(tetris, pool)
    make_afterstates(tetris, snaps)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
