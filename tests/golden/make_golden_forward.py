"""Golden fixture for the FORWARD reset-point producer, written by the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_forward.py

For every (goal, tetrominoes, initial_height_max, max_attempts) setting and seeds 0..N-1 it records what
game/tetris_algo_main/main.py:generate_game / solve_game compute -- TetrisGameGenerator(seed, ...).board / .sequence and
TetrisSolver(board, sequence, goal, max_attempts).solve() -- plus, for one setting, the reset points that
game/tetris.py:translate makes of the winnable games (with random.seed(1234) before the call, because translate draws
the first piece of every reset point from the global stream).  -> forward_games.npz
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ['I', 'J', 'L', 'O', 'S', 'T', 'Z']
SETTINGS = [(10, 30, 4, 1000, 24), (3, 20, 4, 300, 60), (2, 12, 6, 200, 60), (15, 40, 7, 150, 12), (1, 7, 4, 50, 40)]


def pack_rows(board) -> np.ndarray:
    b = np.asarray(board).astype(bool)
    return (b.astype(np.uint16) << np.arange(10, dtype=np.uint16)).sum(axis=1).astype(np.uint16)


def main():
    tetris = refshim.load()                                     # puts game/ on sys.path
    from tetris_algo_main.TetrisGameGenerator import TetrisGameGenerator
    from tetris_algo_main.TetrisSolver import TetrisSolver
    out = {"settings": np.array(SETTINGS, np.int32)}
    for si, (goal, tet, ihm, max_attempts, nseeds) in enumerate(SETTINGS):
        rows = np.zeros((nseeds, 20), np.uint16)
        letters = np.zeros((nseeds, tet), np.uint8)
        solvable = np.zeros(nseeds, np.uint8)
        failed = np.zeros(nseeds, np.int32)
        moves = np.full((nseeds, tet, 3), -1, np.int8)
        nmoves = np.zeros(nseeds, np.uint8)
        games = []
        for seed in range(nseeds):
            g = TetrisGameGenerator(seed=seed, goal=goal, tetrominoes=tet, initial_height_max=ihm)
            rows[seed] = pack_rows(g.board)
            letters[seed] = [tetris.piece_translations[c] for c in g.sequence]
            ok, stack, fa = TetrisSolver(g.board, g.sequence, g.goal, max_attempts=max_attempts).solve()
            solvable[seed], failed[seed] = int(bool(ok)), int(fa)
            if ok:
                nmoves[seed] = len(stack)
                for k, (name, rot, col) in enumerate(stack):
                    moves[seed, k] = (NAMES.index(name), rot, col)
                games.append(g)
        print(f"setting {si} {(goal, tet, ihm, max_attempts)}: {int(solvable.sum())}/{nseeds} winnable")
        for k, v in dict(rows=rows, letters=letters, solvable=solvable, failed=failed, moves=moves, nmoves=nmoves).items():
            out[f"s{si}_{k}"] = v
        if si == 1:                                              # translate (game/tetris.py:19-20) on the winnable games
            random.seed(1234)
            pts = tetris.translate(games)
            out["translate_rows"] = np.array([pack_rows(b) for b, _ in pts], np.uint16)
            out["translate_pieces"] = np.array([p for _, p in pts], np.uint8)
            out["translate_tail"] = np.array([random.random() for _ in range(3)])
    np.savez_compressed(os.path.join(HERE, "forward_games.npz"), **out)


if __name__ == "__main__":
    main()
