"""Pin the oracles to the LIVE reference  --  TEST INFRASTRUCTURE, build container only.

Replays seeded random episodes through (1) the unmodified reference ``game/tetris.py`` imported
from /root/reference, (2) the Python bitrow restatement (piclim_oracle.py) and (3) the C
restatement (piclim_oracle.c), comparing after EVERY move: board (packed rows), remaining pieces,
lines_cleared, moves_used, state.  Boards are adversarial on purpose: heights 0-20,
densities 0.3-0.9, overhangs, cells in row 0, pre-existing full rows, rot in [-2, 7],
loc in [0, 12].

    python oracle/validate_against_reference.py --episodes 100000 --procs 8
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, piclim_oracle as po, refshim  # noqa: E402


def random_board(rng: np.random.Generator):
    H = int(rng.integers(0, 21))
    dens = float(rng.uniform(0.3, 0.9))
    b = np.zeros((20, 10), bool)
    if H:
        b[20 - H:] = rng.random((H, 10)) < dens
    kind = int(rng.integers(0, 6))
    if kind == 0 and H:                     # plant a pre-existing full row
        b[int(rng.integers(20 - H, 20))] = True
    elif kind == 1:                         # a cell in the very top row
        b[0, int(rng.integers(0, 10))] = True
    elif kind == 2 and H >= 3:              # knock a column out (deep well / overhang)
        b[20 - H + 1:, int(rng.integers(0, 10))] = False
    return b


def check_tables(tetris):
    for p, fam in enumerate(tetris.tetrominos):
        assert len(fam) == po.N_ROT[p]
        for r, (shape, prof) in enumerate(fam):
            masks = tuple(int(sum(1 << j for j in range(shape.shape[1]) if shape[i, j])) for i in range(shape.shape[0]))
            assert masks == po.ORIENT_ROWS[p][r], (p, r, masks)
            assert tuple(prof) == po.bottom_profile(masks), (p, r)


def run_chunk(args):
    seed, episodes = args
    tetris = refshim.load()
    rng = np.random.default_rng(seed)
    moves = topouts = wins = 0
    for ep in range(episodes):
        L = int(rng.integers(1, 16))
        M = int(rng.integers(1, 41))
        board = random_board(rng)
        npieces = M + 1
        pieces = [int(x) for x in rng.integers(0, 7, npieces)]
        ref = refshim.inject(tetris, L, M, board, pieces)
        rows = po.rows_from_bool(board)
        py = po.OracleEnv(L, M).load(rows, pieces)
        cst = c_oracle.BatchState(1).load(np.array([rows], np.uint16), np.array([pieces], np.uint8), npieces)
        # play until termination, then up to 2 more moves (nothing stops move() after termination)
        extra = 2
        while ref.pieces and extra >= 0:
            rot = int(rng.integers(-2, 8))
            loc = int(rng.integers(0, 13))
            before = ref.moves_used
            ref.move(rot, loc)
            py.move(rot, loc)
            c_oracle.step_batch(cst, [rot], [loc], L, M)
            moves += 1
            rrows = po.rows_from_bool(ref.board)
            sc = refshim.state_code(ref.state)
            assert rrows == py.rows, (seed, ep, "py rows")
            assert list(ref.pieces) == py.pieces
            assert int(ref.lines_cleared) == py.lines_cleared and ref.moves_used == py.moves_used
            assert sc == py.state
            assert rrows == [int(x) for x in cst.rows[0]], (seed, ep, "c rows")
            assert int(cst.head[0]) == npieces - len(ref.pieces)
            assert int(ref.lines_cleared) == int(cst.lines[0]) and ref.moves_used == int(cst.moves[0])
            assert sc == int(cst.state[0])
            if sc != 0:
                extra -= 1
                if ref.moves_used == before and sc == 2:
                    topouts += 1
                wins += sc == 1
    return moves, topouts, wins


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=100000)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--seed", type=int, default=2024)
    a = ap.parse_args()
    tetris = refshim.load()
    check_tables(tetris)
    c_oracle.build()
    per = (a.episodes + a.procs - 1) // a.procs
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        res = pool.map(run_chunk, [(a.seed + i, per) for i in range(a.procs)])
    moves = sum(r[0] for r in res)
    print(f"OK: {per * a.procs} episodes, {moves} moves, {sum(r[1] for r in res)} top-out terminations, "
          f"{sum(r[2] for r in res)} post-win moves; reference == python oracle == C oracle after every move "
          f"({time.time() - t0:.1f} s on {a.procs} procs)")


if __name__ == "__main__":
    main()
