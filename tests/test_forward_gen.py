"""CPU tier: the native forward generator + solver (csrc/forward_gen.cpp) against what the unmodified reference computes
(tests/golden/forward_games.npz, written by tests/golden/make_golden_forward.py from game/tetris_algo_main)."""
import os
import random

import numpy as np
import pytest

from oracle import refshim


@pytest.fixture(scope="module")
def tp():
    import tetris_piclim
    return tetris_piclim


def test_forward_games_equal_the_reference(tp, golden_dir):
    z = np.load(os.path.join(golden_dir, "forward_games.npz"))
    winnable = 0
    for si, (goal, tet, ihm, max_attempts, nseeds) in enumerate(z["settings"]):
        g = tp.forward_games(int(goal), int(tet), 0, int(nseeds), int(ihm), int(max_attempts), threads=4)
        for k in ("rows", "letters", "solvable", "failed", "nmoves", "moves"):
            assert np.array_equal(g[k], z[f"s{si}_{k}"]), (si, k)
        # every aligned group of 7 is a permutation of the 7 pieces (the 7-bag contract, game/main.py:20-29)
        for seq in g["letters"]:
            for a in range(0, len(seq) - 6, 7):
                assert sorted(seq[a:a + 7]) == list(range(7))
        winnable += int(g["solvable"].sum())
    assert winnable > 100


def test_forward_pool_is_translate_of_generate_batch(tp, golden_dir):
    """game/tetris.py:19-20 on the winnable games of setting 1, incl. the position of the global random stream."""
    z = np.load(os.path.join(golden_dir, "forward_games.npz"))
    goal, tet, ihm, max_attempts, nseeds = (int(v) for v in z["settings"][1])
    random.seed(1234)
    pool = tp.forward_pool(goal, tet, 0, nseeds, ihm, max_attempts, threads=2)
    assert np.array_equal(pool.rows, z["translate_rows"])
    assert np.array_equal(pool.pieces[:, :tet + 1], z["translate_pieces"]) and (pool.npieces == tet + 1).all()
    assert [random.random() for _ in range(3)] == z["translate_tail"].tolist()
    with pytest.raises(ValueError):
        tp.forward_pool(10, 42)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="live reference tree not present")
def test_forward_vs_live_reference(tp):
    refshim.load()
    from tetris_algo_main.TetrisGameGenerator import TetrisGameGenerator
    from tetris_algo_main.TetrisSolver import TetrisSolver
    names = ['I', 'J', 'L', 'O', 'S', 'T', 'Z']
    for goal, tet, ihm, ma, seed0, cnt in [(2, 15, 5, 120, 1000, 6), (4, 25, 8, 80, 77, 4)]:
        g = tp.forward_games(goal, tet, seed0, cnt, ihm, ma, threads=1)
        for k in range(cnt):
            r = TetrisGameGenerator(seed=seed0 + k, goal=goal, tetrominoes=tet, initial_height_max=ihm)
            assert np.array_equal(tp.configs.rows_from_bool(r.board.astype(bool)), g["rows"][k])
            ok, stack, fa = TetrisSolver(r.board, r.sequence, goal, max_attempts=ma).solve()
            assert (bool(ok), int(fa)) == (bool(g["solvable"][k]), int(g["failed"][k]))
            if ok:
                assert [(names.index(a), b, c) for a, b, c in stack] == [tuple(int(v) for v in m) for m in g["moves"][k, :g["nmoves"][k]]]
