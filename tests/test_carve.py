"""CPU tier: the native carving generator (csrc/carve_gen.cpp) against configs produced by the unmodified reference
(tests/golden/carve_pool_L10_M30.npz and kat_carve.npz: random.seed(s); Tetris(L, M, warm_reset=False, debug=True))."""
import ctypes
import os
import random

import numpy as np
import pytest

from oracle import piclim_oracle as po, refshim


@pytest.fixture(scope="module")
def tp():
    import tetris_piclim
    return tetris_piclim


def test_python_random_stream_is_reproduced(tp):
    lib = ctypes.CDLL(tp.build.build_carve())
    for seed, hi in [(0, 3), (0, 6), (1, 9), (1000, 7), (2 ** 40 + 5, 5), (123456789, 0)]:
        out = np.zeros(500, np.int32)
        lib.carve_pyrandom_randints(ctypes.c_uint64(seed), hi, 500, ctypes.c_void_p(out.ctypes.data))
        random.seed(seed)
        assert [random.randint(0, hi) for _ in range(500)] == out.tolist()


def test_carve_pool_equals_reference_pool(tp, golden_dir):
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    K = len(z["rows"])
    pool = tp.carve_pool(K, 10, 30, seed0=1000, threads=4)
    assert np.array_equal(pool.rows, z["rows"])
    assert np.array_equal(pool.npieces, z["npieces"]) and (pool.npieces == 31).all()
    assert np.array_equal(pool.pieces[:, :z["pieces"].shape[1]], z["pieces"])
    assert np.array_equal(pool.nsol, z["nsol"]) and np.array_equal(pool.solutions, z["solutions"])


def test_carve_kats_and_solutions_win(tp, golden_dir):
    z = np.load(os.path.join(golden_dir, "kat_carve.npz"))
    for i in range(int(z["count"])):
        seed, L, M = (int(v) for v in z[f"k{i}_meta"])
        pool = tp.carve_pool(1, L, M, seed0=seed, threads=1)
        assert np.array_equal(pool.rows[0], z[f"k{i}_rows"])
        assert np.array_equal(pool.pieces[0, :M + 1], z[f"k{i}_pieces"])
        sol = pool.solutions[0, :pool.nsol[0]]
        assert np.array_equal(sol, z[f"k{i}_solution"])
        e = po.OracleEnv(L, M).load(pool.rows[0], pool.pieces[0, :M + 1])
        for rot, loc in sol:
            e.move(int(rot), int(loc))
        assert e.state == po.WON                           # game/main.py:49-57
    with pytest.raises(ValueError):
        tp.carve_pool(1, 10, 42)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="live reference tree not present")
def test_carve_vs_live_reference(tp):
    tetris = refshim.load()
    for seed, L, M in [(7, 10, 30), (11, 8, 20), (4, 12, 35)]:
        random.seed(seed)
        g = tetris.Tetris(L, M, warm_reset=False, debug=True)
        pool = tp.carve_pool(1, L, M, seed0=seed, threads=1)
        assert po.rows_from_bool(g.board) == [int(x) for x in pool.rows[0]]
        assert g.pieces == [int(x) for x in pool.pieces[0, :pool.npieces[0]]]
        assert g.solution == [tuple(int(v) for v in s) for s in pool.solutions[0, :pool.nsol[0]]]


# ---- the drop-in facade on Python's GLOBAL random stream (how the reference's own drivers use it) ----------------

def test_facade_ctor_draws_the_reference_config_from_global_random(tp, golden_dir):
    """random.seed(s); Tetris(L, M, warm_reset=False, debug=True) -> the reference's board / pieces / solution
    (fixtures written by the unmodified reference); needs no GPU (the CUDA handle is created by the first move)."""
    z = np.load(os.path.join(golden_dir, "kat_carve.npz"))
    for i in range(int(z["count"])):
        seed, L, M = (int(v) for v in z[f"k{i}_meta"])
        random.seed(seed)
        g = tp.Tetris(L, M, warm_reset=False, debug=True)
        assert g.board.dtype == bool and g.board.shape == (20, 10)
        assert po.rows_from_bool(g.board) == [int(x) for x in z[f"k{i}_rows"]]
        assert g.pieces == [int(x) for x in z[f"k{i}_pieces"]]
        assert g.solution == [tuple(int(v) for v in s) for s in z[f"k{i}_solution"]]
        assert (g.lines_cleared, g.moves_used, g.state) == (0, 0, None)
        g.terminate()
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    for k in range(0, 24):
        random.seed(1000 + k)
        g = tp.Tetris(10, 30, warm_reset=False, debug=True)
        assert po.rows_from_bool(g.board) == [int(x) for x in z["rows"][k]]
        assert g.pieces == [int(x) for x in z["pieces"][k, :31]]


def test_reference_unit_tests_restated_against_the_facade(tp):
    """game/main.py:7-47 with `tetris` = this package's drop-in module (the move-based test is in the GPU tier)."""
    import importlib.util                                           # the repo-root `tetris.py` shim, loaded by path:
    spec = importlib.util.spec_from_file_location(                  # refshim may have put the reference's module of the
        "_dropin_tetris", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tetris.py"))
    tetris = importlib.util.module_from_spec(spec)                  # same name into sys.modules
    spec.loader.exec_module(tetris)
    assert tetris.Tetris is tp.Tetris
    random.seed(5)
    gen = tetris.RandomPieceGenerator()                             # game/main.py:7-18
    for i in range(16):
        (piece, index), regenerated = gen.get_random_piece()
        assert regenerated == (i % 7 == 0) and len(gen) == 7 - (i % 7)
        gen.delete_index(index)
        assert len(gen) == 7 - (i % 7) - 1
    gen = tetris.RandomPieceGenerator()                             # game/main.py:20-29
    seq = gen.get_random_sequence(16)
    assert len(seq) == 16
    for i in range(0, 16, 7):
        assert len(seq[i:i + 7]) == len(set(seq[i:i + 7]))
    L, M = 15, 40                                                   # game/main.py:32-47
    random.seed(3)
    game = tetris.Tetris(L, M, warm_reset=False, debug=True)
    other = tetris.Tetris(L, M, warm_reset=False, debug=True)
    other.board[:, :] = False
    other.board[-L:, :] = True
    for i in range(len(game.solution) - 1, -1, -1):
        rotations, location = game.solution[i]
        assert other.carve(game.pieces[i], rotations, location, i == len(game.solution) - 1)
    game.terminate()
    assert np.array_equal(game.board, other.board)
    t = tetris.get_tetromino(1, 5)
    assert t[0].shape == (3, 2) and tuple(t[1]) == (0, 2)           # rot % n_rot, game/tetris.py:60-61


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="live reference tree not present")
def test_facade_leaves_global_random_where_the_reference_does(tp):
    tetris = refshim.load()
    for seed, L, M in [(0, 10, 30), (9, 8, 20), (2, 15, 40)]:
        random.seed(seed)
        r = tetris.Tetris(L, M, warm_reset=False, debug=True)
        r.reset()                                                   # second config from the continued stream
        tail_ref = [random.random() for _ in range(4)]
        random.seed(seed)
        g = tp.Tetris(L, M, warm_reset=False, debug=True)
        g.reset()
        assert [random.random() for _ in range(4)] == tail_ref
        assert np.array_equal(g.board, r.board) and g.pieces == r.pieces
        gen_r, gen_g = tetris.RandomPieceGenerator(), tp.RandomPieceGenerator()
        random.seed(seed); a = gen_r.get_random_sequence(23)
        random.seed(seed); b = gen_g.get_random_sequence(23)
        assert a == b
        d = g.calculate_drop_deltas(2, (0, 1, 1), 3)
        assert list(d) == list(r.calculate_drop_deltas(2, (0, 1, 1), 3)) and g.calculate_drop(d) == r.calculate_drop(d)
