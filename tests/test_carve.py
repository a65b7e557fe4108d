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
