"""CPU tier: the C-ABI library loads without a GPU and exports exactly what include/tetris_piclim.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tp():
    import tetris_piclim
    return tetris_piclim


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tetris_piclim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tpl_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(tp):
    lib = ctypes.CDLL(tp._lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tetris_piclim.h but not exported"
    assert sorted(tp._lib.ALL_SYMBOLS) == syms, "python binding table and header disagree"
    assert lib.tpl_abi_version() == 2


def test_config_generator_library_exports_its_header(tp):
    src = open(os.path.join(ROOT, "include", "piclim_configs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b((?:carve|forward)_[a-z_0-9]+)\s*\(", src)))
    assert syms == ["carve_apply", "carve_generate", "carve_generate_from_state", "carve_pyrandom_randints", "forward_generate"]
    lib = ctypes.CDLL(tp.build.build_carve())
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/piclim_configs.h but not exported"
    rows = np.zeros(20, np.uint16)
    assert lib.carve_generate(ctypes.c_uint64(0), 1, 0, 30, ctypes.c_void_p(rows.ctypes.data), None, 31, None, None, None, 1) == -1
    assert lib.forward_generate(ctypes.c_uint64(0), 1, 3, 0, 4, 10, ctypes.c_void_p(rows.ctypes.data), None, None, None, None, None, 1) == -1


def test_argument_errors_need_no_gpu(tp):
    L = tp._lib.lib()
    assert L.tpl_step(None, 0, 4, None, None, None, None, None, None, 1, 1, None) == -1
    assert b"null" in L.tpl_last_error()
    assert L.tpl_gen_pieces(ctypes.c_void_p(8), 4, 42 * 256 + 1, 0, 0, None, 0, None) == -2      # sequences of any length up to 42 * 256
    assert L.tpl_afterstates(ctypes.c_void_p(8), 2, 4, ctypes.c_void_p(8), None, None, 1, 1, None) == -2   # stride < n
    assert L.tpl_pack(ctypes.c_void_p(8), 0, 0, 0, ctypes.c_void_p(8), ctypes.c_void_p(8), 1, ctypes.c_void_p(8),
                      None, None, None, None, None) == 0                                                      # n == 0 is a no-op
    with pytest.raises(tp.TplError):
        tp._lib.check(-1, "x")


def test_no_cpu_fallback(tp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        tp.BatchedTetris(8, 10, 30, device="cpu")
    with pytest.raises(Exception):
        tp.BatchedTetris(8, 10, 30)                     # no CUDA device: must fail loudly, not fall back
    with pytest.raises(tp.TplError):
        tp.HostBatchedTetris(8, 10, 30)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if f.endswith(".py"):
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                    assert "piclim_oracle" not in txt and "c_oracle" not in txt, f
                else:
                    assert "oracle/" not in txt and "piclim_oracle" not in txt, f


def test_tetromino_table_matches_oracle(tp):
    from oracle import piclim_oracle as po
    for p, fam in enumerate(tp.tetrominos):
        assert len(fam) == po.N_ROT[p]
        for r, (shape, prof) in enumerate(fam):
            masks = tuple(int(sum(1 << j for j in range(shape.shape[1]) if shape[i, j])) for i in range(shape.shape[0]))
            assert masks == po.ORIENT_ROWS[p][r] and prof == po.bottom_profile(masks)
    assert tp.get_tetromino(1, 5)[1] == tp.get_tetromino(1, 1)[1] and tp.get_tetromino(0, -1)[1] == (3,)


def test_synthetic_pool_is_deterministic_and_well_formed(tp):
    a, b = tp.synthetic_pool(512, seed=3, M=30), tp.synthetic_pool(512, seed=3, M=30)
    assert np.array_equal(a.rows, b.rows) and np.array_equal(a.pieces, b.pieces)
    c = tp.synthetic_pool(512, seed=4, M=30)
    assert not np.array_equal(a.rows, c.rows)
    rows = a.rows
    assert rows.shape == (512, 20) and rows.max() < 0x3FF                      # no full rows
    h = (rows != 0).sum(axis=1)
    assert h.max() <= 12 and h.min() == 0 and len(np.unique(h)) == 13           # H ~ U{0..12}
    for k in range(512):                                                        # filled rows are contiguous from the floor
        assert (rows[k, 20 - h[k]:] != 0).all() and (rows[k, :20 - h[k]] == 0).all()
    dens = np.unpackbits(rows[rows != 0].view(np.uint8)).sum() / (10 * (rows != 0).sum())
    assert 0.55 < dens < 0.66
    assert (a.npieces == 31).all()
    for g in range(0, 28, 7):
        assert (np.sort(a.pieces[:, g:g + 7], axis=1) == np.arange(7)).all()     # 7-bag
    from oracle import piclim_oracle as po
    assert list(a.pieces[5, :31]) == po.gen_pieces(3, 5, 0xB0A2D, 31)
    with pytest.raises(ValueError):
        tp.synthetic_pool(4, M=42)
