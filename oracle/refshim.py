"""Import shim for the *live* reference  --  TEST INFRASTRUCTURE ONLY.

Only usable where ``/root/reference`` exists (the build container).  Nothing that runs on
the GPU box (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this; they use the
committed fixtures under ``tests/golden/`` instead.

Recipe from SURVEY.md section 8c: ``game/tetris.py:6`` does ``from tetris_algo_main import main``,
so ``game/`` itself must be on ``sys.path``; the tree is read-only, so no bytecode may be written;
``warm_reset=False`` always (the default forks two producer processes, ``game/tetris.py:190-211``).
"""
from __future__ import annotations

import os
import sys

REFERENCE_GAME_DIR = "/root/reference/game"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_GAME_DIR, "tetris.py"))


def load():
    """Return the reference's ``tetris`` module (unmodified, imported in place)."""
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    sys.dont_write_bytecode = True
    if REFERENCE_GAME_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_GAME_DIR)
    import tetris  # noqa: E402  (the reference module)
    return tetris


def inject(tetris_mod, L, M, board_bool, pieces):
    """Build a reference ``Tetris`` holding a prescribed reset point without running the slow
    carve generator: set exactly the fields ``move``/``get_state`` touch
    (``game/tetris.py:143-151,186-187``) -- the same injection ``load_warm_reset`` does at ``:447``."""
    import numpy as np
    g = tetris_mod.Tetris.__new__(tetris_mod.Tetris)
    g.L, g.M = L, M
    g.warm_reset = False
    g.render = False
    g.debug = False
    g.lines_cleared = 0
    g.moves_used = 0
    g.state = None
    g.board = np.array(board_bool, dtype=bool).copy()
    g.pieces = [int(p) for p in pieces]
    return g


def state_code(state) -> int:
    """None/True/False -> 0/1/2 (running/won/lost)."""
    if state is None:
        return 0
    return 1 if state is True or (state is not False and bool(state)) else 2
