"""Importable alias of the package directory (whose name, fixed by the build contract, is not a Python
identifier): ``import tetris_piclim as tp; tp.BatchedTetris(...)``."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE_NAME = "reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
