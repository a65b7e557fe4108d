"""Drop-in module: ``import tetris`` (as the reference's drivers do, ``game/main.py:2``, ``game/performance_test.py:1``)
resolves to the B200-backed facade when the repo root is first on ``sys.path``."""
import tetris_piclim as _tp

Tetris = _tp.Tetris
RandomPieceGenerator = _tp.RandomPieceGenerator
tetrominos = _tp.tetrominos
get_tetromino = _tp.get_tetromino
