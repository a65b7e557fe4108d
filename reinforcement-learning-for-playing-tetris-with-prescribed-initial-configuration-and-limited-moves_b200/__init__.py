"""B200-native batched Tetris-piclim simulator: the rollout hot path of the reference's ``game/tetris.py``
behind its own ``Tetris`` interface, on hand-written sm_100a CUDA kernels reached through a C ABI
(``include/tetris_piclim.h``).  Importing the env classes loads ``csrc/libtetris_piclim_sm100.so``; there is no
CPU fallback."""
from .configs import (ConfigPool, carve_pool, concat_pools, forward_games, forward_pool, gen_pieces, load_pool,  # noqa: F401
                      save_pool, synthetic_pool)
from . import build  # noqa: F401
from ._lib import TplError, launch_count  # noqa: F401
from .host_env import HostBatchedTetris, PinnedArray  # noqa: F401
from .tetris import RandomPieceGenerator, Tetris, get_tetromino, tetrominos  # noqa: F401


def __getattr__(name):          # torch is only imported when the device-tensor API is asked for
    if name == "BatchedTetris":
        from .batched import BatchedTetris
        return BatchedTetris
    if name in ("ValueNet", "Model"):
        from . import model
        return getattr(model, name)
    if name == "ValueKernel":
        from .value_kernel import ValueKernel
        return ValueKernel
    raise AttributeError(name)
