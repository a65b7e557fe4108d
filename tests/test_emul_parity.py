"""CPU tier: the per-env device code (compiled for the host by tests/emul) against golden fixtures + oracle."""
import os

import numpy as np
import pytest

from tests import parity_cases as pc


def _pool(golden_dir):
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    pieces = np.zeros((len(z["rows"]), 42), np.uint8)
    pieces[:, :z["pieces"].shape[1]] = z["pieces"]
    return z["rows"], pieces, z["npieces"]


def test_golden_kat(emul, golden_dir):
    pc.case_golden_kat(emul, golden_dir)


def test_golden_moves(emul, golden_dir):
    assert pc.case_golden_moves(emul, golden_dir) > 2000


def test_golden_afterstates(emul, golden_dir):
    assert pc.case_golden_afterstates(emul, golden_dir) >= 300


@pytest.mark.parametrize("L,M,seed", [(10, 30, 1), (15, 40, 2), (1, 1, 3), (3, 41, 4)])
def test_random_moves_vs_oracle(emul, L, M, seed):
    pc.case_random_moves(emul, 3000, min(M + 3, 24), L, M, seed)


@pytest.mark.parametrize("L,M,seed", [(10, 30, 5), (2, 5, 6)])
def test_afterstates_vs_oracle(emul, L, M, seed):
    pc.case_afterstates_vs_oracle(emul, 4000, L, M, seed)


def test_rng(emul):
    pc.case_rng(emul)


def test_reset(emul, golden_dir):
    pc.case_reset(emul, _pool(golden_dir))


def test_rollout_random(emul, golden_dir):
    stats = pc.case_rollout(emul, _pool(golden_dir), 2000, 120, 10, 30, seed=11, env_base=5_000_000_000, chunks=(0.25, 0.75))
    assert stats[0] > 0


def test_rollout_greedy(emul, golden_dir):
    stats = pc.case_rollout(emul, _pool(golden_dir), 600, 80, 10, 30, seed=12, env_base=77,
                            weights=[760, -360, -180, -510, 100000, -100000], chunks=(0.5, 0.5))
    assert stats[1] > 0          # a sensible greedy policy wins some carve configs


def test_fused_step_observe(emul, golden_dir):
    pc.case_fused_step_observe(emul, _pool(golden_dir))


def test_edges(emul):
    pc.case_edges(emul)


def test_long_episodes_queue_refill(emul):
    pc.case_long_episodes(emul)


def test_expand_distinct_without_any_placement():
    """distinct.expand on a batch in which no env has a piece left: `rows` is empty, every slot reads TPL_FLAG_NOPIECE
    (found by scripts/fuzz_gpu.py with M = 1: the host helper indexed an empty array)."""
    import importlib
    dm = importlib.import_module("tetris_piclim").distinct if hasattr(importlib.import_module("tetris_piclim"), "distinct") \
        else importlib.import_module("reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.distinct")
    runs = np.full(5, np.uint32(7) << np.uint32(29), np.uint32)          # piece id 7 = no piece, offset 0
    out = dm.expand(np.zeros(0, np.uint32), runs)
    assert out.shape == (5, 40, 4) and (out[:, :, 0] == (16 << 3) & 0xFF).all() and not out[:, :, 1:].any()
