"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against golden fixtures and the oracle."""
import os

import numpy as np
import pytest

from tests import parity_cases as pc

pytestmark = pytest.mark.gpu


def _pool(golden_dir):
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    pieces = np.zeros((len(z["rows"]), 42), np.uint8)
    pieces[:, :z["pieces"].shape[1]] = z["pieces"]
    return z["rows"], pieces, z["npieces"]


def test_library_loaded(gpu):
    """The native library is what runs: it must be mapped into this process."""
    with open("/proc/self/maps") as f:
        assert "libtetris_piclim_sm100.so" in f.read()


def test_golden_kat(gpu, golden_dir):
    pc.case_golden_kat(gpu, golden_dir)


def test_golden_moves(gpu, golden_dir):
    assert pc.case_golden_moves(gpu, golden_dir) > 2000


def test_golden_afterstates(gpu, golden_dir):
    assert pc.case_golden_afterstates(gpu, golden_dir) >= 300


@pytest.mark.parametrize("L,M,seed", [(10, 30, 1), (15, 40, 2), (1, 1, 3), (3, 41, 4)])
def test_random_moves_vs_oracle(gpu, L, M, seed):
    pc.case_random_moves(gpu, 100_000, min(M + 3, 34), L, M, seed)


@pytest.mark.parametrize("L,M,seed", [(10, 30, 5), (2, 5, 6), (15, 40, 7)])
def test_afterstates_vs_oracle(gpu, L, M, seed):
    pc.case_afterstates_vs_oracle(gpu, 100_000, L, M, seed)


def test_rng(gpu):
    pc.case_rng(gpu)


def test_reset(gpu, golden_dir):
    pc.case_reset(gpu, _pool(golden_dir))


def test_rollout_random_1e5_episodes(gpu, golden_dir):
    """>= 1e5 seeded episodes, fused rollout vs the oracle: final boards, queues, counters, episode numbers
    and the 8 statistics all bit-exact (north-star gate)."""
    stats = pc.case_rollout(gpu, _pool(golden_dir), 20_000, 150, 10, 30, seed=11, env_base=5_000_000_000, chunks=(0.2, 0.8))
    assert stats[0] >= 100_000, stats


def test_rollout_greedy(gpu, golden_dir):
    stats = pc.case_rollout(gpu, _pool(golden_dir), 6_000, 90, 10, 30, seed=12, env_base=77,
                            weights=[760, -360, -180, -510, 100000, -100000], chunks=(0.5, 0.5))
    assert stats[1] > 0


def test_fused_step_observe(gpu, golden_dir):
    pc.case_fused_step_observe(gpu, _pool(golden_dir))


def test_edges(gpu):
    pc.case_edges(gpu)


def test_sorted_afterstates_kernel_opt_in(gpu):
    """The opt-in piece-sorted kernel (TPL_SORTED_AFTERSTATES=1; counting sort in shared memory, warp-uniform alias
    skipping, TMA bulk stores) must produce the same bytes.  The switch is read once per process: run it in a child."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests.engines import GpuEngine\n"
            "from tests import parity_cases as pc\n"
            "g = GpuEngine()\n"
            "pc.case_afterstates_vs_oracle(g, 100_000, 10, 30, 5)\n"
            "pc.case_afterstates_vs_oracle(g, 4100, 15, 40, 9)\n"
            "print('sorted-ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TPL_SORTED_AFTERSTATES="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sorted-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_compact_form_ragged_sizes(gpu, golden_dir):
    """Partial last warp tiles (n % 32 != 0) and odd sizes through the fused kernel and the stand-alone one."""
    for n in (3001, 3002, 4096 + 132):
        pc.case_fused_step_observe(gpu, _pool(golden_dir), n=n, steps=12)
        pc.case_afterstates_vs_oracle(gpu, 40_000 + (n % 4), 10, 30, 3)


def test_fused_step_full_size_vs_oracle(gpu, golden_dir):
    """BASELINE configs[2] size: 2^20 envs through the fused kernel bench.py times, every packed feature word of all 40
    slots, every record and every counter against the C oracle (and against the three stand-alone kernels)."""
    pc.case_fused_step_observe(gpu, _pool(golden_dir), n=1 << 20, steps=12, seed=17, env_base=1 << 33)


def test_fused_step_is_deterministic(gpu, golden_dir):
    """Two runs of the same 80 fused steps (warp-private queues, persistent warps) give identical bytes every step,
    and equal the three stand-alone kernels at the end."""
    import torch
    import tetris_piclim as tp
    pool = tp.load_pool(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    n = 300_000
    envs = [tp.BatchedTetris(n, 10, 30, seed=3, config_pool=pool, env_base=77) for _ in range(3)]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    for t in range(80):
        rot = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        loc = torch.randint(0, 10, (n,), device="cuda", dtype=torch.uint8, generator=g)
        a = envs[0].step_observe(rot, loc, packed=True)
        b = envs[1].step_observe(rot, loc, packed=True)
        for x, y in zip(a, b):
            if x is not None:
                assert torch.equal(x, y), f"step {t}"
        envs[2].move(rot, loc); envs[2].reset(done_only=True)
    assert torch.equal(envs[0].state, envs[1].state) and torch.equal(envs[0].state, envs[2].state)
    assert torch.equal(envs[0].episode, envs[2].episode)
    assert torch.equal(a[3], envs[2].afterstates(packed=True, raw=True)[0])


def test_compact_form_64bit_pointer_chain(gpu, golden_dir):
    """The compact form advances its store pointer with a 32-bit add when the output array sits inside one 4 GB window
    (the usual case) and with a 64-bit chain otherwise; TPL_NO_P32=1 (read once per process: child) forces the latter."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests.engines import GpuEngine\n"
            "from tests import parity_cases as pc\n"
            "from tests.test_gpu_parity import _pool\n"
            "g = GpuEngine()\n"
            "pc.case_afterstates_vs_oracle(g, 100_000, 10, 30, 5)\n"
            "pc.case_fused_step_observe(g, _pool(%r), n=50_000, steps=20)\n"
            "print('chain64-ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), golden_dir)
    env = dict(os.environ, TPL_NO_P32="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "chain64-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
