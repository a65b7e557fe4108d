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
    env = dict(os.environ, TPL_NO_P32="1", TPL_NO_PDL="1")      # (and plain launches instead of programmatic dependent launch)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "chain64-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("log2n", [22, 23])
def test_big_batches_vs_oracle(gpu, golden_dir, log2n):
    """2^22 and 2^23 envs (BASELINE configs[4]: 8 M envs over 2 / 4 GPUs = 4 M / 2 M per GPU) against the C ORACLE, not against the
    stand-alone kernels: fused random rollout (boards, queues, counters, episode numbers, statistics), then two fused steps with
    external actions (compact 40-slot form, then the distinct-placements form) incl. the auto-reset and every afterstate word."""
    import torch
    import tetris_piclim as tp
    from oracle import c_oracle
    n, L, M, seed, base = 1 << log2n, 10, 30, 77, 1 << 36
    prow, ppieces, pnp = _pool(golden_dir)
    pool = tp.ConfigPool(prow, ppieces, pnp)
    env = tp.BatchedTetris(n, L, M, seed=seed, config_pool=pool, env_base=base)
    env.reset()
    ost = c_oracle.BatchState(n)
    oep, ots, ostats = c_oracle.rollout(ost, base, seed, L, M, prow, ppieces, pnp, 0, True)
    env.rollout_random(9)
    _, _, s1 = c_oracle.rollout(ost, base, seed, L, M, prow, ppieces, pnp, 9, False, oep, ots, nthreads=16)
    assert np.array_equal(env.stats.cpu().numpy(), s1)

    def same_state(what):
        f = env.fields()
        for k, o in (("rows", ost.rows), ("state", ost.state), ("lines", ost.lines), ("moves", ost.moves), ("head", ost.head)):
            assert np.array_equal(f[k].cpu().numpy(), o), f"{what}: {k}"
        assert np.array_equal(env.episode.cpu().numpy().view(np.uint32), oep), what
    same_state("after the fused rollout")
    rng = np.random.default_rng(log2n)
    for t, distinct in enumerate((False, True)):
        rot, loc = rng.integers(0, 4, n).astype(np.uint8), rng.integers(0, 10, n).astype(np.uint8)
        d_rot, d_loc = torch.from_numpy(rot).cuda(), torch.from_numpy(loc).cuda()
        if distinct:
            dl, fl, st, rows, runs, used = env.step_observe_distinct(d_rot, d_loc)
        else:
            dl, fl, st, feats, _, _ = env.step_observe(d_rot, d_loc, packed=True)
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(dl.cpu().numpy(), odl) and np.array_equal(st.cpu().numpy(), ost.state)
        c_oracle.reset_done(ost, base, seed, prow, ppieces, pnp, oep, ots)
        same_state(f"fused step {t}")
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=16)
        of[:, :, 0] |= ofl << 3
        del ofl
        if distinct:
            got = env.expand_distinct(rows, runs).permute(1, 0, 2).contiguous().cpu().numpy()
            assert int(used) < 0.62 * 40 * n
        else:
            got = feats.permute(1, 0, 2).contiguous().cpu().numpy()
        assert np.array_equal(got, of), f"afterstates of fused step {t}"
        del got, of


def test_long_episodes_queue_refill(gpu):
    pc.case_long_episodes(gpu, n=5000)
