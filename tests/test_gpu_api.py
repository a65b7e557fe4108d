"""GPU tier: the public Python API (BatchedTetris, HostBatchedTetris, the Tetris facade) and full-size
size-independent properties."""
import os

import numpy as np
import pytest

from oracle import c_oracle, piclim_oracle as po
from tests import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tp(gpu):
    import tetris_piclim
    return tetris_piclim


@pytest.fixture(scope="module")
def carve_pool(tp, golden_dir):
    return tp.load_pool(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))


def test_facade_replays_reference_solutions(tp, carve_pool):
    """game/main.py:49-57 (test_carving_invertability) against the drop-in facade."""
    g = tp.Tetris(10, 30, warm_reset=False, debug=True, config_pool=carve_pool, seed=4)
    for _ in range(5):
        assert len(g.pieces) == 31 and g.solution
        for rot, loc in g.solution:
            g.move(rot, loc)
        assert g.state is True and g.lines_cleared >= 10
        g.reset(fresh=True)
        assert g.state is None and g.moves_used == 0
    g.terminate()


def test_reference_invertability_test_on_global_random(tp):
    """game/main.py:49-57 verbatim in spirit: seed Python's global random, build the env with the reference's ctor
    arguments (no config_pool: the native carve generator draws the reference's config), play the recorded solution."""
    import random
    for seed, (L, M) in enumerate([(15, 40), (10, 30), (15, 40)]):
        random.seed(seed)
        game = tp.Tetris(L, M, warm_reset=False, debug=True)
        assert len(game.pieces) == M + 1
        for rotations, location in game.solution:
            game.move(rotations, location)
        assert game.state is True
        game.terminate()


def test_facade_matches_oracle_move_by_move(tp, carve_pool):
    rng = np.random.default_rng(0)
    for ep in range(6):
        g = tp.Tetris(6, 12, warm_reset=False, config_pool=carve_pool, seed=ep)
        o = po.OracleEnv(6, 12).load(pc.po.rows_from_bool(g.board), g.pieces)
        b0 = g.board
        while g.pieces:
            rot, loc = int(rng.integers(-2, 8)), int(rng.integers(0, 13))
            g.move(rot, loc); o.move(rot, loc)
            assert po.rows_from_bool(g.board) == o.rows and g.pieces == o.pieces
            assert (g.lines_cleared, g.moves_used) == (o.lines_cleared, o.moves_used)
            assert g.state is {0: None, 1: True, 2: False}[o.state]
            if len(g.pieces) >= 2:
                st = g.get_state()
                assert st[0] is g.board and st[1:3] == (o.pieces[0], o.pieces[1])
                assert st[3] == 6 - o.lines_cleared and st[4] == 12 - o.moves_used
        with pytest.raises(IndexError):
            g.move(0, 0)
        g.terminate()
    g = tp.Tetris(6, 12, warm_reset=False, config_pool=carve_pool)
    with pytest.raises(ValueError):
        g.move(0, -1)
    # reference quirk (game/tetris.py:438-443): reset() keeps the counters; attributes are writable
    g.move(0, 0)
    g.reset()
    assert g.moves_used == 1
    g.board[:, :] = False
    g.board[-1, 4:] = True
    g.pieces = [0, 6, 6]
    g.L, g.lines_cleared, g.moves_used, g.state = 1, 0, 0, None
    g.move(0, 0)
    assert g.state is True and not g.board.any()
    ns = tp.Tetris(6, 12, warm_reset=False, config_pool=carve_pool).next_states()
    assert 9 <= len(ns) <= 34
    g.terminate()


def test_batched_api_vs_oracle(tp, carve_pool):
    import torch
    n, L, M = 5000, 10, 30
    env = tp.BatchedTetris(n, L, M, seed=21, config_pool=carve_pool, env_base=1000)
    env.reset()
    ost = c_oracle.BatchState(n)
    c_oracle.rollout(ost, 1000, 21, L, M, carve_pool.rows, carve_pool.pieces, carve_pool.npieces, 0, True)
    rng = np.random.default_rng(1)
    for t in range(12):
        feats, flags = env.afterstates()
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
        assert np.array_equal(feats.cpu().numpy().reshape(n, 40, 4), of)
        assert np.array_equal(flags.cpu().numpy().reshape(n, 40), ofl)
        rot, loc = rng.integers(-3, 9, n), rng.integers(0, 14, n)
        dl, fl, st = env.move(torch.as_tensor(rot), torch.as_tensor(loc)) if t % 2 else env.move(rot, loc)
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(dl.cpu().numpy(), odl) and np.array_equal(st.cpu().numpy(), ost.state)
        boards, cur, nxt, rem_l, rem_m, state = env.get_state()
        assert np.array_equal(boards.cpu().numpy(), ost.rows)
        assert np.array_equal(rem_l.cpu().numpy(), L - ost.lines) and np.array_equal(rem_m.cpu().numpy(), M - ost.moves)
        has2 = ost.head + 1 < ost.npieces
        assert np.array_equal(nxt.cpu().numpy()[has2], ost.pieces[np.arange(n), np.minimum(ost.head + 1, 41)][has2])
    bb = env.get_state(bool_boards=True)[0].cpu().numpy()
    assert np.array_equal(tp.configs.rows_from_bool(bb), ost.rows)
    with pytest.raises(ValueError):
        env.move(np.zeros(n), np.full(n, -1))
    with pytest.raises(ValueError):
        env.move(np.zeros(n + 1), np.zeros(n + 1))
    # counter-based 7-bag sequences through the batched API
    from oracle import c_oracle as co
    assert np.array_equal(env.gen_pieces(31, episode=4).cpu().numpy(), co.gen_pieces(21, 1000, n, 4, 31))
    # explicit boards reset + done_only auto-reset
    env.reset(done_only=True)
    f = env.fields()
    assert not f["state"].any()
    env.terminate()


def test_host_api_vs_oracle(tp, carve_pool):
    n, L, M = 3000, 10, 30
    env = tp.HostBatchedTetris(n, L, M, seed=31, env_base=9, config_pool=carve_pool)
    env.reset()
    ost = c_oracle.BatchState(n)
    oep, ots, _ = c_oracle.rollout(ost, 9, 31, L, M, carve_pool.rows, carve_pool.pieces, carve_pool.npieces, 0, True)
    rng = np.random.default_rng(2)
    pin = {k: tp.PinnedArray(s, d) for k, (s, d) in dict(rot=((n,), np.uint8), loc=((n,), np.uint8), dl=((n,), np.int8),
           fl=((n,), np.uint8), st=((n,), np.int8), feats=((40, n, 4), np.uint8), afl=((40, n), np.uint8)).items()}
    for t in range(40):
        rot, loc = rng.integers(0, 4, n), rng.integers(0, 10, n)
        pin["rot"].array[:] = rot; pin["loc"].array[:] = loc
        env.step_observe(*[pin[k].array for k in ("rot", "loc", "dl", "fl", "st", "feats", "afl")])
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(pin["dl"].array, odl) and np.array_equal(pin["st"].array, ost.state)
        # auto-reset of finished envs, episode e+1 drawn by the counter RNG
        done = np.where((ost.state != 0) | (ost.head >= ost.npieces))[0]
        for i in done:
            oep[i] += 1
            k = po.config_index(31, 9 + int(i), int(oep[i]), carve_pool.K)
            ost.rows[i] = carve_pool.rows[k]; ost.pieces[i] = carve_pool.pieces[k]; ost.npieces[i] = carve_pool.npieces[k]
            ost.head[i] = 0; ost.lines[i] = 0; ost.moves[i] = 0; ost.state[i] = 0
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
        assert np.array_equal(pin["feats"].array.transpose(1, 0, 2), of)
        assert np.array_equal(pin["afl"].array.T, ofl)
    f = env.fields()
    assert np.array_equal(f["rows"], ost.rows) and np.array_equal(f["moves"], ost.moves)
    # feats=None: the features stay on the device in the compact form; only the move's results come back
    import torch
    rot, loc = rng.integers(0, 4, n), rng.integers(0, 10, n)
    pin["rot"].array[:] = rot; pin["loc"].array[:] = loc
    env.step_observe(pin["rot"].array, pin["loc"].array, pin["dl"].array, pin["fl"].array, pin["st"].array, None, None)
    odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
    assert np.array_equal(pin["dl"].array, odl) and np.array_equal(pin["st"].array, ost.state)
    for i in np.where((ost.state != 0) | (ost.head >= ost.npieces))[0]:
        oep[i] += 1
        k = po.config_index(31, 9 + int(i), int(oep[i]), carve_pool.K)
        ost.rows[i] = carve_pool.rows[k]; ost.pieces[i] = carve_pool.pieces[k]; ost.npieces[i] = carve_pool.npieces[k]
        ost.head[i] = 0; ost.lines[i] = 0; ost.moves[i] = 0; ost.state[i] = 0
    of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
    exp = of.copy(); exp[:, :, 0] |= (ofl << 3)
    assert env.feats_device_ptr() != 0

    class _DevView:                                             # zero-copy view of the library's device buffer for torch
        __cuda_array_interface__ = {"shape": (40 * n * 4,), "typestr": "|u1", "data": (env.feats_device_ptr(), False), "version": 2}
    dev = torch.as_tensor(_DevView(), device="cuda")
    assert np.array_equal(dev.cpu().numpy().reshape(40, n, 4).transpose(1, 0, 2), exp)
    with pytest.raises(tp.TplError):
        env.step_observe(pin["rot"].array, pin["loc"].array, pin["dl"].array, pin["fl"].array, pin["st"].array, None, pin["afl"].array)
    dl, fl, st = env.move(np.zeros(n), np.zeros(n))
    feats, flags = env.afterstates()
    assert feats.shape == (n, 4, 10, 4) and flags.shape == (n, 4, 10)
    env.close()


def _oracle_autoreset(ost, oep, seed, env_base, pool):
    for i in np.where((ost.state != 0) | (ost.head >= ost.npieces))[0]:
        oep[i] += 1
        k = po.config_index(seed, env_base + int(i), int(oep[i]), pool.K)
        ost.rows[i] = pool.rows[k]; ost.pieces[i] = 0; ost.pieces[i, :pool.pieces.shape[1]] = pool.pieces[k]
        ost.npieces[i] = pool.npieces[k]
        ost.head[i] = 0; ost.lines[i] = 0; ost.moves[i] = 0; ost.state[i] = 0


@pytest.mark.parametrize("n,chunks", [(3000, None), (600_000, None), (70_001, "3")])
def test_host_api_pipelined_and_distinct(tp, carve_pool, n, chunks, monkeypatch):
    """tpl_env_step_observe (40-slot forms) and tpl_env_step_observe_distinct through HOST buffers, with the batch cut into
    chunks (600 000 envs -> 2 chunks; 70 001 envs forced into 3 ragged chunks): every chunk's results land at the right
    place of the caller's arrays, and the distinct form expands to the oracle's grid."""
    from tests import parity_cases as pc
    if chunks:
        monkeypatch.setenv("TPL_ENV_CHUNKS", chunks)
    L, M, seed, base = 10, 30, 41, 1 << 35
    envs = [tp.HostBatchedTetris(n, L, M, seed=seed, env_base=base, config_pool=carve_pool) for _ in range(2)]
    assert envs[0].chunks() == (int(chunks) if chunks else min(2, max(1, n // 262144)))
    for e in envs:
        e.reset()
    ost = c_oracle.BatchState(n)
    oep, _, _ = c_oracle.rollout(ost, base, seed, L, M, carve_pool.rows, carve_pool.pieces, carve_pool.npieces, 0, True)
    rng = np.random.default_rng(3)
    cap = envs[1].distinct_capacity()
    pin = {k: tp.PinnedArray(s, d) for k, (s, d) in dict(rot=((n,), np.uint8), loc=((n,), np.uint8), dl=((n,), np.int8),
           fl=((n,), np.uint8), st=((n,), np.int8), feats=((40, n, 4), np.uint8), afl=((40, n), np.uint8),
           dl2=((n,), np.int8), fl2=((n,), np.uint8), st2=((n,), np.int8), rows=((cap,), np.uint32), runs=((n,), np.uint32)).items()}
    for t in range(6):
        rot, loc = rng.integers(0, 4, n), rng.integers(0, 10, n)
        pin["rot"].array[:] = rot; pin["loc"].array[:] = loc
        packed = t % 2 == 0
        envs[0].step_observe(*[pin[k].array for k in ("rot", "loc", "dl", "fl", "st", "feats")], None if packed else pin["afl"].array)
        pin["rows"].array[:] = 0xFFFFFFFF
        words = envs[1].step_observe_distinct(*[pin[k].array for k in ("rot", "loc", "dl2", "fl2", "st2", "rows", "runs")])
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        for a, b in (("dl", "dl2"), ("fl", "fl2"), ("st", "st2")):
            assert np.array_equal(pin[a].array, pin[b].array)
        assert np.array_equal(pin["dl"].array, odl) and np.array_equal(pin["st"].array, ost.state)
        _oracle_autoreset(ost, oep, seed, base, carve_pool)
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
        grid = of.copy(); grid[:, :, 0] |= ofl << 3
        if packed:
            assert np.array_equal(pin["feats"].array.transpose(1, 0, 2), grid)
        else:
            assert np.array_equal(pin["feats"].array.transpose(1, 0, 2), of) and np.array_equal(pin["afl"].array.T, ofl)
        dm = pc.distinct_module()
        runs = pin["runs"].array
        assert np.array_equal(dm.expand(pin["rows"].array, runs), grid)
        cnt = dm.tables()[0][runs >> 29].astype(np.int64)
        assert cnt.sum() <= words <= cnt.sum() + 3 * ((n + 31) // 32 + 8)          # only the used words crossed PCIe
        assert words < 0.7 * 40 * n
    for e in envs:
        f = e.fields()
        assert np.array_equal(f["rows"], ost.rows) and np.array_equal(f["moves"], ost.moves)
        e.close()


def test_batched_distinct_forms(tp, carve_pool):
    """BatchedTetris.afterstates_distinct / step_observe_distinct / expand_distinct and the torch gather helper"""
    import torch
    from tests import parity_cases as pc
    n, L, M, seed = 50_000, 10, 30, 5
    env = tp.BatchedTetris(n, L, M, seed=seed, config_pool=carve_pool)
    env.reset()
    ost = c_oracle.BatchState(n)
    oep, _, _ = c_oracle.rollout(ost, 0, seed, L, M, carve_pool.rows, carve_pool.pieces, carve_pool.npieces, 0, True)
    rng = np.random.default_rng(4)
    dm = pc.distinct_module()
    for t in range(5):
        if t == 0:
            rows, runs, used = env.afterstates_distinct()
        else:
            rot, loc = rng.integers(0, 4, n).astype(np.uint8), rng.integers(0, 10, n).astype(np.uint8)
            dl, fl, st, rows, runs, used = env.step_observe_distinct(torch.from_numpy(rot).cuda(), torch.from_numpy(loc).cuda())
            odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
            assert np.array_equal(dl.cpu().numpy(), odl) and np.array_equal(st.cpu().numpy(), ost.state)
            _oracle_autoreset(ost, oep, seed, 0, carve_pool)
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
        grid = of.copy(); grid[:, :, 0] |= ofl << 3
        r = rows.cpu().numpy().view(np.uint32)[:int(used)]
        d = runs.cpu().numpy().view(np.uint32)
        pc.check_distinct(r, d, grid, f"step {t}")
        assert np.array_equal(env.expand_distinct(rows, runs).cpu().numpy().transpose(1, 0, 2), grid)
        idx, valid, slot = dm.gather_index(runs)
        w = rows.view(torch.uint8).view(-1, 4)[idx]                                    # [n, 34, 4] padded placements
        g = torch.from_numpy(grid).cuda()
        pick = g[torch.arange(n, device="cuda")[:, None], slot.clamp(max=39)]          # the grid at each placement's (rot, loc)
        assert bool((w[valid] == pick[valid]).all())
        assert int(valid.sum()) == int(dm.tables()[0][d >> 29].sum())
    f = env.fields()
    assert np.array_equal(f["rows"].cpu().numpy(), ost.rows)


def test_env_on_non_current_device(tp, carve_pool):
    """every C call runs with the env's device current: drive an env while another device is current"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 4096
    env = tp.BatchedTetris(n, 10, 30, device="cuda:1", seed=1, config_pool=carve_pool)
    ref = tp.BatchedTetris(n, 10, 30, device="cuda:0", seed=1, config_pool=carve_pool)
    torch.cuda.set_device(0)
    for e in (env, ref):
        e.reset(); e.rollout_random(5)
    a, b = env.afterstates(packed=True)[0], ref.afterstates(packed=True)[0]
    assert torch.equal(a.cpu(), b.cpu())


def test_error_codes(tp):
    import ctypes
    L = tp._lib.lib()
    assert L.tpl_step(None, 0, 4, None, None, None, None, None, None, 1, 1, None) == -1
    assert b"null" in L.tpl_last_error()
    assert L.tpl_gen_pieces(ctypes.c_void_p(8), 4, 42 * 256 + 1, 0, 0, None, 0, None) == -2
    with pytest.raises(tp.TplError):
        tp._lib.check(L.tpl_afterstates(ctypes.c_void_p(8), 0, 4, None, None, None, 1, 1, None), "x")
    h = tp.HostBatchedTetris(4, 3, 3)
    with pytest.raises(tp.TplError):
        h.reset()                                  # no pool yet
    h.close()


def _torch_features(rows):
    """Independent torch restatement of (holes, bumpiness, agg height) on uint16 bitrows [N,20]."""
    import torch
    r = rows.to(torch.int32)
    cells = ((r[:, :, None] >> torch.arange(10, device=r.device)) & 1)           # [N,20,10]
    filled = torch.cummax(cells, dim=1).values                                    # 1 from the first filled row down
    h = filled.sum(dim=1)                                                         # [N,10]
    agg = h.sum(dim=1)
    bump = (h[:, 1:] - h[:, :-1]).abs().sum(dim=1)
    holes = agg - cells.sum(dim=(1, 2))
    return holes, bump, agg


def test_full_size_properties(tp):
    """BASELINE config sizes (2^20 envs): size-independent properties instead of an oracle pass.
    (1) the features promised by afterstates for the slot that is then played equal the features of the board
        the move really produces; (2) cell conservation: cells' = cells + 4 - 10 * rows_cleared;
    (3) flags agree with the state transition; (4) the fused rollout equals the same steps done one by one."""
    import torch
    n, L, M = 1 << 20, 10, 30
    pool = tp.synthetic_pool(4096, seed=0, M=M)
    env = tp.BatchedTetris(n, L, M, seed=0, config_pool=pool)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ar = torch.arange(n, device="cuda")
    for t in range(6):
        rows0 = env.fields()["rows"].clone()
        cells0 = _torch_features(rows0)[2] - _torch_features(rows0)[0]
        feats, flags = env.afterstates()
        rot = torch.randint(0, 4, (n,), device="cuda", generator=g)
        loc = torch.randint(0, 10, (n,), device="cuda", generator=g)
        pf = feats[ar, rot, loc].to(torch.int32)            # promised (dlines, holes, bump, agg)
        pfl = flags[ar, rot, loc]
        dl, fl, st = env.move(rot, loc)
        rows1 = env.fields()["rows"]
        holes, bump, agg = _torch_features(rows1)
        assert torch.equal(pf[:, 0], dl.to(torch.int32))
        assert torch.equal(pf[:, 1], holes.to(torch.int32)) and torch.equal(pf[:, 2], bump.to(torch.int32))
        assert torch.equal(pf[:, 3], agg.to(torch.int32))
        assert torch.equal(pfl & 7, fl & 7)
        top = (fl & 1).bool()
        cells1 = agg - holes
        assert torch.equal(cells1[~top], (cells0 + 4 - 10 * dl.to(torch.int64))[~top])
        assert torch.equal(rows1.view(torch.int16)[top], rows0.view(torch.int16)[top])
        env.reset(done_only=True)
    # fused rollout == stepwise path on the same counter-RNG schedule (compared through a checksum of records)
    a = tp.BatchedTetris(n, L, M, seed=5, config_pool=pool); a.reset()
    b = tp.BatchedTetris(n, L, M, seed=5, config_pool=pool); b.reset()
    a.rollout_random(7)
    a.rollout_random(6)
    b.rollout_random(13)
    assert torch.equal(a.state, b.state) and torch.equal(a.stats, b.stats) and torch.equal(a.episode, b.episode)
    assert int(a.stats[6]) == 13 * n


def test_all_entry_points_on_ragged_sizes(tp):
    """Every kernel and API layer once on n = 1, 33, 1000, 40 000 (split kernel, tile kernels, ragged tails).
    (compute-sanitizer is closed on this GPU pool; this is the bounds smoke that remains, next to the parity tests.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "sanitize_case.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sanitize case done" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_forward_pool_feeds_the_env(tp):
    """Reset points of the reference's forward producer (game/tetris.py:482-488) as a device-resident config pool."""
    import random
    random.seed(5)
    pool = tp.forward_pool(3, 20, 0, 40, 4, 300)
    assert pool.K > 10 and (pool.npieces == 21).all()
    n = 1000
    env = tp.BatchedTetris(n, 3, 20, seed=2, config_pool=pool)
    idx = np.arange(n, dtype=np.int32) % pool.K
    env.reset(idx=idx)
    f = env.fields(queue=True)
    assert np.array_equal(f["rows"].cpu().numpy(), pool.rows[idx])
    assert np.array_equal(f["queue"].cpu().numpy()[:, :21], pool.pieces[idx, :21])
    env.rollout_greedy(20, [760, -360, -180, -510, 100000, -100000])
    stats = env.stats.cpu().numpy()
    assert stats[0] > 0 and stats[1] > 0                     # the greedy policy wins some of these 3-line games
    env.terminate()


def test_checkpoint_resume(tp, carve_pool, tmp_path):
    """state_dict -> torch.save -> load_state_dict into a fresh object continues the rollout bit for bit (records, episode
    counters, counter-RNG position, statistics)."""
    import torch
    n = 20_000
    a = tp.BatchedTetris(n, 10, 30, seed=5, config_pool=carve_pool, env_base=3)
    a.reset(); a.rollout_random(25); a.rollout_greedy(5, [760, -360, -180, -510, 100000, -100000])
    path = os.path.join(tmp_path, "ckpt.pt")
    torch.save(a.state_dict(), path)
    a.rollout_random(30)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rot = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
    loc = torch.randint(0, 10, (n,), device="cuda", dtype=torch.uint8, generator=g)
    ra = a.step_observe(rot, loc, packed=True)
    b = tp.BatchedTetris(n, 1, 1, seed=0, config_pool=carve_pool)          # different L / M / seed: all restored
    b.load_state_dict(torch.load(path))
    b.rollout_random(30)
    rb = b.step_observe(rot, loc, packed=True)
    assert torch.equal(a.state, b.state) and torch.equal(a.episode, b.episode) and torch.equal(a.stats, b.stats)
    assert torch.equal(ra[3], rb[3]) and (b.L, b.M, b.seed, b.env_base) == (10, 30, 5, 3)
    with pytest.raises(ValueError):
        tp.BatchedTetris(n + 1, 10, 30, config_pool=carve_pool).load_state_dict(torch.load(path))
