"""Parity cases shared by the CPU tier (EmulEngine) and the GPU tier (GpuEngine).

Every expected value comes either from the committed golden fixtures (outputs of the UNMODIFIED reference,
tests/golden/make_golden.py) or from the oracle (oracle/piclim_oracle.{py,c}), itself pinned to the live
reference on >= 1e5 episodes.  All comparisons are bit-exact (integer/byte work).
"""
from __future__ import annotations

import os

import numpy as np

from oracle import c_oracle, piclim_oracle as po

P = 42
FLAG_TOPOUT, FLAG_WIN, FLAG_LOSE, FLAG_ALIAS, FLAG_NOPIECE = 1, 2, 4, 8, 16


def distinct_module():
    """the product's host helper for the distinct-placements form (tables from the library, numpy expansion)"""
    from importlib import import_module
    import tetris_piclim
    return import_module(tetris_piclim.__name__ + ".distinct")


def check_distinct(rows, runs, grid_packed, what=""):
    """rows / runs (distinct-placements form) against the compact 40-slot grid uint8[n, 40, 4] of the same states:
    expand(rows, runs) == grid, run lengths follow the pieces, runs do not overlap, nothing but runs and zero gaps is used."""
    dm = distinct_module()
    count, slot_of, canon_of = dm.tables()
    n = len(runs)
    piece, off = (runs >> 29).astype(np.int64), (runs & 0x1FFFFFFF).astype(np.int64)
    cnt = count[piece].astype(np.int64)
    nopiece = (grid_packed[:, 0, 0] >> 3) & FLAG_NOPIECE != 0
    assert np.array_equal(piece == 7, nopiece), f"{what}: run descriptors disagree with the grid about empty queues"
    assert (off + cnt <= len(rows)).all(), f"{what}: a run leaves the used part of rows"
    assert np.array_equal(dm.expand(rows, runs), grid_packed), f"{what}: expand(distinct) != 40-slot grid"
    used = np.zeros(len(rows) + 1, np.int64)
    np.add.at(used, off, 1); np.add.at(used, off + cnt, -1)
    cover = np.cumsum(used)[:-1]
    assert cover.max(initial=0) <= 1, f"{what}: runs overlap"
    assert not rows[cover == 0].any(), f"{what}: words outside every run must be zero (alignment gaps)"
    assert int(cnt.sum()) <= len(rows) <= int(cnt.sum()) + 3 * ((n + 31) // 32) + 3
    # no alias flag inside the runs
    sel = cover == 1
    assert not ((rows[sel] & 0xFF) >> 3 & FLAG_ALIAS).any(), f"{what}: alias flag in a distinct placement"


# ---------------------------------------------------------------------------------------------
# input generators
# ---------------------------------------------------------------------------------------------
def adversarial_boards(rng: np.random.Generator, n: int) -> np.ndarray:
    """uint16[n, 20]: heights 0..20, densities 0.3..0.9, planted full rows, top-row cells, wells."""
    H = rng.integers(0, 21, n)
    dens = rng.uniform(0.3, 0.9, n)
    cells = rng.random((n, 20, 10)) < dens[:, None, None]
    rowidx = np.arange(20)[None, :, None]
    cells &= rowidx >= (20 - H)[:, None, None]
    kind = rng.integers(0, 6, n)
    k0 = np.where((kind == 0) & (H > 0))[0]
    if len(k0):
        r = 20 - 1 - (rng.integers(0, 1 << 30, len(k0)) % H[k0])
        cells[k0, r, :] = True
    k1 = np.where(kind == 1)[0]
    cells[k1, 0, rng.integers(0, 10, len(k1))] = True
    k2 = np.where((kind == 2) & (H >= 3))[0]
    for i in k2:
        cells[i, 20 - H[i] + 1:, rng.integers(0, 10)] = False
    return (cells.astype(np.uint16) << np.arange(10, dtype=np.uint16)).sum(axis=2).astype(np.uint16)


def alias_mask(piece: int) -> np.ndarray:
    m = np.zeros((4, 10), bool)
    for r in range(4):
        for c in range(10):
            m[r, c] = po.slot_is_alias(piece, r, c)
    return m


def oracle_state(rows, pieces, npieces, lines=None, moves=None, state=None, head=None):
    n = len(rows)
    st = c_oracle.BatchState(n)
    st.load(rows, pieces, npieces)
    if lines is not None: st.lines[:] = lines
    if moves is not None: st.moves[:] = moves
    if state is not None: st.state[:] = state
    if head is not None: st.head[:] = head
    return st


def assert_same(eng, s, ost: c_oracle.BatchState, what=""):
    u = eng.unpack(s)
    assert np.array_equal(u["rows"], ost.rows), f"{what}: boards differ"
    assert np.array_equal(u["lines"], ost.lines), f"{what}: lines_cleared differ"
    assert np.array_equal(u["moves"], ost.moves), f"{what}: moves_used differ"
    assert np.array_equal(u["state"], ost.state), f"{what}: state differs"
    assert np.array_equal(u["head"], ost.head), f"{what}: pieces consumed differ"
    assert np.array_equal(u["npieces"], ost.npieces), f"{what}: npieces differ"
    valid = np.arange(P)[None, :] < ost.npieces[:, None]
    assert np.array_equal(np.where(valid, u["queue"], 0), np.where(valid, ost.pieces, 0)), f"{what}: piece queue differs"
    has = ost.head < ost.npieces
    cur = np.where(has, ost.pieces[np.arange(ost.n), np.minimum(ost.head, P - 1)], 255)
    assert np.array_equal(u["cur"], cur.astype(np.uint8)), f"{what}: current piece differs"


# ---------------------------------------------------------------------------------------------
# golden fixtures (reference outputs)
# ---------------------------------------------------------------------------------------------
def case_golden_kat(eng, golden_dir):
    z = np.load(os.path.join(golden_dir, "kat_carve.npz"))
    for i in range(int(z["count"])):
        seed, L, M = (int(v) for v in z[f"k{i}_meta"])
        pieces = np.zeros((1, P), np.uint8)
        pcs = z[f"k{i}_pieces"]
        pieces[0, :len(pcs)] = pcs
        assert len(pcs) == M + 1                                  # game/tetris.py:281-284
        s = eng.pack(z[f"k{i}_rows"][None], pieces, [len(pcs)])
        for t, (rot, loc) in enumerate(z[f"k{i}_solution"]):
            dl, fl, st = eng.step(s, [rot], [loc], L, M)
            u = eng.unpack(s)
            lines, moves, state = (int(v) for v in z[f"k{i}_trace"][t])
            assert (int(u["lines"][0]), int(u["moves"][0]), int(u["state"][0])) == (lines, moves, state), (i, t)
            assert int(st[0]) == state
        assert int(u["state"][0]) == po.WON                       # game/main.py:49-57
        assert np.array_equal(u["rows"][0], z[f"k{i}_final_rows"])


def case_golden_moves(eng, golden_dir):
    z = np.load(os.path.join(golden_dir, "moves_random.npz"))
    LM = z["LM"]
    checked = 0
    for L, M in sorted({(int(a), int(b)) for a, b in LM}):
        sel = np.where((LM[:, 0] == L) & (LM[:, 1] == M))[0]
        s = eng.pack(z["rows0"][sel], z["pieces"][sel], z["npieces"][sel])
        nm = z["nmoves"][sel]
        prev_rows = z["rows0"][sel].copy()
        prev_meta = np.zeros((len(sel), 4), np.int32)
        prev_meta[:, 3] = z["npieces"][sel]
        for t in range(int(nm.max())):
            act = z["actions"][sel, t]
            live = t < nm
            # finished episodes keep receiving a harmless move; only live ones are compared
            rot = np.where(live, act[:, 0], 0)
            loc = np.where(live, act[:, 1], 0)
            dl, fl, st = eng.step(s, rot, loc, L, M)
            u = eng.unpack(s)
            exp_rows, exp_meta = z["after_rows"][sel, t], z["after_meta"][sel, t]
            assert np.array_equal(u["rows"][live], exp_rows[live]), (L, M, t)
            assert np.array_equal(u["lines"][live], exp_meta[live, 0])
            assert np.array_equal(u["moves"][live], exp_meta[live, 1])
            assert np.array_equal(u["state"][live], exp_meta[live, 2])
            assert np.array_equal((u["npieces"].astype(int) - u["head"])[live], exp_meta[live, 3])
            assert np.array_equal(dl[live], (exp_meta[live, 0] - prev_meta[live, 0]).astype(np.int8))
            topout = exp_meta[live, 1] == prev_meta[live, 1]
            assert np.array_equal((fl[live] & FLAG_TOPOUT) != 0, topout)
            prev_meta = np.where(live[:, None], exp_meta, prev_meta)
            checked += int(live.sum())
    assert checked == int(z["nmoves"].sum())
    return checked


def case_golden_afterstates(eng, golden_dir, with_boards=True):
    z = np.load(os.path.join(golden_dir, "afterstates.npz"))
    meta = z["meta"]
    total = 0
    for L, M in sorted({(int(a), int(b)) for a, b in meta[:, :2]}):
        sel = np.where((meta[:, 0] == L) & (meta[:, 1] == M))[0]
        args = (z["rows"][sel], z["pieces"][sel], z["npieces"][sel], meta[sel, 2], meta[sel, 3])
        s = eng.pack(*args)
        feats, flags, ff = eng.afterstates(s, L, M, f32=True)
        exp_flags = z["flags"][sel].copy()
        for j, i in enumerate(sel):
            exp_flags[j] |= np.where(alias_mask(int(z["pieces"][i, 0])), FLAG_ALIAS, 0).astype(np.uint8)
        assert np.array_equal(feats, z["feats"][sel]), (L, M)
        assert np.array_equal(flags, exp_flags), (L, M)
        assert np.array_equal(ff, z["feats"][sel].astype(np.float32))
        if with_boards:                                   # slot (r, c) must also be the board move(r, c) produces
            for r in range(4):
                for c in range(10):
                    s2 = eng.pack(*args)
                    eng.step(s2, np.full(len(sel), r), np.full(len(sel), c), L, M)
                    assert np.array_equal(eng.unpack(s2)["rows"], z["boards"][sel, r, c]), (L, M, r, c)
        total += len(sel)
    return total


# ---------------------------------------------------------------------------------------------
# randomized parity against the C oracle
# ---------------------------------------------------------------------------------------------
def case_random_moves(eng, n, steps, L, M, seed, wide_actions=True):
    rng = np.random.default_rng(seed)
    rows = adversarial_boards(rng, n)
    npieces = np.full(n, min(M + 1, P), np.uint8)
    pieces = rng.integers(0, 7, (n, P)).astype(np.uint8)
    s = eng.pack(rows, pieces, npieces)
    ost = oracle_state(rows, pieces, npieces)
    assert_same(eng, s, ost, "after pack")
    for t in range(steps):
        rot = rng.integers(-2, 8, n) if wide_actions else rng.integers(0, 4, n)
        loc = rng.integers(0, 13, n) if wide_actions else rng.integers(0, 10, n)
        dl, fl, st = eng.step(s, rot, loc, L, M)
        odl, ofl = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(dl, odl), f"step {t}: dlines"
        assert np.array_equal(fl & (FLAG_TOPOUT | FLAG_NOPIECE), ofl), f"step {t}: flags"
        assert np.array_equal(st, ost.state), f"step {t}: state"
        assert_same(eng, s, ost, f"step {t}")
    return ost


def case_afterstates_vs_oracle(eng, n, L, M, seed, pre_moves=3):
    rng = np.random.default_rng(seed)
    rows = adversarial_boards(rng, n)
    npieces = np.full(n, min(M + 1, P), np.uint8)
    pieces = rng.integers(0, 7, (n, P)).astype(np.uint8)
    lines = rng.integers(0, max(L, 1), n).astype(np.int32)
    moves = rng.integers(0, max(M, 1), n).astype(np.int32)
    head = np.minimum(moves, npieces - 1).astype(np.uint8)
    s = eng.pack(rows, pieces, npieces, lines, moves, None, head)
    ost = oracle_state(rows, pieces, npieces, lines, moves, None, head)
    for _ in range(pre_moves):
        rot, loc = rng.integers(0, 4, n), rng.integers(0, 10, n)
        eng.step(s, rot, loc, L, M)
        c_oracle.step_batch(ost, rot, loc, L, M)
    feats, flags, ff = eng.afterstates(s, L, M, f32=True)
    of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
    assert np.array_equal(feats.reshape(n, 40, 4), of)
    assert np.array_equal(flags.reshape(n, 40), ofl)
    assert np.array_equal(ff.reshape(n, 40, 4), of.astype(np.float32))
    assert_same(eng, s, ost, "afterstates must not modify the state")
    pk = eng.afterstates_packed(s, L, M).reshape(n, 40, 4)            # compact form: byte 0 = dlines | flags << 3
    exp = of.copy(); exp[:, :, 0] |= (ofl << 3)
    assert np.array_equal(pk, exp)
    rows_d, runs_d = eng.afterstates_distinct(s, L, M)                # distinct-placements form, expanded == the oracle's grid
    check_distinct(rows_d, runs_d, exp, "afterstates_distinct")
    assert_same(eng, s, ost, "afterstates_distinct must not modify the state")
    return of, ofl


# ---------------------------------------------------------------------------------------------
# RNG, reset, rollouts
# ---------------------------------------------------------------------------------------------
def case_rng(eng):
    for seed, base, ep, count in [(0, 0, 0, 31), (5, 1 << 33, 3, 42), (0xDEADBEEFCAFE, 12345, 7, 16), (1, 0, 0, 1)]:
        got = eng.gen_pieces(257, count, seed, base, ep)
        exp = c_oracle.gen_pieces(seed, base, 257, ep, count)
        assert np.array_equal(got, exp)
        for e in (0, 100, 256):
            assert list(got[e]) == po.gen_pieces(seed, base + e, ep, count)
        # the reference's 7-bag contract (game/main.py:20-29): every aligned group of 7 has no duplicates
        for g in range(0, count, 7):
            grp = got[:, g:g + 7]
            srt = np.sort(grp, axis=1)
            assert (srt[:, 1:] != srt[:, :-1]).all()
            assert grp.max() <= 6
    # sequences longer than the 42-piece queue (block-wise generation == the oracle's bag-by-bag definition)
    for count in (43, 61, 84, 100, 300):
        got = eng.gen_pieces(33, count, 21, 5, 2)
        assert np.array_equal(got, c_oracle.gen_pieces(21, 5, 33, 2, count))
        assert list(got[7]) == po.gen_pieces(21, 5 + 7, 2, count)
    ep = np.arange(300, dtype=np.uint32) * 3
    got = eng.gen_pieces(300, 31, 9, 77, episode=ep)
    for e in (0, 17, 299):
        assert list(got[e]) == po.gen_pieces(9, 77 + e, int(ep[e]), 31)


def case_reset(eng, pool_arrays, n=1000, seed=3, env_base=1 << 20):
    prow, ppieces, pnp = pool_arrays
    K = len(prow)
    pool = eng.make_pool(prow, ppieces, pnp)
    s = eng.empty_states(n)
    # (1) explicit indices
    rng = np.random.default_rng(0)
    idx = rng.integers(0, K, n).astype(np.int32)
    eng.reset(s, pool, idx=idx)
    u = eng.unpack(s)
    assert np.array_equal(u["rows"], prow[idx]) and np.array_equal(u["npieces"], pnp[idx])
    assert not u["lines"].any() and not u["moves"].any() and not u["state"].any() and not u["head"].any()
    valid = np.arange(P)[None, :] < pnp[idx][:, None]
    assert np.array_equal(np.where(valid, u["queue"], 0), np.where(valid, ppieces[idx], 0))
    # (2) counter-RNG draw == oracle.config_index
    episode = np.full(n, 5, np.uint32)
    eng.reset(s, pool, episode=episode, seed=seed, env_base=env_base)
    u = eng.unpack(s)
    exp = np.array([po.config_index(seed, env_base + i, 5, K) for i in range(n)])
    assert np.array_equal(u["rows"], prow[exp])
    # (3) masked reset leaves the others alone
    L, M = 4, 7
    for _ in range(4):
        eng.step(s, rng.integers(0, 4, n), rng.integers(0, 10, n), L, M)
    before = eng.unpack(s)
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    eng.reset(s, pool, idx=idx, mask=mask, mode=1)
    after = eng.unpack(s)
    m = mask.astype(bool)
    assert np.array_equal(after["rows"][~m], before["rows"][~m]) and np.array_equal(after["moves"][~m], before["moves"][~m])
    assert np.array_equal(after["rows"][m], prow[idx][m]) and not after["moves"][m].any()
    # (3b) masked reset WITHOUT explicit indices starts a new episode for the masked envs (counter bumped before the draw)
    episode = np.full(n, 4, np.uint32)
    tstep = np.full(n, 9, np.uint32)
    eng.reset(s, pool, mask=mask, mode=1, episode=episode, seed=seed, env_base=env_base, tstep=tstep)
    after = eng.unpack(s)
    assert np.array_equal(episode, np.where(m, 5, 4)) and np.array_equal(tstep, np.where(m, 0, 9))
    exp = np.array([po.config_index(seed, env_base + i, 5, K) for i in range(n)])
    assert np.array_equal(after["rows"][m], prow[exp][m]) and np.array_equal(after["rows"][~m], before["rows"][~m])
    # (4) auto-reset: only finished envs, episode counter bumped, draw keyed by the new episode
    for _ in range(5):
        eng.step(s, rng.integers(0, 4, n), rng.integers(0, 10, n), L, M)
    before = eng.unpack(s)
    done = (before["state"] != 0) | (before["head"] >= before["npieces"])
    assert done.any() and (~done).any()
    episode = np.full(n, 9, np.uint32)
    eng.reset(s, pool, mode=2, episode=episode, seed=seed, env_base=env_base)
    after = eng.unpack(s)
    assert np.array_equal(episode, np.where(done, 10, 9))
    exp = np.array([po.config_index(seed, env_base + i, 10, K) for i in range(n)])
    assert np.array_equal(after["rows"][done], prow[exp][done]) and not after["state"][done].any()
    assert np.array_equal(after["rows"][~done], before["rows"][~done])
    # (5) generated pieces replace the pool's
    eng.reset(s, pool, idx=idx, episode=np.full(n, 2, np.uint32), seed=seed, env_base=env_base, gen_count=31)
    u = eng.unpack(s)
    assert (u["npieces"] == 31).all()
    assert np.array_equal(u["queue"][:, :31], c_oracle.gen_pieces(seed, env_base, n, 2, 31))


def case_rollout(eng, pool_arrays, n, steps, L, M, seed, env_base, weights=None, chunks=(1.0,)):
    prow, ppieces, pnp = pool_arrays
    pool = eng.make_pool(prow, ppieces, pnp)
    s = eng.empty_states(n)
    episode, tstep = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    eng.reset(s, pool, episode=episode, seed=seed, env_base=env_base)           # episode 0
    ost = c_oracle.BatchState(n)
    oep, ots, ostats = c_oracle.rollout(ost, env_base, seed, L, M, prow, ppieces, pnp, 0, True)
    assert_same(eng, s, ost, "episode-0 install")
    stats = np.zeros(8, np.int64)
    done = 0
    for frac in chunks:                                                          # rollouts can be continued
        k = int(round(steps * frac))
        stats += eng.rollout(s, pool, episode, tstep, k, seed, env_base, 0, L, M, weights)
        _, _, st2 = c_oracle.rollout(ost, env_base, seed, L, M, prow, ppieces, pnp, k, False, oep, ots, nthreads=8, weights=weights)
        ostats += st2
        done += k
        assert_same(eng, s, ost, f"after {done} rollout steps")
        assert np.array_equal(episode, oep) and np.array_equal(tstep, ots)
        assert np.array_equal(stats, ostats), (stats, ostats)
    assert stats[6] == n * done
    return stats


def case_fused_step_observe(eng, pool_arrays, n=3000, steps=45, L=10, M=30, seed=8, env_base=123456789):
    """The fused kernel (move -> auto-reset -> afterstates) == the three separate calls, and == the oracle."""
    prow, ppieces, pnp = pool_arrays
    pool = eng.make_pool(prow, ppieces, pnp)
    rng = np.random.default_rng(seed)
    a, b, c = eng.empty_states(n), eng.empty_states(n), eng.empty_states(n)
    ep_a, ep_b, ep_c = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    ts_a, ts_c = np.full(n, 77, np.uint32), np.full(n, 77, np.uint32)
    eng.reset(a, pool, episode=ep_a, seed=seed, env_base=env_base, tstep=ts_a)
    eng.reset(b, pool, episode=ep_b, seed=seed, env_base=env_base)
    eng.reset(c, pool, episode=ep_c, seed=seed, env_base=env_base, tstep=ts_c)
    assert not ts_a.any() and not ts_c.any()                                    # every reset zeroes the action counter
    ts_a[:] = 5; ts_c[:] = 5
    ever_reset = np.zeros(n, bool)
    ost = c_oracle.BatchState(n)
    oep, ots, _ = c_oracle.rollout(ost, env_base, seed, L, M, prow, ppieces, pnp, 0, True)
    tot = np.zeros(8, np.int64)
    for t in range(steps):
        rot, loc = rng.integers(0, 4, n), rng.integers(0, 10, n)
        dl, fl, st, feats, afl, stats = eng.step_observe(a, rot, loc, pool, ep_a, seed, env_base, L, M, packed=(t % 2 == 1), tstep=ts_a)
        tot += stats
        # the same step with the afterstates in the distinct-placements form
        dl3, fl3, st3, rows_d, runs_d, stats3 = eng.step_observe(c, rot, loc, pool, ep_c, seed, env_base, L, M, tstep=ts_c, distinct=True)
        assert np.array_equal(dl, dl3) and np.array_equal(fl, fl3) and np.array_equal(st, st3) and np.array_equal(stats, stats3)
        assert np.array_equal(ep_a, ep_c) and np.array_equal(ts_a, ts_c) and np.array_equal(eng.raw(a), eng.raw(c))
        dl2, fl2, st2 = eng.step(b, rot, loc, L, M)
        eng.reset(b, pool, mode=2, episode=ep_b, seed=seed, env_base=env_base)
        f2, g2 = eng.afterstates(b, L, M)
        f2 = f2.reshape(n, 40, 4).transpose(1, 0, 2).copy(); g2 = g2.reshape(n, 40).T
        assert np.array_equal(dl, dl2) and np.array_equal(fl, fl2) and np.array_equal(st, st2)
        assert np.array_equal(ep_a, ep_b) and np.array_equal(eng.raw(a), eng.raw(b))
        grid = f2.copy(); grid[:, :, 0] |= (g2 << 3)
        check_distinct(rows_d, runs_d, np.ascontiguousarray(grid.transpose(1, 0, 2)), f"fused step {t}")
        if afl is None:
            f2[:, :, 0] |= (g2 << 3)
        else:
            assert np.array_equal(afl, g2)
        assert np.array_equal(feats, f2)
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(dl, odl) and np.array_equal(st, ost.state)
        done = np.where((ost.state != 0) | (ost.head >= ost.npieces))[0]
        ever_reset[done] = True
        assert np.array_equal(ts_a, np.where(ever_reset, 0, 5))                  # auto-reset zeroes tstep, nothing else touches it
        for i in done:
            oep[i] += 1
            k = po.config_index(seed, env_base + int(i), int(oep[i]), len(prow))
            ost.rows[i] = prow[k]; ost.pieces[i] = ppieces[k]; ost.npieces[i] = pnp[k]
            ost.head[i] = 0; ost.lines[i] = 0; ost.moves[i] = 0; ost.state[i] = 0
        assert_same(eng, a, ost, f"fused step {t}")
    assert tot[6] == n * steps and (tot[0] > 0 or steps < 30) and tot[7] == tot[0]


def case_long_episodes(eng, n=400, L=200, M=60, seed=13, env_base=1 << 21):
    """Episodes longer than the 42-piece queue (M = 60 -> 61 pieces): the kernels refill the queue from the counter-based
    sequence when it runs dry (game/tetris.py:95-102: sequences of any length).  Driven by a host-side greedy heuristic on the
    oracle's afterstate features so that episodes actually last 60 moves; even steps go through the fused step, odd steps
    through tpl_step + tpl_reset_from_pool(TPL_RESET_DONE), which refills as well."""
    G = M + 1
    rng = np.random.default_rng(seed)
    K = 16
    prow = np.zeros((K, 20), np.uint16)
    prow[1:, 19] = rng.integers(1, 1023, K - 1).astype(np.uint16)               # (almost) empty boards: long episodes
    ppieces, pnp = np.zeros((K, P), np.uint8), np.full(K, 1, np.uint8)
    pool = eng.make_pool(prow, ppieces, pnp)
    idx = rng.integers(0, K, n).astype(np.int32)
    s = eng.empty_states(n)
    ep = np.zeros(n, np.uint32)
    eng.reset(s, pool, idx=idx, episode=ep, seed=seed, env_base=env_base, gen_count=G)
    ost = c_oracle.BatchState(n, P=64)
    seq = c_oracle.gen_pieces(seed, env_base, n, 0, G)
    ost.load(prow[idx], seq, np.full(n, G, np.uint8))
    oep = np.zeros(n, np.uint32)
    longest = 0
    for t in range(M + 25):
        of, ofl, _ = c_oracle.afterstates_batch(ost, L, M)
        bad = (ofl & (FLAG_TOPOUT | FLAG_ALIAS | FLAG_NOPIECE)) != 0
        score = of[:, :, 1].astype(np.int64) * 8 + of[:, :, 3].astype(np.int64) + of[:, :, 2].astype(np.int64) - 20 * of[:, :, 0].astype(np.int64)
        slot = np.where(bad, 1 << 30, score).argmin(axis=1)
        rot, loc = slot // 10, slot % 10
        if t % 2 == 0:
            dl, fl, st, _, _, _ = eng.step_observe(s, rot, loc, pool, ep, seed, env_base, L, M, packed=True, gen_count=G)
        else:
            dl, fl, st = eng.step(s, rot, loc, L, M)
        odl, _ = c_oracle.step_batch(ost, rot, loc, L, M)
        assert np.array_equal(dl, odl) and np.array_equal(st, ost.state), f"step {t}"
        longest = max(longest, int(ost.moves.max()))
        # episodes that ended start the next one: same board class (explicit idx is not available to the fused step, so the
        # config is the counter-RNG draw), a fresh 61-piece sequence keyed by the new episode number
        done = np.where((ost.state != 0) | (ost.head >= ost.npieces))[0]
        for i in done:
            oep[i] += 1
            k = po.config_index(seed, env_base + int(i), int(oep[i]), K)
            ost.rows[i] = prow[k]; ost.pieces[i] = 0
            ost.pieces[i, :G] = c_oracle.gen_pieces(seed, env_base + int(i), 1, int(oep[i]), G)[0]
            ost.npieces[i] = G; ost.head[i] = 0; ost.lines[i] = 0; ost.moves[i] = 0; ost.state[i] = 0
        if t % 2 == 1:
            eng.reset(s, pool, mode=2, episode=ep, seed=seed, env_base=env_base, gen_count=G)
        u = eng.unpack(s)
        assert np.array_equal(ep, oep), f"step {t}: episode numbers"
        assert np.array_equal(u["rows"], ost.rows) and np.array_equal(u["lines"], ost.lines), f"step {t}: boards / lines"
        assert np.array_equal(u["moves"], ost.moves) and np.array_equal(u["state"], ost.state), f"step {t}: moves / state"
        cur = ost.pieces[np.arange(n), np.minimum(ost.head, 63)]
        assert np.array_equal(u["cur"], np.where(ost.head < ost.npieces, cur, 255).astype(np.uint8)), f"step {t}: current piece"
    assert longest == M, f"no episode reached the move limit ({longest})"


# ---------------------------------------------------------------------------------------------
# edge cases the reference's semantics single out (SURVEY.md section 8c micro-cases)
# ---------------------------------------------------------------------------------------------
def _one(eng, rows, pieces, L, M, rot, loc, lines=0, moves=0):
    p = np.zeros((1, P), np.uint8); p[0, :len(pieces)] = pieces
    s = eng.pack(np.array([rows], np.uint16), p, [len(pieces)], [lines], [moves])
    dl, fl, st = eng.step(s, [rot], [loc], L, M)
    return eng.unpack(s), int(dl[0]), int(fl[0]), int(st[0]), s


def case_edges(eng):
    E = [0] * 20
    # untouched pre-existing full row is never cleared (:382-383 inspects only the piece's rows)
    rows = list(E); rows[19] = 0x3FF
    u, dl, fl, st, _ = _one(eng, rows, [6, 0], 5, 5, 0, 0)
    assert dl == 0 and u["rows"][0][19] == 0x3FF and u["rows"][0][18] == 3 and u["rows"][0][17] == 3
    # drop == 0 is legal: a vertical I on a column of height 16
    rows = list(E)
    for r in range(4, 20): rows[r] = 1
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 5, 5, 1, 0)
    assert not fl & FLAG_TOPOUT and u["moves"][0] == 1 and all(u["rows"][0][r] & 1 for r in range(20))
    # one row higher tops out: piece consumed, board and moves untouched, state lost (:356, :372-374)
    rows[3] = 1
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 5, 5, 1, 0)
    assert fl & FLAG_TOPOUT and st == po.LOST and u["moves"][0] == 0 and u["head"][0] == 1
    assert list(u["rows"][0]) == rows
    # the survey's counter-example for "slide from row 0" semantics: only cell (0,0) set, J rot 1 at loc 0 tops out
    rows = list(E); rows[0] = 1
    u, dl, fl, st, _ = _one(eng, rows, [2, 0], 5, 5, 1, 0)
    assert fl & FLAG_TOPOUT and list(u["rows"][0]) == rows
    # rot wraps with Python's % (:61): rot=5 -> 1, rot=-1 -> n_rot-1; loc=99 clamps to 10-w (:364)
    for rot, loc in [(5, 0), (-1, 0), (2, 99), (7, 12)]:
        u, dl, fl, st, _ = _one(eng, E, [1, 0], 5, 5, rot, loc)
        e = po.OracleEnv(5, 5).load(E, [1, 0]); e.move(rot, loc)
        assert list(u["rows"][0]) == e.rows, (rot, loc)
    # L=1, M=1 clearing move: win beats the move limit (:415-422)
    rows = list(E); rows[19] = 0x3FF & ~0xF
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 1, 1, 0, 0)
    assert dl == 1 and fl & FLAG_WIN and st == po.WON and u["rows"][0][19] == 0
    # same move with L=2: cleared but lost by the move limit
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 2, 1, 0, 0)
    assert dl == 1 and fl & FLAG_LOSE and st == po.LOST
    # 3-of-4 non-adjacent clear: vertical I completes rows 16,17,19 but not 18
    rows = list(E)
    for r in (16, 17, 19): rows[r] = 0x3FE
    rows[18] = 0x1FE
    rows[15] = 0x200
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 9, 9, 1, 0)
    e = po.OracleEnv(9, 9).load(rows, [0, 0]); k = e.move(1, 0)
    assert dl == 3 == k and list(u["rows"][0]) == e.rows and u["rows"][0][19] == 0x1FF and u["rows"][0][18] == 0x200
    # 4-line clear (tetris) and lines bookkeeping
    rows = list(E)
    for r in range(16, 20): rows[r] = 0x3FE
    u, dl, fl, st, _ = _one(eng, rows, [0, 0], 9, 9, 1, 0, lines=3)
    assert dl == 4 and u["lines"][0] == 7 and not any(u["rows"][0])
    # empty queue: nothing happens, NOPIECE flagged (the reference raises IndexError)
    p = np.zeros((1, P), np.uint8)
    s = eng.pack(np.array([E], np.uint16), p, [0])
    dl, fl, st = eng.step(s, [0], [0], 5, 5)
    assert fl[0] == FLAG_NOPIECE and eng.unpack(s)["head"][0] == 0
    feats, flags = eng.afterstates(s, 5, 5)
    assert (flags == FLAG_NOPIECE).all() and not feats.any()
    # full 42-piece queue round-trips through the 128-bit packing
    q = (np.arange(42) * 5 % 7).astype(np.uint8)[None]
    s = eng.pack(np.array([E], np.uint16), q, [42])
    u = eng.unpack(s)
    assert np.array_equal(u["queue"], q) and u["cur"][0] == q[0, 0] and u["next"][0] == q[0, 1]
    for t in range(42):
        eng.step(s, [0], [t % 10], 99, 99)
        u = eng.unpack(s)
        assert u["head"][0] == t + 1
        assert u["cur"][0] == (q[0, t + 1] if t + 1 < 42 else 255)
    # ragged batch sizes (not multiples of the warp / block size)
    for n in (1, 31, 33, 129):
        case_random_moves(eng, n, 3, 4, 9, seed=n, wide_actions=True)
