"""CPU tier: the oracles (Python + C restatements) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), the published Philox known-answer vectors, and each other."""
import os

import numpy as np
import pytest

from oracle import c_oracle, piclim_oracle as po, refshim
from tests import parity_cases as pc

P = 42


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        assert po.philox4x32(ctr, key) == exp
        assert c_oracle.philox(ctr, key) == exp


def test_seven_bag_contract():
    """game/main.py:20-29 (test_sequence): right length, every aligned group of 7 free of duplicates."""
    for seed in range(5):
        seq = po.gen_pieces(seed, 17, 3, 16)
        assert len(seq) == 16
        for i in range(0, 16, 7):
            grp = seq[i:i + 7]
            assert len(grp) == len(set(grp))
    # all 5040 permutations are reachable and equally weighted up to the 2^32/5040 rounding
    seen = {tuple(po.bag_from_word((k * 2 ** 32) // 5040 + 1)) for k in range(5040)}
    assert len(seen) == 5040
    a = c_oracle.gen_pieces(9, 100, 64, 2, 42)
    assert all(list(a[i]) == po.gen_pieces(9, 100 + i, 2, 42) for i in range(64))


def test_oracle_golden_kat(golden_dir):
    z = np.load(os.path.join(golden_dir, "kat_carve.npz"))
    assert int(z["count"]) >= 4
    for i in range(int(z["count"])):
        seed, L, M = (int(v) for v in z[f"k{i}_meta"])
        e = po.OracleEnv(L, M).load(z[f"k{i}_rows"], z[f"k{i}_pieces"])
        cst = c_oracle.BatchState(1).load(z[f"k{i}_rows"][None], z[f"k{i}_pieces"][None], len(z[f"k{i}_pieces"]))
        for t, (rot, loc) in enumerate(z[f"k{i}_solution"]):
            e.move(int(rot), int(loc))
            c_oracle.step_batch(cst, [rot], [loc], L, M)
            exp = tuple(int(v) for v in z[f"k{i}_trace"][t])
            assert (e.lines_cleared, e.moves_used, e.state) == exp
            assert (int(cst.lines[0]), int(cst.moves[0]), int(cst.state[0])) == exp
        assert e.state == po.WON and e.rows == [int(x) for x in z[f"k{i}_final_rows"]]
        assert np.array_equal(cst.rows[0], z[f"k{i}_final_rows"])


def test_oracle_golden_moves(golden_dir):
    z = np.load(os.path.join(golden_dir, "moves_random.npz"))
    n_checked = 0
    for e_i in range(len(z["rows0"])):
        L, M = (int(v) for v in z["LM"][e_i])
        npc = int(z["npieces"][e_i])
        py = po.OracleEnv(L, M).load(z["rows0"][e_i], z["pieces"][e_i, :npc])
        cst = c_oracle.BatchState(1).load(z["rows0"][e_i][None], z["pieces"][e_i][None], npc)
        for t in range(int(z["nmoves"][e_i])):
            rot, loc = (int(v) for v in z["actions"][e_i, t])
            py.move(rot, loc)
            c_oracle.step_batch(cst, [rot], [loc], L, M)
            exp_rows = [int(x) for x in z["after_rows"][e_i, t]]
            lines, moves, state, left = (int(v) for v in z["after_meta"][e_i, t])
            assert py.rows == exp_rows and (py.lines_cleared, py.moves_used, py.state, len(py.pieces)) == (lines, moves, state, left)
            assert [int(x) for x in cst.rows[0]] == exp_rows
            assert (int(cst.lines[0]), int(cst.moves[0]), int(cst.state[0]), npc - int(cst.head[0])) == (lines, moves, state, left)
            n_checked += 1
    assert n_checked == int(z["nmoves"].sum()) > 2000


def test_oracle_golden_afterstates(golden_dir):
    z = np.load(os.path.join(golden_dir, "afterstates.npz"))
    S = len(z["rows"])
    for s in range(S):
        L, M, lines, moves = (int(v) for v in z["meta"][s])
        npc = int(z["npieces"][s])
        exp_flags = z["flags"][s] | np.where(pc.alias_mask(int(z["pieces"][s, 0])), 8, 0).astype(np.uint8)
        cst = c_oracle.BatchState(1).load(z["rows"][s][None], z["pieces"][s][None], npc)
        cst.lines[0], cst.moves[0] = lines, moves
        f, fl, b = c_oracle.afterstates_batch(cst, L, M, want_boards=True)
        assert np.array_equal(f.reshape(4, 10, 4), z["feats"][s])
        assert np.array_equal(fl.reshape(4, 10), exp_flags)
        assert np.array_equal(b.reshape(4, 10, 20), z["boards"][s])
        if s % 8 == 0:                                     # the pure-Python oracle is slow: sample it
            e = po.OracleEnv(L, M).load(z["rows"][s], z["pieces"][s, :npc])
            e.lines_cleared, e.moves_used = lines, moves
            pf, pfl, pb = po.afterstates(e)
            assert np.array_equal(pf, z["feats"][s]) and np.array_equal(pfl, exp_flags) and np.array_equal(pb, z["boards"][s])


def test_python_and_c_oracles_agree_on_rollouts(golden_dir):
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    prow, pnp = z["rows"], z["npieces"]
    ppieces = np.zeros((len(prow), P), np.uint8); ppieces[:, :z["pieces"].shape[1]] = z["pieces"]
    n, steps, seed, base = 6, 70, 5, 1 << 34
    st = c_oracle.BatchState(n)
    ep, ts, stats = c_oracle.rollout(st, base, seed, 10, 30, prow, ppieces, pnp, steps, True)
    tot = [0] * 8
    for i in range(n):
        env, episode, s = po.rollout_random(seed, base + i, 10, 30, prow, ppieces, steps, pool_npieces=pnp)
        assert env.rows == [int(x) for x in st.rows[i]] and episode == int(ep[i])
        assert (env.lines_cleared, env.moves_used, env.state) == (int(st.lines[i]), int(st.moves[i]), int(st.state[i]))
        tot = [a + b for a, b in zip(tot, s)]
    assert tot[:7] == [int(v) for v in stats[:7]]


def test_reference_solutions_win_from_pool(golden_dir):
    """The recorded carve solutions (reference debug=True) win when replayed: game/main.py:49-57."""
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    for k in range(0, len(z["rows"]), 16):
        e = po.OracleEnv(10, 30).load(z["rows"][k], z["pieces"][k, :z["npieces"][k]])
        for rot, loc in z["solutions"][k, :z["nsol"][k]]:
            e.move(int(rot), int(loc))
        assert e.state == po.WON


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="live reference tree not present")
def test_oracles_vs_live_reference_sample():
    """A small live replay (the 1e5-episode run is oracle/validate_against_reference.py)."""
    from oracle.validate_against_reference import check_tables, run_chunk
    check_tables(refshim.load())
    moves, topouts, wins = run_chunk((99, 300))
    assert moves > 1500


def test_c_reset_done_matches_python_config_draw(golden_dir):
    """orc_reset_done (the C statement of the fused step's auto-reset) == the per-env Python restatement"""
    import os
    from oracle import c_oracle, piclim_oracle as po
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    prow, ppieces, pnp = z["rows"], z["pieces"], z["npieces"]
    n, seed, base = 500, 7, 1 << 34
    st = c_oracle.BatchState(n)
    ep, ts, _ = c_oracle.rollout(st, base, seed, 10, 30, prow, ppieces, pnp, 12, True)
    ref = st.copy(); ep_ref = ep.copy()
    done = (ref.state != 0) | (ref.head >= ref.npieces)
    assert done.any() and not done.all()
    ts[:] = 5
    cnt = c_oracle.reset_done(st, base, seed, prow, ppieces, pnp, ep, ts)
    assert cnt == int(done.sum())
    for i in range(n):
        if done[i]:
            k = po.config_index(seed, base + i, int(ep_ref[i]) + 1, len(prow))
            assert ep[i] == ep_ref[i] + 1 and ts[i] == 0
            assert np.array_equal(st.rows[i], prow[k]) and st.npieces[i] == pnp[k] and st.head[i] == 0 and st.state[i] == 0
            assert np.array_equal(st.pieces[i, :pnp[k]], ppieces[k, :pnp[k]]) and st.lines[i] == 0 and st.moves[i] == 0
        else:
            assert ep[i] == ep_ref[i] and ts[i] == 5 and np.array_equal(st.rows[i], ref.rows[i])
