"""GPU tier: the fused tensor-core value net (tpl_value_rows) and the action selection (tpl_select_action) against plain
PyTorch.  Floating point: bf16 operands / activations with fp32 accumulation on both sides; the accumulation ORDER differs
(tcgen05 vs. torch matmul), and a last-bit difference can flip one bf16 rounding of an activation, so values are compared
within  |kernel - torch| <= 0.02 + 0.02 * |torch|  (observed: ~1e-3), and an arg-max counts as equal when the value of the
chosen placement is within that tolerance of the best one."""
import os
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GAMMA = 0.99


@pytest.fixture(scope="module")
def mods(gpu):
    import tetris_piclim as tp
    return tp, import_module(tp.__name__ + ".value_kernel"), import_module(tp.__name__ + ".distinct"), import_module(tp.__name__ + ".model")


def _net(model, seed, torch):
    torch.manual_seed(seed)
    net = model.ValueNet().cuda()
    with torch.no_grad():                                   # biases away from zero so that every term of the epilogue matters
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4, net.layer5):
            layer.bias.uniform_(-0.5, 0.5)
    return net


def _close(a, b):
    return (a - b).abs() <= 0.02 + 0.02 * b.abs()


@pytest.mark.parametrize("nrows", [1, 127, 128, 129, 300_007])
def test_value_rows_vs_torch(mods, nrows):
    import torch
    tp, vk, dm, model = mods
    net = _net(model, 3, torch)
    k = vk.ValueKernel(net)
    g = torch.Generator(device="cuda"); g.manual_seed(nrows)
    feats = torch.stack([torch.randint(0, 5, (nrows,), device="cuda", generator=g) | (torch.randint(0, 32, (nrows,), device="cuda", generator=g) << 3),
                         torch.randint(0, 120, (nrows,), device="cuda", generator=g), torch.randint(0, 100, (nrows,), device="cuda", generator=g),
                         torch.randint(0, 201, (nrows,), device="cuda", generator=g)], dim=1).to(torch.uint8)
    rows = feats.contiguous().view(torch.int32).view(-1)
    got = k.values(rows)
    ref = vk.reference_values(net, rows)
    assert torch.isfinite(got).all()
    assert bool(_close(got, ref).all()), float((got - ref).abs().max())
    assert float((got - ref).abs().mean()) < 5e-3
    # a device-side row count: rows past it are left untouched
    if nrows > 200:
        cnt = torch.tensor([nrows - 150], dtype=torch.int32, device="cuda")
        out = torch.full((nrows,), -7.0, device="cuda")
        k.values(rows, cnt, out)
        assert bool((out[nrows - 150:] == -7.0).all()) and bool(_close(out[:nrows - 150], ref[:nrows - 150]).all())
    # new parameters are seen after sync()
    with torch.no_grad():
        net.layer5.bias += 3.0
    k.sync()
    assert bool(_close(k.values(rows), ref + 3.0).all())


def test_select_action_on_env(mods, golden_dir):
    import torch
    tp, vk, dm, model = mods
    pool = tp.load_pool(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    n, L, M = 120_000, 10, 30
    env = tp.BatchedTetris(n, L, M, seed=2, config_pool=pool)
    env.reset(); env.rollout_random(6); env.reset(done_only=True)
    net = _net(model, 5, torch)
    k = vk.ValueKernel(net)
    count, slot_of, _ = dm.tables()
    rng = np.random.default_rng(0)
    for step in range(3):
        rot = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
        loc = torch.from_numpy(rng.integers(0, 10, n).astype(np.uint8)).cuda()
        dl, fl, st, rows, runs, used = env.step_observe_distinct(rot, loc)
        vals = k.values(rows, used.reshape(1))
        ref = vk.reference_values(net, rows[:int(used)])
        assert bool(_close(vals[:int(used)], ref).all())
        # greedy selection
        r_, c_, chosen, q = k.select(rows, runs, vals, GAMMA, 0.0, seed=9, step=step)
        idx, valid, slot = dm.gather_index(runs)
        w = rows[idx]                                                              # [n, 34] words (padded)
        b0 = w & 0xFF
        rew = (b0 & 7).float() + 10.0 * (((b0 >> 3) & 2) != 0).float() - 10.0 * (((b0 >> 3) & 5) != 0).float()
        qref = (rew + GAMMA * ref[idx.clamp(max=int(used) - 1)]).masked_fill(~valid, float("-inf"))
        best = qref.max(dim=1).values
        j = (slot == (r_.long() * 10 + c_.long())[:, None]) & valid                # the placement the kernel chose
        assert bool((j.sum(dim=1) == 1).all()), "rot / loc is not one of the env's distinct placements"
        qsel = (qref.masked_fill(~j, float("-inf"))).max(dim=1).values
        assert bool(_close(qsel, best).all()), "selected placement is not (within tolerance) the arg-max"
        assert bool((w[j] == chosen).all())
        # agreement with the production PyTorch path (rank_bf16) on the arg-max
        vt = net.rank_bf16(torch.stack([(rows[:int(used)] & 7).float(), ((rows[:int(used)] >> 8) & 0xFF).float(),
                                        ((rows[:int(used)] >> 16) & 0xFF).float(), ((rows[:int(used)] >> 24) & 0xFF).float()], dim=1)).float()
        qt = (rew + GAMMA * vt[idx.clamp(max=int(used) - 1)]).masked_fill(~valid, float("-inf"))
        # rank_bf16 rounds its OUTPUT to bf16 (8 bits of mantissa), so near-ties are common and the two arg-max indices need not
        # coincide; what must hold is that the kernel's choice is, under rank_bf16 too, within the tolerance of the best placement
        qt_sel = (qt.masked_fill(~j, float("-inf"))).max(dim=1).values
        assert float(_close(qt_sel, qt.max(dim=1).values).float().mean()) > 0.999
        # exploration: eps = 1 draws uniformly among the distinct placements
        r2, c2, ch2, _ = k.select(rows, runs, vals, GAMMA, 1.0, seed=9, step=step)
        j2 = (slot == (r2.long() * 10 + c2.long())[:, None]) & valid
        assert bool((j2.sum(dim=1) == 1).all())
        frac_first = float((j2.float().argmax(dim=1) == 0).float().mean())
        assert 0.02 < frac_first < 0.12                                            # E[1 / run length] = (3/17 + 3/34 + 1/9) / 7 = 0.054
        # drive the env with the greedy action: the kernel's (rot, loc) is what tpl_step consumes
        env.step_observe_distinct(r_, c_)
