"""The PyTorch side (value net + DQN loop): CPU checks of the network, GPU smoke of the loop."""
import pytest
import torch


def _mod(name):
    import tetris_piclim as tp
    from importlib import import_module
    return import_module(tp.__name__ + "." + name)


def test_model_matches_reference_layer_shape():
    m = _mod("model")
    net = m.Model(217, 14)                                  # the shapes model/train.py:26 asks for
    shapes = [tuple(p.shape) for p in net.parameters()]
    assert shapes == [(128, 217), (128,), (128, 128), (128,), (128, 128), (128,), (128, 128), (128,), (14, 128), (14,)]
    assert net(torch.zeros(3, 217)).shape == (3, 14)
    v = m.ValueNet()
    assert v(torch.zeros(5, 4, dtype=torch.uint8)).shape == (5,)
    x = torch.randint(0, 200, (64, 4)).to(torch.float32)
    assert torch.allclose(v.rank_bf16(x).float(), v(x), atol=0.05, rtol=0.05)      # bf16 ranking forward == fp32 forward (loosely)


def test_hyperparameters_are_the_reference_ones():
    t = _mod("train")
    assert (t.BATCH_SIZE, t.GAMMA, t.EPS_START, t.EPS_END, t.EPS_DECAY, t.TAU, t.LR) == (128, 0.99, 0.9, 0.05, 1000, 0.005, 1e-4)


def test_select_slots_never_picks_alias():
    t = _mod("train")
    g = torch.Generator(); g.manual_seed(0)
    values = torch.randn(40, 64)
    flags = torch.zeros(40, 64, dtype=torch.uint8)
    flags[10:] = 8
    values[10:] = 100.0
    for eps in (0.0, 0.5, 1.0):
        s = t.select_slots(values, flags, eps, g)
        assert int(s.max()) < 10


@pytest.mark.gpu
@pytest.mark.parametrize("value_kernel", [True, False])
def test_dqn_loop_runs_on_gpu(gpu, value_kernel):
    t = _mod("train")
    net, st = t.train(num_envs=4096, iterations=12, optim_steps_per_iter=2, value_kernel=value_kernel)
    assert st.env_steps == 4096 * 12 and st.optim_steps == 24
    assert st.loss == st.loss and st.episodes > 0            # finite loss, episodes finished and were reset


@pytest.mark.gpu
@pytest.mark.parametrize("value_kernel", [True, False])
def test_dqn_loop_learns(gpu, value_kernel):
    """The completed loop (model/train.py's constants) improves the behaviour policy: episodes that end late in a short run clear
    clearly more lines than those that end early -- on both rollout paths (fused tensor-core ranking kernel over the distinct
    placements / PyTorch forward over the 40-slot grid)."""
    import tetris_piclim as tp
    t = _mod("train")
    pool = tp.synthetic_pool(4096, seed=0, M=30)
    log = []

    def cb(it, eps, loss, s, prev):
        d = {k: s[k] - prev.get(k, 0) for k in s}
        log.append((d["lines"] / max(d["episodes"], 1), d["moves"] / max(d["episodes"], 1)))

    net, st = t.train(num_envs=8192, iterations=900, config_pool=pool, optim_steps_per_iter=8, batch_size=1024, log_every=150, log_fn=cb,
                      value_kernel=value_kernel)
    assert len(log) == 6 and st.loss == st.loss
    assert log[-1][0] > 1.5 * log[0][0] and log[-1][1] > log[0][1], log      # more lines and longer episodes
