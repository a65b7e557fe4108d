"""CPU tier: the N>1 path on gloo, world_size 2.  Each rank rolls its shard (the per-env device code compiled for
the host stands in for the GPU kernel -- test infrastructure only), the statistics are all-reduced through the
package's own helper, and the result must equal the single-process run: sharding by global env id is invisible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _pool(golden_dir):
    z = np.load(os.path.join(golden_dir, "carve_pool_L10_M30.npz"))
    pieces = np.zeros((len(z["rows"]), 42), np.uint8)
    pieces[:, :z["pieces"].shape[1]] = z["pieces"]
    return z["rows"], pieces, z["npieces"]


def _roll(eng, pool_arrays, base, count, steps, seed, weights):
    pool = eng.make_pool(*pool_arrays)
    s = eng.empty_states(count)
    ep, ts = np.zeros(count, np.uint32), np.zeros(count, np.uint32)
    eng.reset(s, pool, episode=ep, seed=seed, env_base=base)
    stats = eng.rollout(s, pool, ep, ts, steps, seed, base, 0, 10, 30, weights)
    return eng.raw(s), ep, stats


def _worker(rank, world, port, golden_dir, total, steps, seed, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import tetris_piclim as tp
    from importlib import import_module
    D = import_module(tp.__name__ + ".distributed")
    from tests.engines import EmulEngine
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    base, count = D.shard_bounds(total, world, rank)
    rec, ep, stats = _roll(EmulEngine(), _pool(golden_dir), base, count, steps, seed, None)
    red = D.all_reduce_stats(torch.from_numpy(stats))
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, rec, ep))
    if rank == 0:
        recs = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])])
        eps = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        np.savez(out, recs=recs, eps=eps, stats=red.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    import tetris_piclim as tp
    from importlib import import_module
    D = import_module(tp.__name__ + ".distributed")
    for total, world in [(8 << 20, 8), (1000, 3), (5, 8), (0, 2)]:
        parts = [D.shard_bounds(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == total
        for (b0, c0), (b1, _) in zip(parts, parts[1:]):
            assert b0 + c0 == b1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)
    s = torch.arange(8, dtype=torch.int64)
    assert torch.equal(D.all_reduce_stats(s), s)            # not initialised: identity


@pytest.mark.timeout(300)
def test_two_ranks_equal_one(golden_dir, tmp_path, emul):
    total, steps, seed = 1501, 60, 77
    out = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, _free_port(), golden_dir, total, steps, seed, out), nprocs=2, join=True)
    z = np.load(out)
    rec, ep, stats = _roll(emul, _pool(golden_dir), 0, total, steps, seed, None)
    assert np.array_equal(z["recs"], rec) and np.array_equal(z["eps"], ep)
    assert np.array_equal(z["stats"], stats) and stats[6] == total * steps
