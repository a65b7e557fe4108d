import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench
import tetris_piclim as tp
dev = torch.device("cuda", 0)
n = 1 << 20
env = tp.BatchedTetris(n, bench.L_LINES, bench.M_MOVES, device=dev, seed=bench.SEED, config_pool=bench.make_pool(tp))
env.reset(); env.rollout_random(8); env.reset(done_only=True); env.stats.zero_()
g = torch.Generator(device=dev); g.manual_seed(1234)
K, W = 20, 5
rot = torch.randint(0, 4, (K + W, n), device=dev, dtype=torch.uint8, generator=g)
loc = torch.randint(0, 10, (K + W, n), device=dev, dtype=torch.uint8, generator=g)
for mode in ("sleep", "nosleep", "sleep"):
    for i in range(W): env.step_observe(rot[i], loc[i], packed=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    if mode == "sleep": torch.cuda._sleep(int(2.0e6))
    ev[0].record()
    for i in range(K):
        env.step_observe(rot[W + i], loc[W + i], packed=True)
        ev[i + 1].record()
    torch.cuda.synchronize()
    print(mode, "total/K %.4f" % (ev[0].elapsed_time(ev[K]) / K), " ".join(f"{ev[k].elapsed_time(ev[k + 1]):.4f}" for k in range(K)))
    # without per-step events
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if mode == "sleep": torch.cuda._sleep(int(2.0e6))
    e0.record()
    for i in range(K): env.step_observe(rot[W + i], loc[W + i], packed=True)
    e1.record(); torch.cuda.synchronize()
    print(mode, "no per-step events: %.4f" % (e0.elapsed_time(e1) / K))
