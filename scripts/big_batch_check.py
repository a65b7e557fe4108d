import sys, os
sys.path.insert(0, os.getcwd())
import torch, tetris_piclim as tp
n = 1 << 23
pool = tp.load_pool("tests/golden/carve_pool_L10_M30.npz")
a = tp.BatchedTetris(n, 10, 30, seed=1, config_pool=pool); a.reset()
b = tp.BatchedTetris(n, 10, 30, seed=1, config_pool=pool); b.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for t in range(14):
    rot = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
    loc = torch.randint(0, 10, (n,), device="cuda", dtype=torch.uint8, generator=g)
    ra = a.step_observe(rot, loc, packed=True)
    b.move(rot, loc); b.reset(done_only=True)
    fb = b.afterstates(packed=True, raw=True)[0]
    assert torch.equal(ra[3], fb) and torch.equal(a.state, b.state) and torch.equal(a.episode, b.episode), t
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(10):
    a.step_observe(rot, loc, packed=True)
e1.record(); torch.cuda.synchronize()
print("8M envs: fused == separate kernels for 14 steps;", e0.elapsed_time(e1) / 10, "ms per fused step", n * 40 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9, "G afterstates/s")
