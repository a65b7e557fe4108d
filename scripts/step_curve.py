"""Per-step time of the fused step over the first steps of a batch whose episodes all start together: how long until the mix of
boards is stationary (what bench.py's set-up must reach before timing).
    python scripts/step_curve.py [--pre R]      (R = random fused-rollout moves before the first timed step)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tetris_piclim as tp
ap = argparse.ArgumentParser(); ap.add_argument("--pre", type=int, default=8); ap.add_argument("--steps", type=int, default=96)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n = 1 << 20
env = tp.BatchedTetris(n, bench.L_LINES, bench.M_MOVES, device=dev, seed=bench.SEED, config_pool=bench.make_pool(tp))
env.reset(); env.rollout_random(a.pre); env.reset(done_only=True); env.stats.zero_()
g = torch.Generator(device=dev); g.manual_seed(1234)
rot = torch.randint(0, 4, (a.steps, n), device=dev, dtype=torch.uint8, generator=g)
loc = torch.randint(0, 10, (a.steps, n), device=dev, dtype=torch.uint8, generator=g)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps // 8 + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(a.steps):
    env.step_observe(rot[i], loc[i], packed=True)
    if i % 8 == 7: ev[i // 8 + 1].record()
torch.cuda.synchronize()
s = env.stats.cpu().tolist()
print("pre", a.pre, "ms per step by blocks of 8:", " ".join(f"{ev[k].elapsed_time(ev[k + 1]) / 8:.4f}" for k in range(a.steps // 8)))
print("episodes", s[0], "topouts", s[2], "resets", s[7], "steps", s[6])
