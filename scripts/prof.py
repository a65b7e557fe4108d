"""Short, deterministic driver for ncu captures: W warm-up + K timed passes of one part of the hot path at 2^20 envs.

    python scripts/prof.py --what pipeline|fused|fused_distinct|afterstates_distinct|rollout_random|rollout_greedy|afterstates_f32 [--steps K] [--envs N]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import tetris_piclim as tp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="pipeline")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--nostats", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda", 0)
pool = bench.make_pool(tp)
env = tp.BatchedTetris(a.envs, bench.L_LINES, bench.M_MOVES, device=dev, seed=0, config_pool=pool)
env.count_stats = not a.nostats
env.reset()
env.rollout_random(8)
env.reset(done_only=True)
g = torch.Generator(device=dev); g.manual_seed(1)
tot = a.steps + a.warmup
rot = torch.randint(0, 4, (tot, a.envs), device=dev, dtype=torch.uint8, generator=g)
loc = torch.randint(0, 10, (tot, a.envs), device=dev, dtype=torch.uint8, generator=g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(tot):
    if i == a.warmup:
        e0.record()
    if a.what == "pipeline":
        env.afterstates(packed=True); env.move(rot[i], loc[i]); env.reset(done_only=True)
    elif a.what == "afterstates_f32":
        env.afterstates(f32=True, u8=False); env.move(rot[i], loc[i]); env.reset(done_only=True)
    elif a.what == "fused":
        env.step_observe(rot[i], loc[i], packed=True)
    elif a.what == "fused_distinct":
        env.step_observe_distinct(rot[i], loc[i])
    elif a.what == "afterstates_distinct":
        env.afterstates_distinct(); env.move(rot[i], loc[i]); env.reset(done_only=True)
    elif a.what == "rollout_random":
        env.rollout_random(32)
    elif a.what == "rollout_greedy":
        env.rollout_greedy(8, [760, -360, -180, -510, 100000, -100000])
e1.record()
torch.cuda.synchronize()
if a.what == "pipeline":
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record(); env.afterstates(packed=True); evs[1].record(); env.move(rot[0], loc[0]); evs[2].record(); env.reset(done_only=True); evs[3].record()
    torch.cuda.synchronize()
    print("  kernels ms: afterstates %.4f step %.4f reset %.4f" % tuple(evs[j].elapsed_time(evs[j + 1]) for j in range(3)))
print(f"{a.what}: {e0.elapsed_time(e1) / a.steps:.4f} ms per pass, launches={tp.launch_count()}")
