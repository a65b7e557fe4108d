"""Small end-to-end case for compute-sanitizer (memcheck): every kernel once or twice on ragged sizes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tetris_piclim as tp

pool = tp.concat_pools(tp.synthetic_pool(300, seed=0, M=30), tp.carve_pool(100, 10, 30, seed0=0, with_solutions=False))
for n in (1, 33, 1000, 40_000):                       # split kernel (small n) and tile kernels (large n), ragged tails
    env = tp.BatchedTetris(n, 10, 30, seed=1, config_pool=pool, env_base=7)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    for t in range(3):
        rot = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8, generator=g)
        loc = torch.randint(0, 10, (n,), device="cuda", dtype=torch.uint8, generator=g)
        env.afterstates(); env.afterstates(packed=True); env.afterstates(f32=True)
        env.move(rot, loc); env.reset(done_only=True)
        env.step_observe(rot, loc, packed=True); env.step_observe(rot, loc, packed=False, f32=True)
    env.rollout_random(20); env.rollout_greedy(5, [760, -360, -180, -510, 100000, -100000])
    env.get_state(); env.fields(queue=True); env.gen_pieces(31)
    env.reset(mask=np.ones(n, np.uint8)); env.reset(idx=np.zeros(n, np.int32))
    torch.cuda.synchronize()
h = tp.HostBatchedTetris(5000, 10, 30, seed=3, config_pool=pool)
h.reset(); h.move(np.zeros(5000), np.zeros(5000)); h.afterstates(); h.fields(); h.close()
g1 = tp.Tetris(10, 30, warm_reset=False, config_pool=pool); g1.move(1, 3); g1.afterstates(); g1.terminate()
print("sanitize case done")
