"""One-off randomized sweep on the GPU: the fused step, the stand-alone kernels and the rollouts against the oracle for
random (n, L, M, seed) -- sizes around warp/tile boundaries, tiny and large L/M."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.engines import GpuEngine
from tests import parity_cases as pc
from tests.test_gpu_parity import _pool

g = GpuEngine()
pool = _pool(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    n = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4097, 37_889, 40_000, 65_537]))
    L, M, seed = int(rng.integers(1, 17)), int(rng.integers(1, 42)), int(rng.integers(0, 1 << 30))
    pc.case_fused_step_observe(g, pool, n=n, steps=int(rng.integers(3, 30)), L=L, M=M, seed=seed, env_base=int(rng.integers(0, 1 << 40)))
    pc.case_afterstates_vs_oracle(g, n, L, M, seed)
    pc.case_random_moves(g, n, min(M + 3, 12), L, M, seed)
    if k % 4 == 0:
        m = min(n, 3000)
        pc.case_rollout(g, pool, m, 40, L, M, seed=seed & 0xFFFF, env_base=77, chunks=(0.5, 0.5))
        pc.case_rollout(g, pool, m, 30, L, M, seed=seed & 0xFFFF, env_base=5, weights=[760, -360, -180, -510, 100000, -100000], chunks=(0.4, 0.6))
        pc.case_rollout(g, pool, min(m, 500), 20, L, M, seed=seed & 0xFFFF, env_base=5, weights=[70000, -40000, -180, -510, 100000, -100000], chunks=(1.0,))
    print("ok", k, n, L, M, flush=True)
print("fuzz done")
