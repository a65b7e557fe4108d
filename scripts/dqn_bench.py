"""BASELINE configs[3]: the DQN afterstate-value loop (reference hyper-parameters) driving 65 536 GPU envs.
Prints one JSON line: the loop on the value-kernel path (one and four optimiser steps per env step), the same loop with the
forward pass in PyTorch over the 40-slot grid, and the ranking kernel alone.  bench.py runs this in a fresh process for its
`dqn_loop_65536` extra (CUDA-graph capture empties the caching allocator: inside the bench process that costs seconds)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tetris_piclim as tp  # noqa: E402
from importlib import import_module  # noqa: E402
train = import_module(tp.__name__ + ".train")
vk_mod = import_module(tp.__name__ + ".value_kernel")
N = 65536
pool = tp.concat_pools(tp.synthetic_pool(4096, seed=0, M=30), tp.carve_pool(4096, 10, 30, seed0=0, with_solutions=False))
iters = int(os.environ.get("ITERS", "100"))


def run(value_kernel, optim):
    train.train(num_envs=N, iterations=6, config_pool=pool, optim_steps_per_iter=optim, value_kernel=value_kernel)      # warm-up
    net, st = train.train(num_envs=N, iterations=iters, config_pool=pool, optim_steps_per_iter=optim, value_kernel=value_kernel)
    return net, {"iterations": iters, "optimiser_steps_per_iteration": optim, "env_steps_per_s_env_only": st.env_steps_per_s,
                 "env_steps_per_s_end_to_end": st.e2e_steps_per_s, "ms_per_iteration": st.total_seconds * 1e3 / iters,
                 "optim_steps": st.optim_steps, "loss": st.loss, "episodes": st.episodes}


out = {"config": "65536 envs, DQN afterstate-value loop, model/train.py constants (batch 128); the timed runs include their own 3 eager "
                 "iterations and the CUDA-graph capture of the optimiser block"}
net, out["value_kernel"] = run(True, 1)
out["value_kernel"]["note"] = ("rollout half on the library's kernels (tpl_value_rows on the tensor cores over the distinct placements, "
                               "tpl_select_action, tpl_step_observe_distinct, tpl_replay_push), optimiser half in PyTorch on a second stream: "
                               "one optimisation step per env step, the structure of the DQN loop model/train.py's constants come from")
_, out["value_kernel_4_optimiser_steps"] = run(True, 4)
_, out["pytorch_forward_40_slots"] = run(False, 1)
out["pytorch_forward_40_slots"]["note"] = "same loop with the ranking forward in PyTorch (rank_bf16 over the 40-slot grid, round 1's path)"
# the ranking forward alone: value net over the distinct placements of 65 536 envs
env = tp.BatchedTetris(N, 10, 30, seed=0, config_pool=pool)
env.reset(); env.rollout_random(6); env.reset(done_only=True)
rows, runs, used = env.afterstates_distinct()
vk = vk_mod.ValueKernel(net)
vals = vk.values(rows, used.reshape(1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    vk.values(rows, used.reshape(1), out=vals)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
R = int(used)
out["value_rows_kernel"] = {"ms": ms, "rows": R, "rows_per_s": R / (ms * 1e-3),
                            "TFLOPs_algorithmic": 2.0 * R * (4 * 128 + 3 * 128 * 128 + 128) / (ms * 1e-3) / 1e12,
                            "note": "tpl_value_rows alone: tcgen05 128x128x16 bf16 MMAs, weights in shared memory, accumulators and activations in "
                                    "tensor memory; FLOPs = 2 * rows * (4*128 + 3*128*128 + 128), the net's own"}
print(json.dumps(out))
