"""BASELINE configs[3]: the DQN afterstate-value loop (reference hyper-parameters) driving 65 536 GPU envs, on both rollout
paths: the fused tensor-core ranking kernel over the distinct placements (default) and the PyTorch forward over the 40-slot grid."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tetris_piclim as tp  # noqa: E402
from importlib import import_module  # noqa: E402
train = import_module(tp.__name__ + ".train")
pool = tp.concat_pools(tp.synthetic_pool(4096, seed=0, M=30), tp.carve_pool(4096, 10, 30, seed0=0, with_solutions=False))
out = {"config": "65536 envs, DQN afterstate-value loop, model/train.py constants, " + os.environ.get("OPTIM", "1") + " optimiser steps per iteration"}
iters = int(os.environ.get("ITERS", "100"))
OPT = int(os.environ.get("OPTIM", "1"))
for name, vk in (("value_kernel", True), ("pytorch_forward", False)):
    train.train(num_envs=65536, iterations=5, config_pool=pool, optim_steps_per_iter=OPT, value_kernel=vk)      # warm-up: cuBLAS / allocator start-up
    net, st = train.train(num_envs=65536, iterations=iters, config_pool=pool, optim_steps_per_iter=OPT, value_kernel=vk)
    out[name] = {"iterations": st.env_steps // 65536, "env_only_steps_per_s": st.env_steps_per_s, "end_to_end_steps_per_s": st.e2e_steps_per_s,
                 "ms_per_iteration": st.total_seconds * 1e3 / iters, "optim_steps": st.optim_steps, "loss": st.loss, "episodes": st.episodes,
                 "wins": st.wins}
print(json.dumps(out))
