"""Does the completed DQN loop learn?  Trains the afterstate value net on 16 384 GPU envs and prints, per interval, the
win rate / lines per episode of the episodes that ENDED in that interval (epsilon-greedy behaviour policy)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tetris_piclim as tp  # noqa: E402
from importlib import import_module  # noqa: E402
T = import_module(tp.__name__ + ".train")

iters = int(os.environ.get("ITERS", "3000"))
every = int(os.environ.get("EVERY", "250"))
pool = tp.concat_pools(tp.synthetic_pool(4096, seed=0, M=30), tp.carve_pool(4096, 10, 30, seed0=0, with_solutions=False))
log = []
def cb(it, eps, loss, s, prev):
    d = {k: s[k] - prev.get(k, 0) for k in s}
    ep = max(d["episodes"], 1)
    row = {"iter": it, "eps": round(eps, 3), "loss": round(loss, 4), "episodes": d["episodes"], "win_rate": round(d["wins"] / ep, 4),
           "topout_rate": round(d["topouts"] / ep, 4), "lines_per_episode": round(d["lines"] / ep, 3), "moves_per_episode": round(d["moves"] / ep, 2)}
    log.append(row); print(json.dumps(row), flush=True)
net, st = T.train(num_envs=16384, iterations=iters, config_pool=pool, optim_steps_per_iter=8, batch_size=1024, log_every=every, log_fn=cb)
print(json.dumps({"summary": "train.py on 16384 envs", "env_steps": st.env_steps, "seconds": round(st.total_seconds, 1),
                  "first": log[0], "last": log[-1]}))
