"""Where does a DQN iteration go?  torch profiler over the loop at 65 536 envs (GPU time by kernel, CPU time by op).
    python scripts/dqn_profile.py [0|1]     (1 = value-kernel path, the default; 0 = PyTorch forward over the 40-slot grid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tetris_piclim as tp
from importlib import import_module
train = import_module(tp.__name__ + ".train")
vk = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
pool = tp.synthetic_pool(4096, seed=0, M=30)
train.train(num_envs=65536, iterations=10, config_pool=pool, value_kernel=vk)      # warm-up
from torch.profiler import profile, ProfilerActivity
N = 20
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    net, st = train.train(num_envs=65536, iterations=N, config_pool=pool, value_kernel=vk)
print(f"value_kernel={vk}: {st.total_seconds * 1e3 / N:.3f} ms per iteration")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=12, max_name_column_width=60))
