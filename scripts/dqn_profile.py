import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tetris_piclim as tp
from importlib import import_module
train = import_module(tp.__name__ + ".train")
pool = tp.synthetic_pool(4096, seed=0, M=30)
train.train(num_envs=65536, iterations=10, config_pool=pool)      # warm-up
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    train.train(num_envs=65536, iterations=10, config_pool=pool)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=10, max_name_column_width=60))
