"""Short driver for ncu: the value-net ranking kernel over the distinct placements of 65 536 envs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tetris_piclim as tp
from importlib import import_module
vkm, model = import_module(tp.__name__ + ".value_kernel"), import_module(tp.__name__ + ".model")
env = tp.BatchedTetris(65536, 10, 30, seed=0, config_pool=tp.synthetic_pool(4096, seed=0, M=30))
env.reset(); env.rollout_random(6); env.reset(done_only=True)
rows, runs, used = env.afterstates_distinct()
torch.manual_seed(0)
vk = vkm.ValueKernel(model.ValueNet().cuda())
vals = vk.values(rows, used.reshape(1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    vk.values(rows, used.reshape(1), out=vals)
e1.record(); torch.cuda.synchronize()
print(f"value_rows: {e0.elapsed_time(e1) / 5:.4f} ms for {int(used)} rows")
