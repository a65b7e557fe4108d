"""Developer aid: per-phase cycle stamps of the value kernel (build with -DTPL_VALUE_TRACE)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tetris_piclim as tp
from importlib import import_module
vkm, model = import_module(tp.__name__ + ".value_kernel"), import_module(tp.__name__ + ".model")
env = tp.BatchedTetris(65536, 10, 30, seed=0, config_pool=tp.synthetic_pool(4096, seed=0, M=30))
env.reset(); env.rollout_random(6); env.reset(done_only=True)
rows, runs, used = env.afterstates_distinct()
vk = vkm.ValueKernel(model.ValueNet().cuda())
vals = torch.zeros(rows.numel(), device="cuda")
vk.values(rows, used.reshape(1), out=vals); vk.values(rows, used.reshape(1), out=vals)
torch.cuda.synchronize()
t = vals[rows.numel() - 4096: rows.numel() - 96].cpu().view(-1, 2)
agg = collections.defaultdict(list)
for tag, d in t.tolist():
    if tag > 0: agg[int(tag)].append(d)
names = {1: "loop top / after issue -> before wait", 2: "mbarrier wait (MMA)", 3: "tcgen05.ld x4 + wait", 4: "cvt + tcgen05.st + wait", 5: "fence + group barrier", 6: "MMA issue + commit (thread 0)", 7: "last wait"}
for k in sorted(agg):
    v = agg[k][20:]
    if v: print(f"{k} {names.get(k)}: n={len(v)} mean {sum(v)/len(v):.0f} cycles, min {min(v):.0f}, max {max(v):.0f}")
