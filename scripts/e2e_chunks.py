"""How many chunks should the pipelined host step use?  End-to-end time of tpl_env_step_observe_distinct / tpl_env_step_observe at
2^20 envs for TPL_ENV_CHUNKS = 1..6 (one process per setting: the variable is read when the handle is created)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import time
    import numpy as np
    import tetris_piclim as tp
    n = 1 << 20
    pool = tp.synthetic_pool(4096, seed=0, M=30)
    h = tp.HostBatchedTetris(n, 10, 30, seed=0, config_pool=pool); h.reset()
    cap = h.distinct_capacity()
    pin = {k: tp.PinnedArray(s, d) for k, (s, d) in dict(rot=((n,), np.uint8), loc=((n,), np.uint8), dl=((n,), np.int8), fl=((n,), np.uint8),
           st=((n,), np.int8), feats=((40, n, 4), np.uint8), rows=((cap,), np.uint32), runs=((n,), np.uint32)).items()}
    rng = np.random.default_rng(0)
    pin["rot"].array[:] = rng.integers(0, 4, n); pin["loc"].array[:] = rng.integers(0, 10, n)
    b5 = [pin[k].array for k in ("rot", "loc", "dl", "fl", "st")]
    out = {"chunks": h.chunks()}
    for name, call in (("distinct", lambda: h.step_observe_distinct(*b5, pin["rows"].array, pin["runs"].array)),
                       ("slots40", lambda: h.step_observe(*b5, pin["feats"].array, None))):
        for _ in range(3): call()
        t0 = time.perf_counter()
        for _ in range(10): call()
        out[name + "_ms"] = (time.perf_counter() - t0) * 100
    print(json.dumps(out))
else:
    for c in (1, 2, 3, 4, 6, 8):
        env = dict(os.environ, TPL_ENV_CHUNKS=str(c))
        print(subprocess.run([sys.executable, __file__, "x"], env=env, capture_output=True, text=True).stdout.strip())
