"""Debug aid: first afterstate mismatches between the GPU and the C oracle, with the env they belong to."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle
from tests import parity_cases as pc
from tests.engines import GpuEngine

eng = GpuEngine()
n, L, M, seed = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 10, 30, 5
rng = np.random.default_rng(seed)
rows = pc.adversarial_boards(rng, n)
npieces = np.full(n, 31, np.uint8)
pieces = rng.integers(0, 7, (n, pc.P)).astype(np.uint8)
s = eng.pack(rows, pieces, npieces)
ost = pc.oracle_state(rows, pieces, npieces)
feats, flags = eng.afterstates(s, L, M)[:2]
of, ofl, _ = c_oracle.afterstates_batch(ost, L, M, nthreads=8)
feats = feats.reshape(n, 40, 4); flags = flags.reshape(n, 40)
bad = np.argwhere((feats != of).any(axis=2) | (flags != ofl))
print("mismatching slots:", len(bad), "of", n * 40)
import collections
print("by slot:", sorted(collections.Counter(int(b[1]) for b in bad).items()))
print("by piece:", sorted(collections.Counter(int(pieces[b[0], 0]) for b in bad).items()))
for e, sl in bad[:12]:
    cols = [(rows[e] >> c) & 1 for c in range(10)]
    H = [20 - int(np.argmax(cols[c])) if cols[c].any() else 0 for c in range(10)]
    print(f"env {e} piece {pieces[e,0]} slot {sl} got {feats[e,sl].tolist()}/{flags[e,sl]} exp {of[e,sl].tolist()}/{ofl[e,sl]} H={H}")
pk = eng.afterstates_packed(s, L, M).reshape(n, 40, 4)
exp = of.copy(); exp[:, :, 0] |= (ofl << 3)
print("packed mismatches:", int((pk != exp).any(axis=2).sum()))
