import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb
from ferox_b200 import capi, scenes
L = fb.lib(); DT = scenes.DT
spec = scenes.mixed_pile(3000, width=100.0)
ss = scenes.StripScene(L, spec, cuts=[50.0], margin=5.0, ghost_cap=1024)
ss.step(1)
for s in range(400):
    ss.step(1)
    for k, w in enumerate(ss.worlds):
        m = fb.manifolds(L, w)
        bad = m[(m["a"] == m["b"])]
        if len(bad) and (s < 5 or s % 50 == 0 or len(bad) < 4):
            n_owned = L.frGetBodyCountInWorld(w)
            c = fb.candidate_pairs(L, w)
            hit0 = c[(c[:, 2] == 1) & ((c[:, 0] < 3) | (c[:, 1] >= n_owned))]
            print("step", s, "strip", k, "self manifolds", len(bad), "first", (int(bad["a"][0]), int(bad["count"][0]), float(bad["contacts"][0][0]["depth"])),
                  "n_owned", n_owned, "cands", len(c), "static/ghost hits sample", hit0[:6].tolist())
            break
    else:
        continue
    if s > 3 and len(bad) >= 4:
        pass
fb.check_error(L)
