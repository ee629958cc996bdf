import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb
from ferox_b200 import capi, scenes
L = fb.lib(); DT = scenes.DT
spec = scenes.mixed_pile(3000, width=100.0)
dyn = spec.body_type != capi.BODY_STATIC
for cuts in ([30.0], [70.0], [50.0]):
    ss = scenes.StripScene(L, spec, cuts=cuts, margin=5.0, ghost_cap=1024)
    for chunk in range(6):
        ss.step(100 + (1 if chunk == 0 else 0))
        got = ss.gather_states()
        x0 = spec.pos[:, 0]
        row = []
        for lo in range(0, 100, 10):
            m = dyn & (x0 >= lo) & (x0 < lo + 10)
            row.append("%.1f" % np.linalg.norm(got["vel"][m], axis=1).mean())
        print("cut", cuts, "after", (chunk + 1) * 100, "steps: mean |v| per 10-unit bin:", " ".join(row))
    ss.release()
fb.check_error(L)
