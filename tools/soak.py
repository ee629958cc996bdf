"""dev probe: long run of the headline scene -- counters must stay sane (no overflow, colours tight)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ferox_b200 as fb
from ferox_b200 import scenes
L = fb.lib(); L.frxSetDevice(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
spec = scenes.mixed_pile(n)
sc = scenes.Scene(L, spec, prime=False)
L.frStepWorld(sc.world, scenes.DT)
for k in range(10):
    L.frxStepWorldN(sc.world, scenes.DT, 1000)
    st = fb.body_states(L, sc.world); c = fb.step_counters(L, sc.world)
    v = np.hypot(st["vel"][:, 0], st["vel"][:, 1])
    print((k + 1) * 1000, "mean|v| %.3f max|v| %.1f max|w| %.1f y[%.1f,%.1f] M %d colors %d rounds %d ovf %d" % (
        v.mean(), v.max(), np.abs(st["omega"]).max(), st["pos"][:, 1].min(), st["pos"][:, 1].max(), c["manifolds"], c["colors"], c["colorRounds"], c["overflow"]), flush=True)
print(L.frxGetLastErrorString())
