import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb
from ferox_b200 import capi, scenes
L = fb.lib(); DT = scenes.DT
spec = scenes.SceneSpec(name="two", cell=2.0)
spec.shapes = [scenes.ShapeSpec("rect", (40.0, 2.0)), scenes.ShapeSpec("circle", (0.6,))]
scenes._finish(spec, [capi.BODY_STATIC, capi.BODY_DYNAMIC, capi.BODY_DYNAMIC, capi.BODY_DYNAMIC],
               [0, 1, 1, 1], [(50.0, 1.0), (49.45, -0.6), (50.55, -0.6), (50.0, -1.6)])
ss = scenes.StripScene(L, spec, cuts=[50.0], margin=4.0, ghost_cap=16)
sc = scenes.Scene(L, spec, prime=False); L.frStepWorld(sc.world, DT)
ss.step(1)
for s in range(6):
    ss.step(1); L.frStepWorld(sc.world, DT)
    got = ss.gather_states(); want = fb.body_states(L, sc.world)
    print("step", s, "strips pos", got["pos"][1:].round(4).tolist(), "vel", got["vel"][1:].round(4).tolist())
    print("       single pos", want["pos"][1:].round(4).tolist(), "vel", want["vel"][1:].round(4).tolist())
    for k, w in enumerate(ss.worlds):
        m = fb.manifolds(L, w)
        print("   strip", k, fb.step_counters(L, w)["candidates"], [(int(a), int(b), int(c), float(np.round(d[0]['depth'], 4)), float(np.round(d[0]['normalScalar'], 4))) for a, b, c, d in zip(m["a"], m["b"], m["count"], m["contacts"])])
    m = fb.manifolds(L, sc.world)
    print("   single", [(int(a), int(b), int(c), float(np.round(d[0]['depth'], 4)), float(np.round(d[0]['normalScalar'], 4))) for a, b, c, d in zip(m["a"], m["b"], m["count"], m["contacts"])])
fb.check_error(L)
