"""Developer probe: where does a 2-strip run differ from the undivided world?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb
from ferox_b200 import capi, scenes
L = fb.lib()
DT = scenes.DT
spec = scenes.mixed_pile(3000, width=100.0)
sc = scenes.Scene(L, spec, prime=False); L.frStepWorld(sc.world, DT); L.frxStepWorldN(sc.world, DT, 600)
want = fb.body_states(L, sc.world); wc = fb.step_counters(L, sc.world)
ss = scenes.StripScene(L, spec, cuts=[50.0], margin=5.0, ghost_cap=1024)
ss.step(601)
got = ss.gather_states()
dyn = spec.body_type != capi.BODY_STATIC
x = want["pos"][:, 0]
for lo in range(0, 100, 10):
    m = dyn & (x >= lo) & (x < lo + 10)
    print("x in [%d,%d): n=%d mean y single %.3f strips %.3f  |v| single %.3f strips %.3f" % (lo, lo + 10, m.sum(), want["pos"][m, 1].mean(), got["pos"][m, 1].mean(),
          np.linalg.norm(want["vel"][m], axis=1).mean(), np.linalg.norm(got["vel"][m], axis=1).mean()))
print("single counters", wc)
for w in ss.worlds:
    print("strip counters", fb.step_counters(L, w))
    m = fb.manifolds(L, w); d = np.concatenate([m["contacts"]["depth"][:, 0], m["contacts"]["depth"][m["count"] > 1, 1]])
    print("  mean depth %.4f max %.3f" % (d.mean(), d.max()))
m = fb.manifolds(L, sc.world); d = np.concatenate([m["contacts"]["depth"][:, 0], m["contacts"]["depth"][m["count"] > 1, 1]])
print("single mean depth %.4f max %.3f" % (d.mean(), d.max()))
w = ss.worlds[0]
m = fb.manifolds(L, w)
deep = m[m["contacts"]["depth"][:, 0] > 1.0]
print("deep manifolds", len(deep), "of", len(m))
n_owned = L.frGetBodyCountInWorld(w)
print("n_owned", n_owned, "sample", [(int(a), int(b), int(c), float(d[0]["depth"])) for a, b, c, d in zip(deep["a"][:12], deep["b"][:12], deep["count"][:12], deep["contacts"][:12])])
import collections
print("a histogram", collections.Counter(deep["a"].tolist()).most_common(5), "b>=n_owned:", int((deep["b"] >= n_owned).sum()))
c = fb.candidate_pairs(L, w)
same = c[c[:, 0] == c[:, 1]]
print("candidates", len(c), "self pairs", len(same), same[:5].tolist())
big = c[(c[:, 0] == 0)]
print("pairs with body 0:", len(big), "partners >= n_owned:", int((big[:, 1] >= n_owned).sum()), big[-5:].tolist())
st = fb.body_states(L, w)
print("floor aabb", st["aabb"][0], "walls", st["aabb"][1], st["aabb"][2])
