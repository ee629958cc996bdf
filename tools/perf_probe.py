"""Developer probe (not a pytest): time one scene family at a given size on the GPU and print the
stage breakdown + counters.  python tests/perf_probe.py {pile|column|boxes} N [settle] [steps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb  # noqa: E402
from ferox_b200 import scenes  # noqa: E402

kind, n = sys.argv[1], int(sys.argv[2])
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 300
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
L = fb.lib()
t0 = time.time()
spec = {"pile": lambda: scenes.mixed_pile(n), "column": lambda: scenes.column_collapse(n, cols=max(64, int((n / 4) ** 0.5))),
        "boxes": lambda: scenes.box_stack(max(1, n // 8), 8)}[kind]()
sc = scenes.Scene(L, spec, prime=False)
L.frStepWorld(sc.world, scenes.DT)
print("built %s n=%d in %.1fs" % (spec.name, spec.n, time.time() - t0))
t0 = time.time()
L.frxStepWorldN(sc.world, scenes.DT, settle)
L.frxSynchronizeWorld(sc.world)
print("settle %d steps: %.3f s wall (%.3f ms/step)" % (settle, time.time() - t0, 1e3 * (time.time() - t0) / max(1, settle)))
fb.check_error(L)
L.frxSetProfiling(sc.world, 1)
ms = np.zeros(steps, np.float32)
L.frxTimedSteps(sc.world, scenes.DT, steps, 1, ms)
st = (C.c_double * 5)()
k = L.frxGetStageTimes(sc.world, st)
c = fb.step_counters(L, sc.world)
fb.check_error(L)
print("ms/step %.4f (min %.4f)  body-steps/s %.3e" % (ms.mean(), ms.min(), spec.n / (ms.mean() * 1e-3)))
print("stages ms:", {n_: round(st[i] / k, 4) for i, n_ in enumerate(["broad", "narrow", "cache", "solve", "integrate"])})
print("counters:", c)
M, Q, N, K, Cn = c["manifolds"], c["contacts"], c["bodies"], c["cellEntries"], c["candidates"]
solve_b = 964.0 * M + 376.0 * Q
print("solve alg GB/s %.1f   step alg bytes/body %.0f" % (solve_b / (st[3] / k * 1e-3) / 1e9, (N * 168 + 104 * K + Cn * 72 + 1060 * M + 376 * Q) / N))
s = fb.body_states(L, sc.world)
print("finite:", bool(np.isfinite(s["pos"]).all()), "max|v| %.2f" % np.abs(s["vel"]).max())

# ---- end-to-end call breakdown (host buffers every step) ----
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
nb = L.frGetBodyCountInWorld(sc.world)
force = np.zeros((nb, 2), np.float32); force[:, 0] = 0.01
t_apply = t_step = t_get = 0.0
for _ in range(30):
    t0 = time.perf_counter(); L.frxApplyForces(sc.world, ptr(force), None, nb); t1 = time.perf_counter()
    L.frStepWorld(sc.world, scenes.DT); t2 = time.perf_counter()
    L.frxGetBodyStates(sc.world, ptr(s["pos"]), ptr(s["angle"]), ptr(s["sincos"]), ptr(s["vel"]), ptr(s["omega"]), ptr(s["aabb"])); t3 = time.perf_counter()
    t_apply += t1 - t0; t_step += t2 - t1; t_get += t3 - t2
print("e2e per step: apply %.3f ms, step(launch) %.3f ms, get(sync+D2H+copy) %.3f ms" % (t_apply / 30 * 1e3, t_step / 30 * 1e3, t_get / 30 * 1e3))
