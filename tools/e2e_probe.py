"""dev probe: where the end-to-end step time goes (host calls around a 65k-body world)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ferox_b200 as fb
from ferox_b200 import scenes
L = fb.lib(); L.frxSetDevice(0)
spec = scenes.mixed_pile(65536)
sc = scenes.Scene(L, spec, prime=False)
dt = scenes.DT
L.frStepWorld(sc.world, dt); L.frxStepWorldN(sc.world, dt, 300); L.frxSynchronizeWorld(sc.world)
n = L.frGetBodyCountInWorld(sc.world)
force = np.zeros((n, 2), np.float32); force[:, 0] = 0.01
st = fb.body_states(L, sc.world)
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
T = np.zeros(5)
K = 100
for _ in range(K):
    t0 = time.perf_counter()
    L.frxApplyForces(sc.world, ptr(force), None, n)
    t1 = time.perf_counter()
    L.frStepWorld(sc.world, dt)
    t2 = time.perf_counter()
    L.frxSynchronizeWorld(sc.world)
    t3 = time.perf_counter()
    L.frGetBodyPosition(sc.bodies[5])          # triggers the mirror pull alone
    t4 = time.perf_counter()
    L.frxGetBodyStates(sc.world, ptr(st["pos"]), ptr(st["angle"]), ptr(st["sincos"]), ptr(st["vel"]), ptr(st["omega"]), ptr(st["aabb"]))
    t5 = time.perf_counter()
    T += (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)
print("per step, ms: apply %.3f  step(call) %.3f  sync %.3f  pull %.3f  scatter %.3f  total %.3f" % (*(T / K * 1e3), T.sum() / K * 1e3))
