"""dev probe: SHA-1 of the state after N coloured steps of a seeded pile (compare library builds: FEROX_B200_LIB=...)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ferox_b200 as fb
from ferox_b200 import scenes
L = fb.lib(); L.frxSetDevice(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sc = scenes.Scene(L, scenes.mixed_pile(n), prime=False)
L.frStepWorld(sc.world, scenes.DT); L.frxStepWorldN(sc.world, scenes.DT, steps)
st = fb.body_states(L, sc.world); fb.check_error(L)
print("HASH", n, steps, hashlib.sha1(b"".join(st[k].tobytes() for k in ("pos", "angle", "vel", "omega"))).hexdigest(), fb.step_counters(L, sc.world))
