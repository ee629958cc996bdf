"""dev probe: manifolds per colour of the settled config-2 pile (how much of a solve's ~88 phases is spent on sparse tail colours)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ferox_b200 as fb
from ferox_b200 import scenes
L = fb.lib(); L.frxSetDevice(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sc = scenes.Scene(L, scenes.mixed_pile(n), prime=False)
L.frStepWorld(sc.world, scenes.DT)
L.frxStepWorldN(sc.world, scenes.DT, 600)
h = np.zeros(64, np.int32)
k = L.frxGetColorHistogram(sc.world, h)
print("colours", k, "manifolds per colour", h[:k].tolist(), fb.step_counters(L, sc.world))
