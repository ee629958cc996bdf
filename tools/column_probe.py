"""dev probe: watch the config-3 block collapse for runaway bodies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ferox_b200 as fb
from ferox_b200 import scenes
import bench
L = fb.lib(); L.frxSetDevice(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 96
if os.environ.get("STOREYS"):
    spec, _, _ = scenes.strip_pile(n, 0, 1, storeys=int(os.environ["STOREYS"]), max_rows=rows, pitch=float(os.environ.get("PITCH", "1.2")), circle_fraction=1.0)
else:
    spec = scenes.column_collapse(n, cols=max(64, n // rows), pitch=float(os.environ.get('PITCH', '1.2')), jitter=float(os.environ.get('JIT', '0.2')), floor_width=32768.0,
                                  friction=float(os.environ.get('FRIC', '0.5')), stagger=os.environ.get('STAG', '0') == '1',
                                  row_pitch=float(os.environ['ROWP']) if 'ROWP' in os.environ else None)
sc = bench.build_bulk_world(L, spec)
L.frStepWorld(sc.world, scenes.DT)
for k in range(14):
    L.frxStepWorldN(sc.world, scenes.DT, 50)
    st = fb.body_states(L, sc.world)
    c = fb.step_counters(L, sc.world)
    v = np.hypot(st["vel"][:, 0], st["vel"][:, 1])
    bad = ~np.isfinite(st["pos"]).all(1)
    i = int(np.nanargmax(v))
    print((k + 1) * 50, "max|v| %.2f at body %d pos %s  nan %d  x[%.1f,%.1f] y[%.1f,%.1f] max|w| %.1f M %d colors %d ovf %d" % (
        v[i], i, st["pos"][i], bad.sum(), np.nanmin(st["pos"][:, 0]), np.nanmax(st["pos"][:, 0]), np.nanmin(st["pos"][:, 1]),
        np.nanmax(st["pos"][:, 1]), np.nanmax(np.abs(st["omega"])), c["manifolds"], c["colors"], c["overflow"]), flush=True)
    if c["overflow"]:
        break
print(L.frxGetLastErrorString())
