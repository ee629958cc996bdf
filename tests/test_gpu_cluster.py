"""The cluster solver (fx_solve.cu k_sweep_cluster: one thread-block cluster, velocities in distributed shared
memory, cluster barriers, bulk-async prefetch) against the grid solver (k_solve_colored): same colours, same phase
order, same per-contact arithmetic, so the two must agree BIT FOR BIT -- state, contact cache and counters -- after
hundreds of steps, for both cluster sizes.  FEROX_CLUSTER is read once per process, hence one subprocess per variant."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = """
import hashlib, sys, numpy as np, ferox_b200 as fb
from ferox_b200 import scenes
L = fb.lib(); L.frxSetDevice(0)
kind, n, width, steps = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
spec = scenes.box_stack(96, 12) if kind == 'boxes' else scenes.mixed_pile(n, width=width if width > 0 else None)
sc = scenes.Scene(L, spec, prime=False)
L.frStepWorld(sc.world, scenes.DT)
for chunk in range(3):
    L.frxStepWorldN(sc.world, scenes.DT, steps // 3)
    st = fb.body_states(L, sc.world); fb.check_error(L)
    m = fb.manifolds(L, sc.world)
    c = fb.step_counters(L, sc.world)
    h = hashlib.sha1(b''.join(st[k].tobytes() for k in ('pos', 'angle', 'vel', 'omega')) + m.tobytes()).hexdigest()
    print('HASH', chunk, h, c['manifolds'], c['contacts'], c['colors'], c['overflow'], L.frxGetWorldSolverKernel(sc.world))
"""

CASES = [("pile", 6000, 150.0, 300), ("pile", 20000, 0.0, 240), ("boxes", 0, 0.0, 300), ("pile", 65536, 0.0, 150)]


@pytest.mark.parametrize("kind,n,width,steps", CASES, ids=["pile6k", "pile20k", "boxes1k", "config2_65k"])
def test_cluster_solver_equals_grid_solver_bit_for_bit(kind, n, width, steps):
    out = {}
    for variant in ("0", "16", "8"):
        env = dict(os.environ, FEROX_CLUSTER=variant, FEROX_LOCAL_SOLVE="0", PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-c", CODE, kind, str(n), str(width), str(steps)], capture_output=True, text=True, timeout=600,
                           cwd=ROOT, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[variant] = [l.split()[1:] for l in r.stdout.splitlines() if l.startswith("HASH")]
        assert len(out[variant]) == 3, r.stdout[-1000:]
    assert all(l[-1] == "0" for l in out["0"]), out["0"]
    if not all(l[-1] == "16" for l in out["16"]):
        pytest.skip("a 16-CTA cluster is not launchable on this device: %r" % (out["16"],))
    assert all(l[-1] == "8" for l in out["8"]), out["8"]
    strip = lambda rows: [l[:-1] for l in rows]
    assert strip(out["0"]) == strip(out["16"]) == strip(out["8"]), out
    assert int(out["0"][-1][2]) > 1000 and out["0"][-1][5] == "0"
