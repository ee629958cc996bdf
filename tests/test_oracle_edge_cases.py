"""CPU: pins the oracle against the UNMODIFIED reference on the corner cases the GPU edge-case tests
exercise (tests/test_gpu_edge_cases.py): material mixing incl. negative values, restitution, static /
kinematic / flagged bodies, negative gravity scale, coincident circles and deep overlaps, 3..8-gons in
rotated stacks.  Bit-exact body state and contact cache (array order), every few steps."""
import math

import numpy as np
import pytest

from ferox_b200 import capi, scenes
from ferox_b200.scenes import SceneSpec, ShapeSpec, _finish
from oracle import orc
from tests.helpers import assert_caches_equal, assert_states_equal, records_to_entries

D, S, K = capi.BODY_DYNAMIC, capi.BODY_STATIC, capi.BODY_KINEMATIC


def poly(n, rad, phase=0.0, **kw):
    return ShapeSpec("poly", tuple((float(np.float32(rad * math.cos(phase + 2 * math.pi * k / n))),
                                    float(np.float32(rad * math.sin(phase + 2 * math.pi * k / n)))) for k in range(n)), **kw)


def materials():
    spec = SceneSpec(name="edge_materials", cell=2.0)
    spec.shapes = [ShapeSpec("rect", (30.0, 1.0), friction=0.9, restitution=0.6),
                   ShapeSpec("circle", (0.5,), friction=0.1, restitution=0.8),
                   ShapeSpec("rect", (1.2, 0.8), friction=0.4, restitution=0.0),
                   ShapeSpec("circle", (0.6,), friction=-1.0, restitution=-0.5)]      # clamped to 0 (src/world.c:401-402)
    types, shapes, pos = [S], [0], [(0.0, 3.0)]
    for i, s in enumerate([1, 2, 3, 1, 2, 3, 1, 2]):
        types.append(D); shapes.append(s); pos.append((-4.0 + 1.3 * i, -1.0 - 0.7 * (i % 4)))
    return _finish(spec, types, shapes, pos), None


def kinds_and_flags():
    spec = SceneSpec(name="edge_kinds", cell=2.0)
    spec.shapes = [ShapeSpec("rect", (30.0, 1.0)), ShapeSpec("rect", (1.0, 1.0), friction=0.75), ShapeSpec("circle", (0.5,))]
    types, shapes, pos = [S, K], [0, 1], [(0.0, 3.0), (-6.0, 1.5)]
    vel, omega = [(0.0, 0.0), (1.5, 0.0)], [0.0, 0.7]                                  # the kinematic body moves and spins
    for i in range(8):
        types.append(D); shapes.append(1 if i % 2 else 2); pos.append((-4.0 + 1.1 * i, 1.0 - 0.3 * (i % 3)))
        vel.append((0.0, 0.0)); omega.append(0.0)
    spec = _finish(spec, types, shapes, pos, vel=vel, omega=omega)

    def tweak(lib, bodies):
        lib.frSetBodyFlags(bodies[4], capi.FLAG_INFINITE_INERTIA)
        lib.frSetBodyFlags(bodies[5], capi.FLAG_INFINITE_MASS)
        lib.frSetBodyGravityScale(bodies[6], -0.5)
    return spec, tweak


def coincident():
    spec = SceneSpec(name="edge_coincident", cell=2.0)
    spec.shapes = [ShapeSpec("circle", (0.5,)), ShapeSpec("circle", (1.0,))]
    return _finish(spec, [D, D, D, S], [0, 0, 1, 1], [(0.0, 0.0), (0.0, 0.0), (0.1, 0.05), (0.0, 0.0)]), None


def polygons():
    spec = SceneSpec(name="edge_polygons", cell=2.0)
    spec.shapes = [ShapeSpec("rect", (40.0, 1.0)), poly(8, 0.7), poly(3, 0.9, 0.3), poly(5, 0.65), poly(8, 0.6, 0.2), poly(4, 0.7, 0.785)]
    types, shapes, pos, angle = [S], [0], [(0.0, 3.0)], [0.0]
    for i in range(30):
        types.append(D); shapes.append(1 + i % 5); pos.append((-8.0 + 1.45 * (i % 12), 1.2 - 1.5 * (i // 12))); angle.append(0.37 * i)
    return _finish(spec, types, shapes, pos, angle), None


CASES = [("materials", materials, 160), ("kinds_and_flags", kinds_and_flags, 150), ("coincident", coincident, 60), ("polygons", polygons, 200)]


@pytest.mark.parametrize("name,make,steps", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_on_edge_cases(ref, name, make, steps):
    spec, tweak = make()
    sc = scenes.Scene(ref, spec)
    if tweak is not None:
        tweak(ref, sc.bodies)
    ow = orc.OracleWorld.from_scene(sc)
    rec = scenes.ManifoldRecorder(sc)
    for s in range(steps):
        sc.step(1)
        ow.step(1)
        got = rec.take()
        if s % 7 and s != steps - 1:
            continue
        assert_states_equal(sc.snapshot(), ow.state(), "%s step %d" % (name, s))
        oc = ow.cache()
        assert_caches_equal(records_to_entries(got), oc[oc["count"] > 0], "%s step %d" % (name, s))
    rec.remove()
    sc.release()
    ow.close()
