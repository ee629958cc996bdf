"""GPU: edge cases of the step path, each run on the product (reference-order mode) next to the
UNMODIFIED reference through identical ferox.h calls, compared bit for bit: empty worlds, a world
without a spatial hash, shapeless / static / kinematic bodies, mass flags, restitution, coincident
circles, 8-vertex polygons, shape and gravity edits in mid-run, frClearWorld, capacity."""
import ctypes as C

import numpy as np
import pytest

import ferox_b200 as fb
from ferox_b200 import capi, scenes
from tests.helpers import assert_states_equal

pytestmark = pytest.mark.gpu
DT = scenes.DT


def both(ref, gpu):
    return ((ref, False), (gpu, True))


def new_world(lib, is_prod, cell=2.0, gravity=(0.0, 9.8)):
    w = lib.frCreateWorld(capi.Vector2(*gravity), cell)
    if is_prod:
        lib.frxSetWorldSolverMode(w, fb.SOLVER_SERIAL)
    return w


def run_pair(ref, gpu, build, steps, between=None, cell=2.0):
    """build(lib) -> list of bodies added to a fresh world; returns the two final snapshots."""
    out = []
    for lib, is_prod in both(ref, gpu):
        w = new_world(lib, is_prod, cell)
        bodies = build(lib, w)
        for s in range(steps):
            if between is not None:
                between(lib, w, bodies, s)
            lib.frStepWorld(w, DT)
        out.append((scenes.snapshot(lib, bodies), lib.frGetBodyCountInWorld(w)))
    fb.check_error(gpu)
    assert out[0][1] == out[1][1]
    assert_states_equal(out[0][0], out[1][0], "edge case")
    return out[0][0]


def mat(density=1.0, friction=0.5, restitution=0.0):
    return capi.Material(density, friction, restitution)


def add(lib, w, body):
    assert lib.frAddBodyToWorld(w, body)
    return body


def test_empty_world_and_zero_dt(gpu, ref):
    for lib, is_prod in both(ref, gpu):
        w = new_world(lib, is_prod)
        for _ in range(3):
            lib.frStepWorld(w, DT)
        lib.frStepWorld(w, 0.0)              # dt <= 0: no-op (src/world.c:205)
        lib.frStepWorld(w, -1.0)
        assert lib.frGetBodyCountInWorld(w) == 0
        lib.frReleaseWorld(w)
    fb.check_error(gpu)


def test_world_without_spatial_hash_never_collides(gpu, ref):
    """cellSize <= 0 -> frCreateSpatialHash returns NULL and every hash call is a no-op
    (src/broad_phase.c:58,100,134): bodies fall through each other."""
    def build(lib, w):
        s = lib.frCreateCircle(mat(), 0.5)
        g = lib.frCreateRectangle(mat(), 20.0, 1.0)
        return [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 2), g)),
                add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0, 0), s)),
                add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.2, -0.5), s))]
    st = run_pair(ref, gpu, build, 80, cell=0.0)
    assert st["pos"][1, 1] > 5.0             # fell through the ground


def test_shapeless_static_kinematic_and_flagged_bodies(gpu, ref):
    def build(lib, w):
        c = lib.frCreateCircle(mat(), 0.5)
        r = lib.frCreateRectangle(mat(friction=0.75), 1.0, 1.0)
        g = lib.frCreateRectangle(mat(), 30.0, 1.0)
        bodies = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 3), g))]
        kin = lib.frCreateBodyFromShape(capi.BODY_KINEMATIC, capi.Vector2(-6, 1.5), r)
        lib.frSetBodyVelocity(kin, capi.Vector2(1.5, 0.0))          # kinematic bodies move, are never pushed
        lib.frSetBodyAngularVelocity(kin, 0.7)
        bodies.append(add(lib, w, kin))
        bodies.append(add(lib, w, lib.frCreateBody(capi.BODY_DYNAMIC, capi.Vector2(1, 1))))   # no shape: zero AABB, no mass
        for i in range(8):
            b = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(-4 + 1.1 * i, 1.0 - 0.3 * (i % 3)), r if i % 2 else c)
            if i == 2:
                lib.frSetBodyFlags(b, capi.FLAG_INFINITE_INERTIA)
            if i == 3:
                lib.frSetBodyFlags(b, capi.FLAG_INFINITE_MASS)
            if i == 4:
                lib.frSetBodyGravityScale(b, -0.5)
            bodies.append(add(lib, w, b))
        return bodies
    run_pair(ref, gpu, build, 150)


def test_restitution_and_material_mixing(gpu, ref):
    def build(lib, w):
        g = lib.frCreateRectangle(mat(friction=0.9, restitution=0.6), 30.0, 1.0)
        bouncy = lib.frCreateCircle(mat(friction=0.1, restitution=0.8), 0.5)
        dull = lib.frCreateRectangle(mat(friction=0.4, restitution=0.0), 1.2, 0.8)
        neg = lib.frCreateCircle(mat(friction=-1.0, restitution=-0.5), 0.6)     # clamped to 0 (src/world.c:401-402)
        bodies = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 3), g))]
        for i, s in enumerate([bouncy, dull, neg, bouncy, dull]):
            bodies.append(add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(-3 + 1.6 * i, -1.0 - 0.7 * i), s)))
        return bodies
    run_pair(ref, gpu, build, 200)


def test_coincident_circles_and_deep_overlap(gpu, ref):
    """Coincident centres: direction (0, 1), depth r1 + r2 (src/collision.c:279-280)."""
    def build(lib, w):
        c = lib.frCreateCircle(mat(), 0.5)
        big = lib.frCreateCircle(mat(), 1.0)
        return [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0, 0), c)),
                add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0, 0), c)),
                add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.1, 0.05), big)),
                add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 0), big))]
    run_pair(ref, gpu, build, 60)


def test_octagons_triangles_and_rotated_stacks(gpu, ref):
    def poly(lib, n, rad, phase=0.0):
        v = capi.Vertices()
        v.count = n
        for k in range(n):
            a = phase + 2 * np.pi * k / n
            v.data[k] = capi.Vector2(float(np.float32(rad * np.cos(a))), float(np.float32(rad * np.sin(a))))
        return lib.frCreatePolygon(mat(), C.byref(v))

    def build(lib, w):
        g = lib.frCreateRectangle(mat(), 40.0, 1.0)
        bodies = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 3), g))]
        shapes = [poly(lib, 8, 0.7), poly(lib, 3, 0.9, 0.3), poly(lib, 5, 0.65), poly(lib, 8, 0.6, 0.2), poly(lib, 4, 0.7, 0.785)]
        for i in range(30):
            b = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(-8 + 1.45 * (i % 12), 1.2 - 1.5 * (i // 12)), shapes[i % 5])
            lib.frSetBodyAngle(b, 0.37 * i)
            bodies.append(add(lib, w, b))
        return bodies
    run_pair(ref, gpu, build, 220)


def test_shape_gravity_and_type_edits_in_mid_run(gpu, ref):
    state = {}

    def build(lib, w):
        c = lib.frCreateCircle(mat(), 0.5)
        r = lib.frCreateRectangle(mat(), 1.0, 1.0)
        g = lib.frCreateRectangle(mat(), 30.0, 1.0)
        state[id(lib)] = (c, r)
        bodies = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_STATIC, capi.Vector2(0, 3), g))]
        for i in range(12):
            bodies.append(add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(-5 + 1.2 * i, 1.0 - 0.6 * (i % 2)), c if i % 2 else r)))
        return bodies

    def between(lib, w, bodies, s):
        c, r = state[id(lib)]
        if s == 40:
            lib.frSetCircleRadius(c, 0.7)                       # every body sharing the shape changes size
        if s == 60:
            lib.frSetWorldGravity(w, capi.Vector2(2.0, 6.0))
        if s == 80:
            lib.frSetBodyShape(bodies[3], r)                    # circle -> box, mass recomputed
            lib.frSetBodyShape(bodies[4], None)                 # detach: zero AABB, zero mass
        if s == 100:
            lib.frSetShapeFriction(r, 0.05)                     # only NEW cache entries see it (App. A.5)
            lib.frSetBodyType(bodies[6], capi.BODY_KINEMATIC)
        if s == 120:
            lib.frSetRectangleDimensions(r, 1.4, 0.6)
    run_pair(ref, gpu, build, 180, between)


def test_clear_world_and_reuse(gpu, ref):
    for lib, is_prod in both(ref, gpu):
        w = new_world(lib, is_prod)
        c = lib.frCreateCircle(mat(), 0.5)
        first = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(i * 0.9, 0), c)) for i in range(6)]
        for _ in range(20):
            lib.frStepWorld(w, DT)
        lib.frClearWorld(w)                                     # members dropped, nothing freed (src/world.c:138-144)
        assert lib.frGetBodyCountInWorld(w) == 0 and not lib.frIsBodyInWorld(w, first[0])
        second = [add(lib, w, lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(i * 0.8, 5), c)) for i in range(4)]
        for _ in range(20):
            lib.frStepWorld(w, DT)
        assert lib.frGetBodyCountInWorld(w) == 4
        p = lib.frGetBodyPosition(first[0])
        assert p.y > 0.0 and lib.frGetBodyPosition(second[0]).y > 5.0
        for b in first:
            lib.frReleaseBody(b)
        lib.frReleaseWorld(w)
    fb.check_error(gpu)


def test_default_capacity_is_the_reference_macro(gpu):
    """Without frxSetWorldCapacity the world holds FR_WORLD_MAX_OBJECT_COUNT (2048) bodies like the
    reference (src/world.c:148-150)."""
    L = gpu
    w = L.frCreateWorld(capi.Vector2(0, 9.8), 2.0)
    c = L.frCreateCircle(mat(), 0.5)
    for rounds in range(2):
        for i in range(1500):
            L.frAddBodyToWorld(w, L.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(2.0 * (i % 200), -2.0 * (i // 200) - 40 * rounds), c))
        L.frStepWorld(w, DT)
    # 3,000 requests were accepted by the queue over two steps (the length check is made at enqueue
    # time only, exactly like the reference); the third batch is refused because length >= 2048
    assert L.frGetBodyCountInWorld(w) == 3000
    assert not L.frAddBodyToWorld(w, L.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0, 0), c))
    assert L.frxGetWorldCapacity(w) == 2048
    fb.check_error(L)


def test_body_moved_by_a_setter_every_frame_keeps_the_graph_and_the_reference_trajectory(gpu, ref):
    """A kinematic platform nudged with frSetBodyPosition before EVERY step (what reference programs do, e.g.
    examples/src/cows.c): no sizing pre-pass (a blocking extra broadphase per step in round 1), the captured graph of the
    step keeps being replayed, and in reference order the trajectory equals the unmodified reference bit for bit
    (frSetBodyPosition's absolute AABB shift, App. A.9, included)."""
    L = gpu
    base = scenes.mixed_pile(300, width=30.0)
    spec = scenes.SceneSpec(name="platform", cell=base.cell, shapes=list(base.shapes) + [scenes.ShapeSpec("rect", (6.0, 1.0))])
    scenes._finish(spec, np.concatenate([base.body_type, [capi.BODY_KINEMATIC]]), np.concatenate([base.body_shape, [len(base.shapes)]]),
                   np.concatenate([base.pos, [[15.0, -30.0]]]), np.concatenate([base.angle, [0.0]]))
    runs = []
    for lib, mode in ((ref, None), (L, fb.SOLVER_SERIAL), (L, fb.SOLVER_COLORED)):
        sc = scenes.Scene(lib, spec, prime=False)
        if mode is not None:
            lib.frxSetWorldSolverMode(sc.world, mode)
        lib.frStepWorld(sc.world, DT)
        plat = sc.bodies[-1]
        for s in range(240):
            lib.frSetBodyPosition(plat, capi.Vector2(15.0 + 4.0 * float(np.sin(0.05 * s)), -30.0 + 0.02 * s))
            lib.frStepWorld(sc.world, DT)
        runs.append((sc, scenes.snapshot(lib, sc.bodies)))
    fb.check_error(L)
    for k in ("pos", "angle", "vel", "omega"):
        assert runs[0][1][k].tobytes() == runs[1][1][k].tobytes(), k
    assert L.frxGetGraphReplays(runs[2][0].world) >= 200
    assert np.isfinite(runs[2][1]["pos"]).all()
    for sc, _ in runs:
        sc.release()
