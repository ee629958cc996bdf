"""CPU: the C-ABI shared object loads, exports every symbol include/ferox_b200.h declares, keeps the
reference's POD layouts, and its host-side set-up code (shapes, mass properties, AABBs, the add
queue, the error convention) behaves like the reference.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ferox_b200 as fb
from ferox_b200 import capi, scenes
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ferox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frx?[A-Z]\w+)\s*\(", text)) - {"frHashQueryFunc", "frCollisionEventFunc", "frRaycastQueryFunc"})


def test_every_declared_symbol_is_exported(prod):
    names = declared_symbols()
    assert len(names) >= 88 + 20, len(names)
    missing = [n for n in names if not hasattr(prod, n)]
    assert not missing, missing
    # the reference API: 87 defined functions + frSetBodyTransform
    for n in list(capi.PROTOTYPES) + list(capi.OPTIONAL_PROTOTYPES):
        assert n in names, n


def test_pod_layouts():
    assert C.sizeof(capi.Vector2) == 8 and C.sizeof(capi.AABB) == 16 and C.sizeof(capi.ContextNode) == 16
    assert C.sizeof(capi.Contact) == 36 and C.sizeof(capi.Collision) == 92 and C.sizeof(capi.Transform) == 20
    assert C.sizeof(capi.Vertices) == 68 and C.sizeof(capi.RaycastHit) == 32 and C.sizeof(capi.Ray) == 20
    assert C.sizeof(capi.Material) == 12 and C.sizeof(capi.BodyPair) == 16 and C.sizeof(fb.Manifold) == 100


def test_static_archive_is_built():
    assert os.path.exists(os.path.join(ROOT, "ferox_b200", "_build", "libferox.a"))


def test_shape_tables_match_reference(ref, prod):
    """Hull order, normals, area, mass, inertia: the tables the kernels consume."""
    spec = scenes.straddle_origin(400)
    sa = [scenes.make_shape(ref, s) for s in spec.shapes]
    sb = [scenes.make_shape(prod, s) for s in spec.shapes]
    for x, y in zip(orc.describe_shapes(ref, sa), orc.describe_shapes(prod, sb)):
        for u, v in zip(x, y):
            assert (u.tobytes() == v.tobytes()) if isinstance(u, np.ndarray) else (u == v)
    for s1, s2 in zip(sa, sb):
        for fn in ("frGetShapeArea", "frGetShapeMass", "frGetShapeInertia", "frGetShapeDensity"):
            assert getattr(ref, fn)(s1) == getattr(prod, fn)(s2), fn
    # degenerate inputs: fewer than 3 points -> empty hull, zero area (src/geometry.c:362)
    v = capi.Vertices()
    v.count = 2
    v.data[0], v.data[1] = capi.Vector2(0, 0), capi.Vector2(1, 0)
    for lib in (ref, prod):
        s = lib.frCreatePolygon(capi.Material(1, 0, 0), C.byref(v))
        assert lib.frGetPolygonVertices(s).contents.count == 0 and lib.frGetShapeArea(s) == 0.0
    # collinear points and a non-convex input go through the same gift wrapping
    pts = [(0, 0), (1, 0), (2, 0), (2, 2), (1, 0.5), (0, 2)]
    v.count = len(pts)
    for i, p in enumerate(pts):
        v.data[i] = capi.Vector2(*map(float, p))
    hulls = []
    for lib in (ref, prod):
        s = lib.frCreatePolygon(capi.Material(1, 0, 0), C.byref(v))
        pv = lib.frGetPolygonVertices(s).contents
        hulls.append([(pv.data[i].x, pv.data[i].y) for i in range(pv.count)] + [lib.frGetShapeArea(s), lib.frGetShapeInertia(s)])
    assert hulls[0] == hulls[1]


def test_body_setup_matches_reference(ref, prod):
    spec = scenes.straddle_origin(200)
    sa = [scenes.make_shape(ref, s) for s in spec.shapes]
    sb = [scenes.make_shape(prod, s) for s in spec.shapes]
    for i in range(spec.n):
        pair = []
        for lib, shapes in ((ref, sa), (prod, sb)):
            b = lib.frCreateBodyFromShape(int(spec.body_type[i]), capi.Vector2(*map(float, spec.pos[i])), shapes[spec.body_shape[i]])
            lib.frSetBodyAngle(b, float(spec.angle[i]))
            lib.frSetBodyPosition(b, capi.Vector2(1.5, 2.5))        # AABB shifts by the absolute position (App. A.9)
            lib.frApplyForceToBody(b, capi.Vector2(2.0, 3.0), capi.Vector2(0.5, -1.0))
            lib.frApplyImpulseToBody(b, capi.Vector2(1.0, 2.0), capi.Vector2(0.25, 0.5))
            lib.frApplyGravityToBody(b, capi.Vector2(0.0, 9.8))
            if i % 3 == 0:
                lib.frSetBodyFlags(b, capi.FLAG_INFINITE_INERTIA)
            snap = scenes.snapshot(lib, [b])
            f = lib.frGetBodyForce(b)
            extra = (lib.frGetBodyMass(b), lib.frGetBodyInverseMass(b), lib.frGetBodyInertia(b), lib.frGetBodyInverseInertia(b),
                     f.x, f.y, lib.frGetBodyTorque(b), lib.frGetBodyGravityScale(b), lib.frGetBodyType(b), lib.frGetBodyFlags(b))
            pair.append((snap, extra))
        for k in pair[0][0]:
            assert pair[0][0][k].tobytes() == pair[1][0][k].tobytes(), (i, k)
        assert pair[0][1] == pair[1][1], i


def test_null_and_invalid_arguments_are_neutral(prod):
    """Error convention of the reference: NULL / invalid in, neutral value out, never abort."""
    L = prod
    assert L.frCreateCircle(capi.Material(1, 0, 0), 0.0) is None
    assert L.frCreateRectangle(capi.Material(1, 0, 0), -1.0, 1.0) is None
    assert L.frCreatePolygon(capi.Material(1, 0, 0), None) is None
    assert L.frCreateBody(capi.BODY_UNKNOWN, capi.Vector2(0, 0)) is None
    assert L.frCreateSpatialHash(0.0) is None
    assert L.frGetBodyType(None) == capi.BODY_UNKNOWN and L.frGetBodyMass(None) == 0.0
    assert L.frGetShapeType(None) == capi.SHAPE_UNKNOWN and L.frGetCircleRadius(None) == 0.0
    assert L.frGetBodyCountInWorld(None) == 0 and L.frGetBodyInWorld(None, 0) is None
    assert not L.frAddBodyToWorld(None, None) and not L.frComputeCollision(None, None, None)
    L.frStepWorld(None, 1.0 / 60.0)
    L.frUpdateWorld(None, 1.0 / 60.0)
    L.frReleaseWorld(None), L.frReleaseBody(None), L.frReleaseShape(None), L.frReleaseSpatialHash(None)
    t = L.frGetBodyTransform(None)
    assert (t.position.x, t.position.y, t.angle) == (0.0, 0.0, 0.0)


def test_spatial_hash_and_raycast_match_reference(ref, prod):
    """The stand-alone query API (host side): ascending unique ids, truncation toward zero."""
    out = []
    for lib in (ref, prod):
        sh = lib.frCreateSpatialHash(2.0)
        rng = np.random.default_rng(5)
        for i in range(200):
            x, y, w, h = rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(0.1, 5), rng.uniform(0.1, 5)
            lib.frInsertIntoSpatialHash(sh, capi.AABB(x, y, w, h), i)
        got = []
        cb = capi.HashQueryFunc(lambda node: got.append(node.id) or True)
        lib.frQuerySpatialHash(sh, capi.AABB(-0.5, -0.5, 3.0, 3.0), cb, None)
        lib.frQuerySpatialHash(sh, capi.AABB(-7.0, 3.0, 9.0, 1.0), cb, None)
        s = lib.frCreateRectangle(capi.Material(1, .5, 0), 2.0, 1.0)
        b = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(3.0, 0.2), s)
        lib.frSetBodyAngle(b, 0.3)
        hit = capi.RaycastHit()
        ok = lib.frComputeRaycast(b, capi.Ray(capi.Vector2(0, 0), capi.Vector2(1, 0.05), 10.0), C.byref(hit))
        c = lib.frCreateCircle(capi.Material(1, .5, 0), 0.75)
        bc = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(-2.0, 0.1), c)
        hit2 = capi.RaycastHit()
        ok2 = lib.frComputeRaycast(bc, capi.Ray(capi.Vector2(0, 0), capi.Vector2(-1, 0), 10.0), C.byref(hit2))
        out.append((got, ok, hit.point.t(), hit.normal.t(), hit.inside, ok2, hit2.point.t(), hit2.distance,
                    lib.frBodyContainsPoint(b, capi.Vector2(3.1, 0.2)), lib.frBodyContainsPoint(bc, capi.Vector2(5, 5))))
    assert out[0] == out[1]


def test_add_queue_capacity_matches_reference(ref, prod):
    """At most FR_WORLD_MAX_OBJECT_COUNT - 1 requests fit between two steps (ring with one dead slot,
    src/external/ferox_utils.h:197-205); bodies are not members before the next post-step."""
    counts = []
    for lib, cap in ((ref, 8192), (prod, 2048)):
        w = lib.frCreateWorld(capi.Vector2(0, 9.8), 2.0)
        s = lib.frCreateCircle(capi.Material(1, .5, 0), 0.5)
        ok = 0
        bodies = []
        for i in range(cap + 10):
            b = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(float(i), 0.0), s)
            bodies.append(b)
            ok += bool(lib.frAddBodyToWorld(w, b))
        counts.append((ok - cap, lib.frGetBodyCountInWorld(w), lib.frIsBodyInWorld(w, bodies[0])))
    assert counts[0] == counts[1] == (-1, 0, False)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_step_refuses_without_gpu(prod):
    """No CPU fallback: without a CUDA device the step is refused loudly and the state is untouched."""
    L = prod
    w = L.frCreateWorld(capi.Vector2(0, 9.8), 2.0)
    assert L.frxGetLastError() != 0
    s = L.frCreateCircle(capi.Material(1, .5, 0), 0.5)
    b = L.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(1.0, 2.0), s)
    assert L.frAddBodyToWorld(w, b)
    L.frStepWorld(w, 1.0 / 60.0)
    assert L.frGetBodyCountInWorld(w) == 0 and L.frGetBodyPosition(b).t() == (1.0, 2.0)
