"""ctypes view of the ferox C ABI.

The boundary of this project IS the reference's public C API (reference include/ferox.h:276-592):
existing C programs relink against our host library.  This module binds that same ABI from Python so
that the parity tests and the bench can drive EITHER shared object through identical calls:

  * ``ferox_b200/_build/libferox.so``  -- the product (host C++ + sm_100a kernels), and
  * ``oracle/_ref/libferox_ref_<cap>.so`` -- the unmodified reference, test infrastructure only.

POD layouts follow reference include/ferox.h:122-271 (sizes per SURVEY.md section 8: frContact 36 B,
frCollision 92 B, frTransform 20 B, frVertices 68 B, frRaycastHit 32 B).
"""
import ctypes as C
import os

MAX_VERTS = 8  # FR_GEOMETRY_MAX_VERTEX_COUNT, reference include/ferox.h:43-46

# body / shape enums, reference include/ferox.h:204-241
SHAPE_UNKNOWN, SHAPE_CIRCLE, SHAPE_POLYGON = 0, 1, 2
BODY_UNKNOWN, BODY_STATIC, BODY_KINEMATIC, BODY_DYNAMIC = 0, 1, 2, 3
FLAG_NONE, FLAG_INFINITE_MASS, FLAG_INFINITE_INERTIA = 0, 1, 2


class Vector2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]

    def t(self):
        return (self.x, self.y)


class AABB(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("width", C.c_float), ("height", C.c_float)]


class ContextNode(C.Structure):
    _fields_ = [("id", C.c_int), ("ctx", C.c_void_p)]


class ContactCache(C.Structure):
    _fields_ = [("normalMass", C.c_float), ("normalScalar", C.c_float),
                ("tangentMass", C.c_float), ("tangentScalar", C.c_float)]


class Contact(C.Structure):
    _fields_ = [("id", C.c_int), ("depth", C.c_float), ("timestamp", C.c_float),
                ("point", Vector2), ("cache", ContactCache)]


class Collision(C.Structure):
    _fields_ = [("count", C.c_int), ("direction", Vector2), ("contacts", Contact * 2),
                ("friction", C.c_float), ("restitution", C.c_float)]


class Ray(C.Structure):
    _fields_ = [("origin", Vector2), ("direction", Vector2), ("maxDistance", C.c_float)]


class RaycastHit(C.Structure):
    _fields_ = [("body", C.c_void_p), ("point", Vector2), ("normal", Vector2),
                ("distance", C.c_float), ("inside", C.c_bool)]


class Material(C.Structure):
    _fields_ = [("density", C.c_float), ("friction", C.c_float), ("restitution", C.c_float)]


class Vertices(C.Structure):
    _fields_ = [("data", Vector2 * MAX_VERTS), ("count", C.c_int)]


class Rotation(C.Structure):
    _fields_ = [("sin_", C.c_float), ("cos_", C.c_float)]


class Transform(C.Structure):
    _fields_ = [("position", Vector2), ("rotation", Rotation), ("angle", C.c_float)]


class BodyPair(C.Structure):
    _fields_ = [("first", C.c_void_p), ("second", C.c_void_p)]


HashQueryFunc = C.CFUNCTYPE(C.c_bool, ContextNode)
CollisionEventFunc = C.CFUNCTYPE(None, BodyPair, C.POINTER(Collision))
RaycastQueryFunc = C.CFUNCTYPE(None, RaycastHit, C.c_void_p)


class CollisionHandler(C.Structure):
    _fields_ = [("preStep", CollisionEventFunc), ("postStep", CollisionEventFunc)]


assert C.sizeof(Contact) == 36 and C.sizeof(Collision) == 92 and C.sizeof(Transform) == 20
assert C.sizeof(Vertices) == 68 and C.sizeof(RaycastHit) == 32 and C.sizeof(ContextNode) == 16

P = C.c_void_p
F = C.c_float
I = C.c_int
B = C.c_bool

# name -> (restype, argtypes); one row per function declared in reference include/ferox.h:281-592.
PROTOTYPES = {
    # broad_phase (ferox.h:281-299)
    "frCreateSpatialHash": (P, [F]),
    "frReleaseSpatialHash": (None, [P]),
    "frClearSpatialHash": (None, [P]),
    "frGetSpatialHashCellSize": (F, [P]),
    "frInsertIntoSpatialHash": (None, [P, AABB, I]),
    "frQuerySpatialHash": (None, [P, AABB, HashQueryFunc, P]),
    # collision (ferox.h:307-310)
    "frComputeCollision": (B, [P, P, C.POINTER(Collision)]),
    "frComputeRaycast": (B, [P, Ray, C.POINTER(RaycastHit)]),
    # geometry (ferox.h:315-399)
    "frCreateCircle": (P, [Material, F]),
    "frCreateRectangle": (P, [Material, F, F]),
    "frCreatePolygon": (P, [Material, C.POINTER(Vertices)]),
    "frReleaseShape": (None, [P]),
    "frGetShapeType": (I, [P]),
    "frGetShapeMaterial": (Material, [P]),
    "frGetShapeDensity": (F, [P]),
    "frGetShapeFriction": (F, [P]),
    "frGetShapeRestitution": (F, [P]),
    "frGetShapeArea": (F, [P]),
    "frGetShapeMass": (F, [P]),
    "frGetShapeInertia": (F, [P]),
    "frGetShapeAABB": (AABB, [P, Transform]),
    "frGetCircleRadius": (F, [P]),
    "frGetPolygonVertex": (Vector2, [P, I]),
    "frGetPolygonVertices": (C.POINTER(Vertices), [P]),
    "frGetPolygonNormal": (Vector2, [P, I]),
    "frGetPolygonNormals": (C.POINTER(Vertices), [P]),
    "frSetShapeType": (None, [P, I]),
    "frSetShapeMaterial": (None, [P, Material]),
    "frSetShapeDensity": (None, [P, F]),
    "frSetShapeFriction": (None, [P, F]),
    "frSetShapeRestitution": (None, [P, F]),
    "frSetCircleRadius": (None, [P, F]),
    "frSetRectangleDimensions": (None, [P, F, F]),
    "frSetPolygonVertices": (None, [P, C.POINTER(Vertices)]),
    # rigid_body (ferox.h:404-531)
    "frCreateBody": (P, [I, Vector2]),
    "frCreateBodyFromShape": (P, [I, Vector2, P]),
    "frReleaseBody": (None, [P]),
    "frGetBodyType": (I, [P]),
    "frGetBodyFlags": (C.c_uint, [P]),
    "frGetBodyShape": (P, [P]),
    "frGetBodyTransform": (Transform, [P]),
    "frGetBodyPosition": (Vector2, [P]),
    "frGetBodyAngle": (F, [P]),
    "frGetBodyMass": (F, [P]),
    "frGetBodyInverseMass": (F, [P]),
    "frGetBodyInertia": (F, [P]),
    "frGetBodyInverseInertia": (F, [P]),
    "frGetBodyGravityScale": (F, [P]),
    "frGetBodyVelocity": (Vector2, [P]),
    "frGetBodyAngularVelocity": (F, [P]),
    "frGetBodyForce": (Vector2, [P]),
    "frGetBodyTorque": (F, [P]),
    "frGetBodyAABB": (AABB, [P]),
    "frGetBodyUserData": (P, [P]),
    "frSetBodyType": (None, [P, I]),
    "frSetBodyFlags": (None, [P, C.c_uint]),
    "frSetBodyShape": (None, [P, P]),
    "frSetBodyPosition": (None, [P, Vector2]),
    "frSetBodyAngle": (None, [P, F]),
    "frSetBodyGravityScale": (None, [P, F]),
    "frSetBodyVelocity": (None, [P, Vector2]),
    "frSetBodyAngularVelocity": (None, [P, F]),
    "frSetBodyUserData": (None, [P, P]),
    "frBodyContainsPoint": (B, [P, Vector2]),
    "frClearBodyForces": (None, [P]),
    "frApplyForceToBody": (None, [P, Vector2, Vector2]),
    "frApplyGravityToBody": (None, [P, Vector2]),
    "frApplyImpulseToBody": (None, [P, Vector2, Vector2]),
    "frApplyAccumulatedImpulses": (None, [P, P, C.POINTER(Collision)]),
    "frIntegrateForBodyVelocity": (None, [P, F]),
    "frIntegrateForBodyPosition": (None, [P, F]),
    "frResolveCollision": (None, [P, P, C.POINTER(Collision), F]),
    # timer (ferox.h:536)
    "frGetCurrentTime": (F, []),
    # world (ferox.h:544-592)
    "frCreateWorld": (P, [Vector2, F]),
    "frReleaseWorld": (None, [P]),
    "frClearWorld": (None, [P]),
    "frAddBodyToWorld": (B, [P, P]),
    "frRemoveBodyFromWorld": (B, [P, P]),
    "frIsBodyInWorld": (B, [P, P]),
    "frGetBodyInWorld": (P, [P, I]),
    "frGetBodyCountInWorld": (I, [P]),
    "frGetWorldGravity": (Vector2, [P]),
    "frSetWorldCollisionHandler": (None, [P, CollisionHandler]),
    "frSetWorldGravity": (None, [P, Vector2]),
    "frStepWorld": (None, [P, F]),
    "frUpdateWorld": (None, [P, F]),
    "frComputeWorldRaycast": (None, [P, Ray, RaycastQueryFunc, P]),
}
# declared at ferox.h:476 but defined nowhere in the reference; the product defines it.
OPTIONAL_PROTOTYPES = {"frSetBodyTransform": (None, [P, Transform])}


def bind(lib, table, required=True):
    missing = []
    for name, (res, args) in table.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and required:
        raise OSError("C ABI symbols missing from %s: %s" % (lib._name, ", ".join(missing)))
    return missing


def load_abi(path):
    """dlopen a library exporting the ferox.h ABI and attach prototypes."""
    if not os.path.exists(path):
        raise OSError("shared library not found: %s" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    bind(lib, PROTOTYPES)
    bind(lib, OPTIONAL_PROTOTYPES, required=False)
    return lib
