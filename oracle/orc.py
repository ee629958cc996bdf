"""ctypes front end of oracle/step_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
`OracleWorld.from_scene` copies a world that was built through the ferox.h C ABI (on the reference
or on the product) into the oracle's flat arrays using public getters only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libstep_oracle.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

ENTRY_DTYPE = np.dtype([
    ("a", np.int32), ("b", np.int32), ("count", np.int32), ("dir", np.float32, 2),
    ("contacts", [("id", np.int32), ("depth", np.float32), ("timestamp", np.float32),
                  ("point", np.float32, 2), ("normalMass", np.float32), ("normalScalar", np.float32),
                  ("tangentMass", np.float32), ("tangentScalar", np.float32)], 2),
    ("friction", np.float32), ("restitution", np.float32)])
COLLISION_DTYPE = np.dtype([(n, ENTRY_DTYPE.fields[n][0]) for n in
                            ("count", "dir", "contacts", "friction", "restitution")])
assert ENTRY_DTYPE.itemsize == 100 and COLLISION_DTYPE.itemsize == 92

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_shape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, f32p, f32p,
                                    C.c_float, C.c_float]
        L.orc_set_bodies.argtypes = [C.c_void_p, i32p, i32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p]
        L.orc_set_forces.argtypes = [C.c_void_p, f32p, f32p]
        L.orc_set_timestamp.argtypes = [C.c_void_p, C.c_float]
        L.orc_set_solve_order.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong]
        L.orc_last_colors.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_float]
        L.orc_get_state.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, f32p, f32p]
        L.orc_candidate_count.argtypes = [C.c_void_p]
        L.orc_get_candidates.argtypes = [C.c_void_p, i32p]
        L.orc_cache_count.argtypes = [C.c_void_p]
        L.orc_get_cache.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_counters.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.int64)]
        L.orc_collide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        assert L.orc_entry_size() == ENTRY_DTYPE.itemsize
        _lib = L
    return _lib


def describe_shapes(abi, shapes):
    """(type, radius, nv, verts[8,2], normals[8,2], friction, restitution) per frShape handle."""
    out = []
    for s in shapes:
        t = abi.frGetShapeType(s)
        verts = np.zeros((8, 2), np.float32)
        normals = np.zeros((8, 2), np.float32)
        nv = 0
        if t == 2:
            pv = abi.frGetPolygonVertices(s).contents
            pn = abi.frGetPolygonNormals(s).contents
            nv = pv.count
            for k in range(nv):
                verts[k] = (pv.data[k].x, pv.data[k].y)
                normals[k] = (pn.data[k].x, pn.data[k].y)
        out.append((t, abi.frGetCircleRadius(s), nv, verts, normals,
                    abi.frGetShapeFriction(s), abi.frGetShapeRestitution(s)))
    return out


class OracleWorld:
    def __init__(self, n, shape_desc, cell, gravity):
        self.L = lib()
        self.n = n
        self.h = self.L.orc_create(n, len(shape_desc), cell, gravity[0], gravity[1])
        for i, (t, r, nv, verts, normals, fr, re) in enumerate(shape_desc):
            self.L.orc_set_shape(self.h, i, t, r, nv, np.ascontiguousarray(verts.reshape(-1)),
                                 np.ascontiguousarray(normals.reshape(-1)), fr, re)

    @classmethod
    def from_scene(cls, scene):
        """Copy a scenes.Scene (any ferox-ABI library) into the oracle via public getters."""
        abi, spec = scene.lib, scene.spec
        n = len(scene.bodies)
        w = cls(n, describe_shapes(abi, scene.shapes), spec.cell, spec.gravity)
        shape_index = {s: i for i, s in enumerate(scene.shapes)}
        typ = np.empty(n, np.int32); shp = np.empty(n, np.int32)
        mass4 = np.empty((n, 4), np.float32); gs = np.empty(n, np.float32)
        for i, b in enumerate(scene.bodies):
            typ[i] = abi.frGetBodyType(b)
            shp[i] = shape_index.get(abi.frGetBodyShape(b), -1)
            mass4[i] = (abi.frGetBodyMass(b), abi.frGetBodyInverseMass(b),
                        abi.frGetBodyInertia(b), abi.frGetBodyInverseInertia(b))
            gs[i] = abi.frGetBodyGravityScale(b)
        st = scene.snapshot()
        w.set_bodies(typ, shp, st["pos"], st["angle"], st["sincos"], st["vel"], st["omega"], mass4, gs,
                     st["aabb"])
        return w

    def set_bodies(self, typ, shp, pos, angle, sincos, vel, omega, mass4, gs, aabb):
        c = np.ascontiguousarray
        self.L.orc_set_bodies(self.h, c(typ, np.int32), c(shp, np.int32), c(pos, np.float32).reshape(-1),
                              c(angle, np.float32), c(sincos, np.float32).reshape(-1),
                              c(vel, np.float32).reshape(-1), c(omega, np.float32),
                              c(mass4, np.float32).reshape(-1), c(gs, np.float32),
                              c(aabb, np.float32).reshape(-1))

    def set_forces(self, force, torque):
        self.L.orc_set_forces(self.h, np.ascontiguousarray(force, np.float32).reshape(-1),
                              np.ascontiguousarray(torque, np.float32))

    def set_solve_order(self, mode, seed=1):
        """0: reference (array) order; 1: greedy colour-major; 2: random permutation per step.  1 and 2 are NOT the
        reference: they measure how much a pure re-ordering of the sweeps moves the statistics (App. A.13)."""
        self.L.orc_set_solve_order(self.h, {"reference": 0, "colored": 1, "random": 2}.get(mode, mode), seed)

    def set_timestamp(self, ts):
        self.L.orc_set_timestamp(self.h, ts)

    def step(self, n=1, dt=float(np.float32(1.0 / 60.0))):
        for _ in range(n):
            self.L.orc_step(self.h, dt)

    def state(self):
        n = self.n
        out = {"pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
               "sincos": np.empty((n, 2), np.float32), "vel": np.empty((n, 2), np.float32),
               "omega": np.empty(n, np.float32), "aabb": np.empty((n, 4), np.float32)}
        self.L.orc_get_state(self.h, out["pos"].reshape(-1), out["angle"], out["sincos"].reshape(-1),
                             out["vel"].reshape(-1), out["omega"], out["aabb"].reshape(-1))
        return out

    def candidates(self):
        m = self.L.orc_candidate_count(self.h)
        out = np.empty((m, 3), np.int32)
        if m:
            self.L.orc_get_candidates(self.h, out.reshape(-1))
        return out

    def cache(self):
        m = self.L.orc_cache_count(self.h)
        out = np.empty(m, ENTRY_DTYPE)
        if m:
            self.L.orc_get_cache(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def counters(self):
        out = np.zeros(4, np.int64)
        self.L.orc_get_counters(self.h, out)
        return dict(K=int(out[0]), C=int(out[1]), M=int(out[2]), Q=int(out[3]))

    def collide(self, a, b):
        out = np.zeros(1, COLLISION_DTYPE)
        hit = self.L.orc_collide(self.h, a, b, out.ctypes.data_as(C.c_void_p))
        return bool(hit), out[0]

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None
