"""ferox_b200 -- B200-native world step behind the ferox.h C ABI.

The product is ``ferox_b200/_build/libferox.so`` (host C++ + hand-written sm_100a kernels, built by
``ferox_b200/csrc/Makefile``).  This package only loads it and attaches ctypes prototypes; there is
no Python or CPU implementation of the step to fall back to: if the shared object is missing,
`lib()` raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("FEROX_B200_LIB") or os.path.join(HERE, "_build", "libferox.so")   # env: A/B builds in dev probes

SOLVER_COLORED, SOLVER_SERIAL = 0, 1


class StepCounters(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in
                ("bodies", "cellEntries", "candidates", "manifolds", "contacts", "colors", "steps", "overflow", "colorRounds")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Manifold(C.Structure):
    _fields_ = [("first", C.c_int), ("second", C.c_int), ("value", capi.Collision)]


P, F, I = C.c_void_p, C.c_float, C.c_int
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

# include/ferox_b200.h, part 2
EXT_PROTOTYPES = {
    "frxSetDevice": (I, [I]),
    "frxGetLastError": (I, []),
    "frxGetLastErrorString": (C.c_char_p, []),
    "frxSetWorldCapacity": (C.c_bool, [P, I]),
    "frxGetWorldCapacity": (I, [P]),
    "frxCreateWorldBatch": (P, [capi.Vector2, F, I]),
    "frxAddBodyToBatch": (C.c_bool, [P, I, P]),
    "frxGetBatchWorldCount": (I, [P]),
    "frxStripConfigure": (I, [P, F, F, F, I, C.c_uint, C.c_uint]),
    "frxStripRegisterShapes": (I, [P, P, I]),
    "frxStripLinkLocal": (I, [P, P]),
    "frxStripGetNcclUniqueId": (I, [P]),
    "frxStripInitNccl": (I, [P, I, I, P]),
    "frxStripStep": (None, [C.POINTER(C.c_void_p), I, F]),
    "frxStripMigrate": (I, [C.POINTER(C.c_void_p), I]),
    "frxStripGetTransport": (I, [P]),
    "frxSetWorldSolverMode": (None, [P, I]),
    "frxGetWorldSolverMode": (I, [P]),
    "frxSetWorldTimestamp": (None, [P, F]),
    "frxGetWorldTimestamp": (F, [P]),
    "frxStepWorldN": (None, [P, F, I]),
    "frxSynchronizeWorld": (None, [P]),
    "frxGetStepCounters": (None, [P, C.POINTER(StepCounters)]),
    "frxResetStepCounters": (None, [P]),
    "frxGetBodyStates": (I, [P, P, P, P, P, P, P]),
    "frxViewBodyStates": (I, [P, P, P, P, P]),
    "frxClearLastError": (None, []),
    "frxMapBodyForces": (I, [P, P]),
    "frxApplyForces": (I, [P, P, P, I]),
    "frxSetBodyVelocities": (I, [P, P, P, I]),
    "frxCreateBodies": (I, [P, I, i32p, i32p, P, f32p, P, P]),
    "frxCreateBodiesEx": (I, [P, I, i32p, i32p, P, f32p, P, P, P, P, P]),
    "frxGetCandidatePairs": (I, [P, P, I]),
    "frxGetManifolds": (I, [P, P, I]),
    "frxComputeCollisions": (I, [P, i32p, I, P, P]),
    "frxRaycastBatch": (I, [P, P, I, P, i32p, I]),
    "frxPointQueryBatch": (I, [P, P, I, i32p, i32p, I]),
    "frxDeviceSinCos": (I, [f32p, I, f32p, f32p]),
    "frxSetProfiling": (None, [P, I]),
    "frxGetStageTimes": (I, [P, C.POINTER(C.c_double)]),
    "frxRecordEvent": (None, [P, I]),
    "frxEventElapsedMs": (F, [P, I, I]),
    "frxGetLaunchCount": (C.c_longlong, []),
    "frxGetGraphReplays": (C.c_longlong, [P]),
    "frxGetWorldSolverKernel": (I, [P]),
    "frxGetColorHistogram": (I, [P, i32p]),
    "frxGetSolveTrace": (I, [P, C.POINTER(C.c_ulonglong), I]),
    "frxProfilerRange": (None, [I]),
    "frxTimedSteps": (I, [P, F, I, I, f32p]),
}

_lib = None


def build(verbose=False):
    """Compile the shared object in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libferox.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib():
    """The loaded product library with ferox.h + frx* prototypes attached.  Never a fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the world step)" % LIB_PATH)
        L = capi.load_abi(LIB_PATH)
        capi.bind(L, EXT_PROTOTYPES)
        _lib = L
    return _lib


def check_error(L=None):
    L = L or lib()
    code = L.frxGetLastError()
    if code:
        raise RuntimeError("libferox error %d: %s" % (code, L.frxGetLastErrorString().decode()))


def step_counters(L, world):
    c = StepCounters()
    L.frxGetStepCounters(world, C.byref(c))
    return c.as_dict()


def body_states(L, world):
    n = L.frGetBodyCountInWorld(world)
    out = {"pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
           "sincos": np.empty((n, 2), np.float32), "vel": np.empty((n, 2), np.float32),
           "omega": np.empty(n, np.float32), "aabb": np.empty((n, 4), np.float32)}
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    L.frxGetBodyStates(world, ptr(out["pos"]), ptr(out["angle"]), ptr(out["sincos"]), ptr(out["vel"]),
                       ptr(out["omega"]), ptr(out["aabb"]))
    return out


def body_state_views(L, world):
    """Zero-copy numpy views of the pinned host mirror (valid until the world is stepped or edited)."""
    f32p = C.POINTER(C.c_float)
    p, a, v, b = f32p(), f32p(), f32p(), f32p()
    n = L.frxViewBodyStates(world, C.byref(p), C.byref(a), C.byref(v), C.byref(b))
    if n == 0:
        z = np.zeros((0, 4), np.float32)
        return {"posrot": z, "angle": np.zeros(0, np.float32), "vel": z, "aabb": z}
    arr = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
    return {"posrot": arr(p, (n, 4)), "angle": arr(a, (n,)), "vel": arr(v, (n, 4)), "aabb": arr(b, (n, 4))}


def candidate_pairs(L, world):
    n = L.frxGetCandidatePairs(world, None, 0)
    out = np.empty((n, 3), np.int32)
    if n:
        L.frxGetCandidatePairs(world, out.ctypes.data_as(C.c_void_p), n)
    return out


MANIFOLD_DTYPE = np.dtype([
    ("a", np.int32), ("b", np.int32), ("count", np.int32), ("dir", np.float32, 2),
    ("contacts", [("id", np.int32), ("depth", np.float32), ("timestamp", np.float32),
                  ("point", np.float32, 2), ("normalMass", np.float32), ("normalScalar", np.float32),
                  ("tangentMass", np.float32), ("tangentScalar", np.float32)], 2),
    ("friction", np.float32), ("restitution", np.float32)])
assert MANIFOLD_DTYPE.itemsize == C.sizeof(Manifold) == 100


def manifolds(L, world):
    n = L.frxGetManifolds(world, None, 0)
    out = np.zeros(n, MANIFOLD_DTYPE)
    if n:
        L.frxGetManifolds(world, out.ctypes.data_as(C.c_void_p), n)
    return out
