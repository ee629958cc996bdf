"""Shared helpers of the parity tests."""
import ctypes as C

import numpy as np

from ferox_b200 import capi, scenes
from oracle import orc


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.tobytes() != b.tobytes():
        bad = np.argwhere(bits(a.astype(np.float32, copy=False)) != bits(b.astype(np.float32, copy=False)))
        i = tuple(bad[0])
        raise AssertionError("%s: %d of %d words differ, first at %s: %r vs %r" % (what, len(bad), a.size, i, a[i], b[i]))


def assert_states_equal(a, b, what=""):
    for k in ("pos", "angle", "sincos", "vel", "omega", "aabb"):
        assert_bits_equal(a[k], b[k], "%s %s" % (what, k))


def records_to_entries(recs):
    out = np.zeros(len(recs), orc.ENTRY_DTYPE)
    for k, r in enumerate(recs):
        out[k]["a"], out[k]["b"], out[k]["count"] = r[0], r[1], r[2]
        out[k]["dir"] = r[3:5]
        out[k]["friction"], out[k]["restitution"] = r[5], r[6]
        for c in range(2):
            o = 7 + 8 * c
            ct = out[k]["contacts"][c]
            ct["id"], ct["depth"] = r[o], r[o + 1]
            ct["point"] = r[o + 2:o + 4]
            ct["normalMass"], ct["normalScalar"], ct["tangentMass"], ct["tangentScalar"] = r[o + 4:o + 8]
    return out


CACHE_FIELDS = ("a", "b", "count", "dir", "friction", "restitution")
CONTACT_FIELDS = ("id", "depth", "point", "normalMass", "normalScalar", "tangentMass", "tangentScalar")


def assert_caches_equal(x, y, what="", timestamps=False):
    """Two contact-cache dumps (ENTRY_DTYPE / MANIFOLD_DTYPE) entry for entry, field for field."""
    assert len(x) == len(y), (what, len(x), len(y))
    for f in CACHE_FIELDS:
        assert x[f].tobytes() == y[f].tobytes(), "%s: cache field %s differs" % (what, f)
    for f in CONTACT_FIELDS + (("timestamp",) if timestamps else ()):
        assert np.ascontiguousarray(x["contacts"][f]).tobytes() == np.ascontiguousarray(y["contacts"][f]).tobytes(), \
            "%s: contact field %s differs" % (what, f)


class RefCallLog:
    """The reference's narrow-phase call sequence of a step, via oracle/ref_shim.c."""

    class Call(C.Structure):
        _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("hit", C.c_int)]

    def __init__(self, ref):
        self.ref = ref
        ref.refshim_count.restype = C.c_size_t
        ref.refshim_calls.restype = C.POINTER(self.Call)

    def __enter__(self):
        self.ref.refshim_enable(1)
        self.ref.refshim_reset()
        return self

    def __exit__(self, *a):
        self.ref.refshim_enable(0)

    def reset(self):
        self.ref.refshim_reset()

    def calls(self, scene):
        n, p = self.ref.refshim_count(), self.ref.refshim_calls()
        return np.array([(scene.index_of[p[i].a], scene.index_of[p[i].b], p[i].hit) for i in range(n)],
                        np.int32).reshape(-1, 3)


def random_collision(rng, signed_zero=True):
    col = capi.Collision()
    col.count = int(rng.integers(1, 3))
    d = rng.normal(size=2)
    d /= np.linalg.norm(d)
    if rng.random() < 0.3:
        d = np.array([0.0, -1.0])
    col.direction = capi.Vector2(float(np.float32(d[0])), float(np.float32(d[1])))
    col.friction = float(np.float32(rng.choice([0, 0.5, 0.75])))
    col.restitution = float(np.float32(rng.choice([0, 0.2])))
    zeros = [0.0, -0.0] if signed_zero else [0.0]
    for k in range(2):
        ct = col.contacts[k]
        ct.depth = float(np.float32(rng.choice([0.0, 0.01, 0.02, -0.01, 0.5 * rng.random()])))
        ct.point = capi.Vector2(float(np.float32(rng.normal())), float(np.float32(99 + rng.normal())))
        ct.cache.normalMass = float(np.float32(0.4))
        ct.cache.tangentMass = float(np.float32(0.39))
        ct.cache.normalScalar = float(np.float32(rng.choice(zeros + [0.03])))
        ct.cache.tangentScalar = float(np.float32(rng.choice(zeros + [0.01, -0.02])))
    return col


def energy_stats(state, masses, gravity_y=9.8):
    """(kinetic, potential) of a body-state dict; y points down (gravity +y)."""
    v2 = (state["vel"].astype(np.float64) ** 2).sum(1)
    ke = 0.5 * (masses * v2).sum()
    pe = -(masses * gravity_y * state["pos"][:, 1].astype(np.float64)).sum()
    return ke, pe
