"""Deterministic scene recipes for the BASELINE.json configs.

A scene is first described as plain numpy arrays (`SceneSpec`) and then instantiated through the
ferox.h C ABI on whichever library the caller passes (the product or the compiled reference), so both
sides of a parity test see byte-identical inputs.  Body sizing follows SURVEY.md finding 3: contact
lever arms stay >= ~0.4 units because the reference solver (normalised lever arm, reference
src/rigid_body.c:465-474) diverges below ~0.35.
"""
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import capi

DT = float(np.float32(1.0 / 60.0))
GRAVITY = (0.0, 9.8)  # FR_WORLD_DEFAULT_GRAVITY, reference include/ferox.h:63-66 (y is down)


class SplitMix64:
    """splitmix64; the seeded generator SURVEY.md section 8d names for config 2."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def uniform(self):
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def randint(self, lo, hi):
        return lo + self.next() % (hi - lo + 1)


def splitmix_uniform(seed, n):
    """n uniforms in [0,1) from splitmix64, vectorised (same stream as SplitMix64.uniform)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class ShapeSpec:
    kind: str                      # "circle" | "rect" | "poly"
    params: tuple                  # (radius,) | (w, h) | ((x,y), ...)
    density: float = 1.0
    friction: float = 0.5
    restitution: float = 0.0


@dataclass
class SceneSpec:
    name: str
    cell: float
    gravity: tuple = GRAVITY
    shapes: list = field(default_factory=list)           # list[ShapeSpec]
    body_type: np.ndarray = None                         # int32 [n]
    body_shape: np.ndarray = None                        # int32 [n] index into shapes
    pos: np.ndarray = None                               # float32 [n,2]
    angle: np.ndarray = None                             # float32 [n]
    vel: np.ndarray = None                               # float32 [n,2] (optional)
    omega: np.ndarray = None                             # float32 [n]   (optional)

    @property
    def n(self):
        return int(self.body_type.shape[0])


def _finish(spec, types, shapes, pos, angle=None, vel=None, omega=None):
    spec.body_type = np.asarray(types, dtype=np.int32)
    spec.body_shape = np.asarray(shapes, dtype=np.int32)
    spec.pos = np.asarray(pos, dtype=np.float32).reshape(-1, 2)
    n = spec.n
    spec.angle = np.zeros(n, np.float32) if angle is None else np.asarray(angle, np.float32)
    spec.vel = np.zeros((n, 2), np.float32) if vel is None else np.asarray(vel, np.float32)
    spec.omega = np.zeros(n, np.float32) if omega is None else np.asarray(omega, np.float32)
    return spec


# --------------------------------------------------------------------------------------------------
# config 1: falling-box stack (SURVEY.md section 8d "Config 1"); runs on the unmodified reference
# --------------------------------------------------------------------------------------------------
def box_stack(cols=128, rows=8):
    spec = SceneSpec(name="config1_box_stack_%dx%d" % (cols, rows), cell=1.5)
    spec.shapes = [
        ShapeSpec("rect", (cols * 1.5 + 8.0, 2.0), friction=0.75),
        ShapeSpec("rect", (1.0, 1.0), friction=0.75),
    ]
    types, shapes, pos = [capi.BODY_STATIC], [0], [(cols * 0.75, 101.0)]
    for r in range(rows):
        for c in range(cols):
            types.append(capi.BODY_DYNAMIC)
            shapes.append(1)
            pos.append((1.5 * c + 0.75, 99.4 - 1.1 * r))
    return _finish(spec, types, shapes, pos)


# --------------------------------------------------------------------------------------------------
# config 2: mixed circles + convex polygons pile in a walled box (SURVEY.md section 8d "Config 2")
# --------------------------------------------------------------------------------------------------
def _poly_radius(nv):
    # keep the inradius >= 0.42 so face-contact lever arms stay in the reference's stable range
    return max(0.6, 0.42 / math.cos(math.pi / nv))


def mixed_pile(n=65536, seed=1234, width=None, pitch=1.7, shared_shapes=False, max_rows=96):
    """n dynamic bodies (50 % circles r=0.5, 50 % jittered regular polygons, 3-8 vertices) on a
    jittered lattice inside a box made of 3 static rectangles, dropped from rest.

    The box is at least 256 units wide and widened so that the lattice has at most `max_rows` rows:
    the reference's solver (and therefore ours) diverges in piles a few hundred rows deep ([probe]:
    a 353-row column reaches |omega| > 600 after 1,400 steps on the CPU oracle, a 95-row pile stays
    at KE/body ~0.5 for 2,000 steps), the same normalised-lever-arm pathology as SURVEY.md finding 3."""
    rng = SplitMix64(seed)
    if width is None:
        width = max(256.0, 2.0 + pitch * math.ceil(n / max_rows))
    spec = SceneSpec(name="config2_mixed_pile_%d" % n, cell=2.0)
    per_row = max(1, int((width - 2.0) // pitch))
    n_rows = (n + per_row - 1) // per_row
    height = n_rows * pitch + 4.0
    floor_y = 0.0
    spec.shapes = [
        ShapeSpec("rect", (width + 4.0, 2.0)),       # floor
        ShapeSpec("rect", (2.0, height + 2.0)),      # walls
        ShapeSpec("circle", (0.5,)),
    ]
    types = [capi.BODY_STATIC] * 3
    shapes = [0, 1, 1]
    pos = [(width * 0.5, floor_y + 1.0), (-1.0, floor_y - height * 0.5 + 1.0),
           (width + 1.0, floor_y - height * 0.5 + 1.0)]
    angle = [0.0, 0.0, 0.0]
    poly_cache = {}
    for k in range(n):
        row, col = divmod(k, per_row)
        x = 1.0 + pitch * (col + 0.5 * (row & 1)) + 0.05 * (rng.uniform() - 0.5) + 0.5 * pitch * 0.5
        y = floor_y - 1.0 - pitch * row - 0.05 * rng.uniform()
        is_poly = rng.uniform() < 0.5
        a = rng.uniform() * 2.0 * math.pi
        if is_poly:
            nv = rng.randint(3, 8)
            if shared_shapes and nv in poly_cache:
                sid = poly_cache[nv]
                for _ in range(nv):
                    rng.uniform()
            else:
                rad = _poly_radius(nv)
                verts = []
                for j in range(nv):
                    th = (j + 0.2 * (rng.uniform() - 0.5)) * 2.0 * math.pi / nv
                    verts.append((rad * math.cos(th), rad * math.sin(th)))
                spec.shapes.append(ShapeSpec("poly", tuple(verts)))
                sid = len(spec.shapes) - 1
                poly_cache[nv] = sid
            shapes.append(sid)
        else:
            shapes.append(2)
        types.append(capi.BODY_DYNAMIC)
        pos.append((x, y))
        angle.append(a)
    return _finish(spec, types, shapes, pos, angle)


# --------------------------------------------------------------------------------------------------
# config 5: one wide mixed pile, generated strip by strip (second row)
# --------------------------------------------------------------------------------------------------
def strip_pile_shapes(strip_width, height):
    """The shape table every strip of a strip_pile registers (same list, same order): floor segment,
    end wall, circle, regular 3..8-gons."""
    shapes = [ShapeSpec("rect", (strip_width + 128.0, 2.0)), ShapeSpec("rect", (2.0, height + 2.0)), ShapeSpec("circle", (0.5,))]
    for nv in range(3, 9):
        rad = _poly_radius(nv)
        shapes.append(ShapeSpec("poly", tuple((rad * math.cos(2.0 * math.pi * j / nv), rad * math.sin(2.0 * math.pi * j / nv))
                                              for j in range(nv))))
    return shapes


def strip_pile(per_strip, k, n_strips, seed=1234, pitch=1.7, max_rows=96, storeys=1, circle_fraction=0.5):
    """Strip k of ONE pile of n_strips * per_strip dynamic bodies (50 % circles, 50 % regular 3..8-gons
    at random angles) resting on `storeys` floors stacked vertically, each pile at most `max_rows`
    lattice rows deep (see mixed_pile; several storeys keep an 8M-body world narrow enough for
    float32 positions: 8 x 1M bodies on 8 storeys is 18.6k units wide, ulp 0.002).  The strip covers x
    in [k W, (k+1) W); its static bodies are its own floor segments (overlapping the neighbours' by 64
    units, statics are never ghosted) and the end wall(s) of the whole pile.
    Returns (spec, cut_lo, cut_hi)."""
    per_storey = (per_strip + storeys - 1) // storeys
    per_row = max(8, int(math.ceil(per_storey / max_rows)))
    width = per_row * pitch
    n_rows = (per_storey + per_row - 1) // per_row
    storey_h = n_rows * pitch + 40.0
    height = storeys * storey_h
    spec = SceneSpec(name="config5_strip_pile_%dx%d" % (n_strips, per_strip), cell=2.0)
    spec.shapes = strip_pile_shapes(width, height)
    x0 = k * width
    u = splitmix_uniform(seed + 7919 * k, 4 * per_strip).reshape(per_strip, 4)
    i = np.arange(per_strip)
    storey, j = i // per_storey, i % per_storey
    row, col = j // per_row, j % per_row
    x = x0 + pitch * (col + 0.25 + 0.5 * (row & 1) * 0.5) + 0.05 * (u[:, 0] - 0.5)
    y = -1.0 - pitch * row - 0.05 * u[:, 1] - storey * storey_h
    cf = float(circle_fraction)
    shape = np.where(u[:, 2] < cf, 2, 3 + np.minimum((u[:, 2] - cf) * (6.0 / max(1e-9, 1.0 - cf)), 5.0).astype(np.int64))
    ang = u[:, 3] * 2.0 * math.pi
    types, shapes, pos, angle = [], [], [], []
    for f in range(storeys):
        types.append(capi.BODY_STATIC); shapes.append(0); pos.append((x0 + 0.5 * width, 1.0 - f * storey_h)); angle.append(0.0)
    if k == 0:
        types.append(capi.BODY_STATIC); shapes.append(1); pos.append((-1.0, -height * 0.5 + 1.0)); angle.append(0.0)
    if k == n_strips - 1:
        types.append(capi.BODY_STATIC); shapes.append(1); pos.append((n_strips * width + 1.0, -height * 0.5 + 1.0)); angle.append(0.0)
    ns = len(types)
    spec = _finish(spec, np.concatenate([np.array(types, np.int32), np.full(per_strip, capi.BODY_DYNAMIC, np.int32)]),
                   np.concatenate([np.array(shapes, np.int64), shape]),
                   np.concatenate([np.array(pos, np.float64).reshape(ns, 2), np.stack([x, y], 1)]),
                   np.concatenate([np.array(angle, np.float64), ang]))
    lo = -math.inf if k == 0 else x0
    hi = math.inf if k == n_strips - 1 else x0 + width
    return spec, lo, hi


# --------------------------------------------------------------------------------------------------
# config 3: granular column collapse, uniform circles (SURVEY.md section 8d "Config 3")
# --------------------------------------------------------------------------------------------------
def column_collapse(n=1048576, cols=512, seed=1, floor_width=16384.0, pitch=1.05, jitter=0.02, friction=0.5, stagger=False,
                    row_pitch=None):
    spec = SceneSpec(name="config3_column_collapse_%d" % n, cell=2.0)
    rows = (n + cols - 1) // cols
    spec.shapes = [
        ShapeSpec("rect", (floor_width, 2.0)),
        ShapeSpec("rect", (2.0, rows * pitch + 8.0)),
        ShapeSpec("circle", (0.5,), friction=friction),
    ]
    k = np.arange(n)
    row, col = k // cols, k % cols
    u = splitmix_uniform(seed, 2 * n)
    rp = pitch if row_pitch is None else row_pitch
    x = 0.5 + pitch * col + jitter * (u[0::2] - 0.5) + (0.5 * pitch * (row & 1) if stagger else 0.0)
    y = -0.5 - rp * row + jitter * (u[1::2] - 0.5) - 0.5 * jitter
    pos = np.empty((n + 2, 2), np.float64)
    # floor top edge at y = 0, left wall right edge at x = 0; the right wall is "removed at t = 0"
    pos[0] = (floor_width * 0.5 - 64.0, 1.0)
    pos[1] = (-1.0, -(rows * pitch + 8.0) * 0.5 + 1.0)
    pos[2:, 0], pos[2:, 1] = x, y
    types = np.full(n + 2, capi.BODY_DYNAMIC, np.int32)
    types[:2] = capi.BODY_STATIC
    shapes = np.full(n + 2, 2, np.int32)
    shapes[0], shapes[1] = 0, 1
    return _finish(spec, types, shapes, pos)


# --------------------------------------------------------------------------------------------------
# config 4: one small RL-style world (256 bodies); many of these are batched / sharded
# --------------------------------------------------------------------------------------------------
def small_world(world_index=0, n_dynamic=252):
    rng = SplitMix64(0xC0FFEE + world_index)
    spec = SceneSpec(name="config4_world_%d" % world_index, cell=2.0)
    spec.shapes = [
        ShapeSpec("rect", (44.0, 2.0)),
        ShapeSpec("rect", (2.0, 60.0)),
        ShapeSpec("rect", (6.0, 1.0)),               # kinematic paddle
        ShapeSpec("circle", (0.5,)),
        ShapeSpec("rect", (1.0, 1.0)),
    ]
    types = [capi.BODY_STATIC, capi.BODY_STATIC, capi.BODY_STATIC, capi.BODY_KINEMATIC]
    shapes = [0, 1, 1, 2]
    pos = [(20.0, 1.0), (-1.0, -28.0), (41.0, -28.0), (20.0, -6.0)]
    angle = [0.0] * 4
    # the paddle is a kinematic stirrer: it spins in place (a translating paddle eventually crushes
    # bodies against a wall and the squeezed bodies are shot out of the world)
    vel = [(0.0, 0.0)] * 4
    omega = [0.0, 0.0, 0.0, 1.0]
    per_row = 24
    for k in range(n_dynamic):
        row, col = divmod(k, per_row)
        x = 1.5 + 1.6 * col + 0.1 * (rng.uniform() - 0.5)
        y = -10.0 - 1.6 * row - 0.1 * rng.uniform()
        box = rng.uniform() < 0.5
        types.append(capi.BODY_DYNAMIC)
        shapes.append(4 if box else 3)
        pos.append((x, y))
        angle.append(rng.uniform() * 2.0 * math.pi if box else 0.0)
        vel.append((0.0, 0.0))
        omega.append(0.0)
    return _finish(spec, types, shapes, pos, angle, vel, omega)


# --------------------------------------------------------------------------------------------------
# small hand-made scenes for unit-level parity
# --------------------------------------------------------------------------------------------------
def straddle_origin(n=1500, seed=7):
    """Mixed rotated bodies spread across x<0 and x>0 so that the truncate-toward-zero cell rule
    (reference src/broad_phase.c:104-108) is exercised on both sides of 'cell 0' (SURVEY App. A.2)."""
    rng = SplitMix64(seed)
    spec = SceneSpec(name="straddle_origin_%d" % n, cell=2.0)
    spec.shapes = [ShapeSpec("rect", (120.0, 2.0)), ShapeSpec("circle", (0.5,))]
    types, shapes, pos, angle = [capi.BODY_STATIC], [0], [(0.0, 1.0)], [0.0]
    side = int(math.sqrt(n)) + 1
    for k in range(n):
        row, col = divmod(k, side)
        x = (col - side / 2) * 1.15 + 0.2 * (rng.uniform() - 0.5)
        y = -0.6 - row * 1.15 + 0.2 * (rng.uniform() - 0.5)
        if rng.uniform() < 0.5:
            nv = rng.randint(3, 8)
            rad = _poly_radius(nv)
            verts = tuple((rad * math.cos((j + 0.2 * (rng.uniform() - 0.5)) * 2 * math.pi / nv),
                           rad * math.sin((j + 0.2 * (rng.uniform() - 0.5)) * 2 * math.pi / nv))
                          for j in range(nv))
            spec.shapes.append(ShapeSpec("poly", verts))
            shapes.append(len(spec.shapes) - 1)
        else:
            shapes.append(1)
        types.append(capi.BODY_DYNAMIC)
        pos.append((x, y))
        angle.append(rng.uniform() * 2.0 * math.pi)
    return _finish(spec, types, shapes, pos, angle)


# --------------------------------------------------------------------------------------------------
# instantiation through the C ABI
# --------------------------------------------------------------------------------------------------
class Scene:
    """A SceneSpec realised on one library: owns the frWorld, frBody and frShape handles."""

    def __init__(self, lib, spec, prime=True):
        self.lib, self.spec = lib, spec
        self.world = lib.frCreateWorld(capi.Vector2(*spec.gravity), spec.cell)
        if hasattr(lib, "frxSetWorldCapacity") and spec.n + 1 > 2047:
            # product only: the reference's capacity is the compile-time FR_WORLD_MAX_OBJECT_COUNT
            lib.frxSetWorldCapacity(self.world, spec.n + 16)
        self.shapes = [make_shape(lib, s) for s in spec.shapes]
        self.bodies = []
        pending = 0
        for i in range(spec.n):
            b = lib.frCreateBodyFromShape(int(spec.body_type[i]),
                                          capi.Vector2(float(spec.pos[i, 0]), float(spec.pos[i, 1])),
                                          self.shapes[int(spec.body_shape[i])])
            if spec.angle[i] != 0.0:
                lib.frSetBodyAngle(b, float(spec.angle[i]))
            if spec.vel[i, 0] != 0.0 or spec.vel[i, 1] != 0.0:
                lib.frSetBodyVelocity(b, capi.Vector2(float(spec.vel[i, 0]), float(spec.vel[i, 1])))
            if spec.omega[i] != 0.0:
                lib.frSetBodyAngularVelocity(b, float(spec.omega[i]))
            self.bodies.append(b)
            if not lib.frAddBodyToWorld(self.world, b):
                # the add queue holds < FR_WORLD_MAX_OBJECT_COUNT requests between two steps
                # (reference src/external/ferox_utils.h:197-205): flush with a step and retry.
                raise RuntimeError("frAddBodyToWorld refused body %d (capacity?)" % i)
            pending += 1
        self.index_of = {b: i for i, b in enumerate(self.bodies)}
        if prime:
            # bodies only enter the world in the post-step of the next step (reference
            # src/world.c:450-479): one priming frStepWorld on the still-empty world.
            lib.frStepWorld(self.world, DT)

    def step(self, n=1, dt=DT):
        for _ in range(n):
            self.lib.frStepWorld(self.world, dt)

    def snapshot(self):
        return snapshot(self.lib, self.bodies)

    def release(self):
        lib = self.lib
        in_world = set()
        for i in range(lib.frGetBodyCountInWorld(self.world)):
            in_world.add(lib.frGetBodyInWorld(self.world, i))
        lib.frReleaseWorld(self.world)          # frees bodies currently in the world (world.c:124-125)
        for b in self.bodies:
            if b not in in_world:
                lib.frReleaseBody(b)
        for s in self.shapes:
            lib.frReleaseShape(s)
        self.bodies, self.shapes, self.world = [], [], None


def make_shape(lib, s):
    mat = capi.Material(s.density, s.friction, s.restitution)
    if s.kind == "circle":
        return lib.frCreateCircle(mat, float(s.params[0]))
    if s.kind == "rect":
        return lib.frCreateRectangle(mat, float(s.params[0]), float(s.params[1]))
    v = capi.Vertices()
    v.count = len(s.params)
    for j, (x, y) in enumerate(s.params):
        v.data[j] = capi.Vector2(float(np.float32(x)), float(np.float32(y)))
    return lib.frCreatePolygon(mat, C.byref(v))


def snapshot(lib, bodies):
    """Dynamic state of `bodies` through the public getters (reference include/ferox.h:413-461)."""
    n = len(bodies)
    out = {
        "pos": np.empty((n, 2), np.float32), "angle": np.empty(n, np.float32),
        "sincos": np.empty((n, 2), np.float32), "vel": np.empty((n, 2), np.float32),
        "omega": np.empty(n, np.float32), "aabb": np.empty((n, 4), np.float32),
    }
    for i, b in enumerate(bodies):
        tx = lib.frGetBodyTransform(b)
        out["pos"][i] = (tx.position.x, tx.position.y)
        out["angle"][i] = tx.angle
        out["sincos"][i] = (tx.rotation.sin_, tx.rotation.cos_)
        v = lib.frGetBodyVelocity(b)
        out["vel"][i] = (v.x, v.y)
        out["omega"][i] = lib.frGetBodyAngularVelocity(b)
        a = lib.frGetBodyAABB(b)
        out["aabb"][i] = (a.x, a.y, a.width, a.height)
    return out


class ManifoldRecorder:
    """Installs a postStep (and optionally preStep) collision handler that records every
    (key, *value) the library reports -- the public window onto the contact cache in its array
    order (reference src/world.c:254-259)."""

    def __init__(self, scene):
        self.scene = scene
        self.records = []
        self._post = capi.CollisionEventFunc(self._on_post)
        handler = capi.CollisionHandler(capi.CollisionEventFunc(), self._post)
        scene.lib.frSetWorldCollisionHandler(scene.world, handler)

    def _on_post(self, key, value):
        c = value.contents
        idx = self.scene.index_of
        rec = (idx.get(key.first, -1), idx.get(key.second, -1), c.count,
               c.direction.x, c.direction.y, c.friction, c.restitution)
        for k in range(2):
            ct = c.contacts[k]
            rec += (ct.id, ct.depth, ct.point.x, ct.point.y, ct.cache.normalMass,
                    ct.cache.normalScalar, ct.cache.tangentMass, ct.cache.tangentScalar)
        self.records.append(rec)

    def take(self):
        out, self.records = self.records, []
        return out

    def remove(self):
        self.scene.lib.frSetWorldCollisionHandler(
            self.scene.world, capi.CollisionHandler(capi.CollisionEventFunc(), capi.CollisionEventFunc()))


class BatchScene:
    """Several SceneSpecs realised as ONE batch of independent sub-worlds (product only,
    frxCreateWorldBatch): a single device pipeline steps all of them with one launch set."""

    def __init__(self, lib, specs):
        self.lib, self.specs = lib, specs
        total = sum(s.n for s in specs)
        self.world = lib.frxCreateWorldBatch(capi.Vector2(*specs[0].gravity), specs[0].cell, len(specs))
        lib.frxSetWorldCapacity(self.world, total + 16)
        self.shapes, self.bodies, self.ranges = [], [], []
        for k, spec in enumerate(specs):
            shapes = [make_shape(lib, s) for s in spec.shapes]
            self.shapes += shapes
            start = len(self.bodies)
            for i in range(spec.n):
                b = lib.frCreateBodyFromShape(int(spec.body_type[i]), capi.Vector2(float(spec.pos[i, 0]), float(spec.pos[i, 1])),
                                              shapes[int(spec.body_shape[i])])
                if spec.angle[i] != 0.0:
                    lib.frSetBodyAngle(b, float(spec.angle[i]))
                if spec.vel[i, 0] != 0.0 or spec.vel[i, 1] != 0.0:
                    lib.frSetBodyVelocity(b, capi.Vector2(float(spec.vel[i, 0]), float(spec.vel[i, 1])))
                if spec.omega[i] != 0.0:
                    lib.frSetBodyAngularVelocity(b, float(spec.omega[i]))
                self.bodies.append(b)
                if not lib.frxAddBodyToBatch(self.world, k, b):
                    raise RuntimeError("frxAddBodyToBatch refused body %d of world %d" % (i, k))
            self.ranges.append((start, len(self.bodies)))

    def release(self):
        self.lib.frReleaseWorld(self.world)
        for s in self.shapes:
            self.lib.frReleaseShape(s)


class StripScene:
    """One SceneSpec cut into vertical strips, each strip a frWorld of the product (frxStrip*): dynamic
    and kinematic bodies go to the strip whose x-interval holds their position, static bodies to every
    strip.  All strips live in this process (in-process neighbour links) unless `rank` is given, in
    which case only strip `rank` is built (NCCL neighbours, one process per GPU)."""

    def __init__(self, lib, spec, cuts, margin=4.0, ghost_cap=4096, rank=None, devices=None):
        self.lib, self.spec = lib, spec
        edges = [-math.inf] + [float(c) for c in cuts] + [math.inf]
        self.n_strips = len(edges) - 1
        self.local = list(range(self.n_strips)) if rank is None else [rank]
        stride = spec.n + 64
        owner = np.searchsorted(np.asarray(cuts, np.float32), spec.pos[:, 0], side="right")
        self.worlds, self.bodies, self.owned = [], {}, {}
        for k in self.local:
            if devices is not None:
                lib.frxSetDevice(devices[k])
            w = lib.frCreateWorld(capi.Vector2(*spec.gravity), spec.cell)
            lib.frxSetWorldCapacity(w, spec.n + 16)
            rc = lib.frxStripConfigure(w, edges[k], edges[k + 1], margin, ghost_cap, k * stride, self.n_strips * stride)
            assert rc == 0, "frxStripConfigure failed"
            shapes = [make_shape(lib, s) for s in spec.shapes]
            arr = (C.c_void_p * len(shapes))(*shapes)
            assert lib.frxStripRegisterShapes(w, arr, len(shapes)) == 0
            mine = []
            for i in range(spec.n):
                static = int(spec.body_type[i]) == capi.BODY_STATIC
                if not static and owner[i] != k:
                    continue
                b = lib.frCreateBodyFromShape(int(spec.body_type[i]), capi.Vector2(float(spec.pos[i, 0]), float(spec.pos[i, 1])),
                                              shapes[int(spec.body_shape[i])])
                if spec.angle[i] != 0.0:
                    lib.frSetBodyAngle(b, float(spec.angle[i]))
                if spec.vel[i, 0] != 0.0 or spec.vel[i, 1] != 0.0:
                    lib.frSetBodyVelocity(b, capi.Vector2(float(spec.vel[i, 0]), float(spec.vel[i, 1])))
                if spec.omega[i] != 0.0:
                    lib.frSetBodyAngularVelocity(b, float(spec.omega[i]))
                assert lib.frAddBodyToWorld(w, b)
                self.bodies[(k, i)] = b
                mine.append(i)
            self.worlds.append(w)
            self.owned[k] = np.array(mine, np.int64)
        if rank is None:
            for a, b in zip(self.worlds[:-1], self.worlds[1:]):
                assert lib.frxStripLinkLocal(a, b) == 0
        self._arr = (C.c_void_p * len(self.worlds))(*self.worlds)

    @classmethod
    def local(cls, lib, spec, k, n_strips, cut_lo, cut_hi, margin=4.0, ghost_cap=4096, device=None):
        """Strip k only, from a spec that already holds just this strip's bodies (strip_pile); bodies
        are created in bulk.  Neighbours are linked by the caller (frxStripInitNccl)."""
        self = cls.__new__(cls)
        self.lib, self.spec, self.n_strips, self.local = lib, spec, n_strips, [k]
        if device is not None:
            lib.frxSetDevice(device)
        stride = 2 * spec.n + 64          # id slice of a strip: room for as many immigrants (frxStripMigrate) as it has bodies
        w = lib.frCreateWorld(capi.Vector2(*spec.gravity), spec.cell)
        lib.frxSetWorldCapacity(w, 2 * spec.n + 16)
        assert lib.frxStripConfigure(w, cut_lo, cut_hi, margin, ghost_cap, k * stride, n_strips * stride) == 0
        shapes = [make_shape(lib, s) for s in spec.shapes]
        arr = (C.c_void_p * len(shapes))(*shapes)
        assert lib.frxStripRegisterShapes(w, arr, len(shapes)) == 0
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        bt, bs = np.ascontiguousarray(spec.body_type, np.int32), np.ascontiguousarray(spec.body_shape, np.int32)
        pos, ang = np.ascontiguousarray(spec.pos, np.float32), np.ascontiguousarray(spec.angle, np.float32)
        got = lib.frxCreateBodiesEx(w, spec.n, bt, bs, arr, pos.reshape(-1), ptr(ang), None, None, None, None)
        assert got == spec.n, (got, spec.n)
        self.worlds, self.bodies, self.owned = [w], {}, {k: np.arange(spec.n, dtype=np.int64)}
        self._arr = (C.c_void_p * 1)(w)
        return self

    def step(self, n=1, dt=DT):
        for _ in range(n):
            self.lib.frxStripStep(self._arr, len(self.worlds), dt)

    def gather_states(self):
        """body state of every non-static body, indexed like the spec (owned strips only)."""
        import ferox_b200 as fb
        n = self.spec.n
        out = {"pos": np.full((n, 2), np.nan, np.float32), "vel": np.full((n, 2), np.nan, np.float32),
               "angle": np.full(n, np.nan, np.float32), "omega": np.full(n, np.nan, np.float32)}
        for w, k in zip(self.worlds, self.local):
            st = fb.body_states(self.lib, w)
            idx = self.owned[k]
            for f in out:
                out[f][idx] = st[f]
        return out

    def release(self):
        for w in self.worlds:
            self.lib.frReleaseWorld(w)
