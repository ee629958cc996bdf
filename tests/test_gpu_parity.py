"""GPU parity tests: the CUDA world step, called through the C ABI, against the CPU oracle
(oracle/step_oracle.c), the unmodified reference (oracle/_ref) and the golden vectors.

Levels (BASELINE.json north_star):
  1. broadphase pair sets, SAT contact counts and feature ids ........ bit-exact
  2. contact points, normals, depths ................................. bit-exact here (<= 1e-5 rel asked)
  3. post-step state: reference-order solver mode .................... bit-exact, many steps
     graph-coloured mode (different Gauss-Seidel order) .............. exact where manifolds are
        independent; one step from identical state within the ordering tolerance of SURVEY App. A.13;
        1,000-step energy / penetration / contact statistics within 1-2 %
"""
import ctypes as C
import os

import numpy as np
import pytest

import ferox_b200 as fb
from ferox_b200 import capi, scenes
from oracle import orc
from tests.helpers import (RefCallLog, assert_bits_equal, assert_caches_equal, assert_states_equal, random_collision,
                           records_to_entries)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DT = scenes.DT


def make_world(L, spec, mode):
    sc = scenes.Scene(L, spec, prime=False)
    L.frxSetWorldSolverMode(sc.world, mode)
    L.frStepWorld(sc.world, DT)                      # priming step: queued bodies become members
    assert L.frGetBodyCountInWorld(sc.world) == spec.n
    fb.check_error(L)
    return sc


# ------------------------------------------------------------------------------------------------
# the reference's own unit tests, run against the product (tests/src/collision.c:46-201)
# ------------------------------------------------------------------------------------------------
EPS = float(np.finfo(np.float32).eps)


def test_reference_ut_circle_vs_circle(gpu):
    L = gpu
    zero = capi.Material(0, 0, 0)
    s1, s2 = L.frCreateCircle(zero, 1.0), L.frCreateCircle(zero, 1.15)
    b1 = L.frCreateBodyFromShape(capi.BODY_KINEMATIC, capi.Vector2(0, 0), s1)
    b2 = L.frCreateBodyFromShape(capi.BODY_KINEMATIC, capi.Vector2(0, 0), s2)
    col = capi.Collision()
    for (x1, x2, dirx, diry, depth, px, py) in [(-0.5, 1.65, 1, 0, 0.0, 0.5, 0.0), (-0.5, 1.5, 1, 0, 0.15, 0.5, 0.0),
                                               (0.0, 0.0, 0, 1, 2.15, 0.0, 1.0)]:
        L.frSetBodyPosition(b1, capi.Vector2(x1, 0))
        L.frSetBodyPosition(b2, capi.Vector2(x2, 0))
        L.frComputeCollision(b1, b2, C.byref(col))
        assert col.count == 1
        assert abs(col.direction.x - dirx) <= EPS and abs(col.direction.y - diry) <= EPS
        assert abs(col.contacts[0].depth - depth) <= EPS
        assert abs(col.contacts[0].point.x - px) <= EPS and abs(col.contacts[0].point.y - py) <= EPS


def test_reference_ut_circle_vs_polygon(gpu):
    L = gpu
    zero = capi.Material(0, 0, 0)
    s1, s2 = L.frCreateCircle(zero, 1.0), L.frCreateRectangle(zero, 2.0, 2.0)
    b1 = L.frCreateBodyFromShape(capi.BODY_KINEMATIC, capi.Vector2(0, 0), s1)
    b2 = L.frCreateBodyFromShape(capi.BODY_KINEMATIC, capi.Vector2(0, 0), s2)
    col = capi.Collision()
    L.frSetBodyPosition(b1, capi.Vector2(-1.5, 0)); L.frSetBodyPosition(b2, capi.Vector2(1.5, 0))
    assert not L.frComputeCollision(b1, b2, C.byref(col)) and col.count == 0      # a miss leaves *collision alone
    for (x1, x2, depth, px) in [(-1.0, 1.0, 0.0, 0.0), (-0.5, 0.5, 1.0, 0.5)]:
        L.frSetBodyPosition(b1, capi.Vector2(x1, 0)); L.frSetBodyPosition(b2, capi.Vector2(x2, 0))
        assert L.frComputeCollision(b1, b2, C.byref(col)) and col.count == 1
        assert abs(col.direction.x - 1.0) <= EPS and abs(col.direction.y) <= EPS
        assert abs(col.contacts[0].depth - depth) <= EPS and abs(col.contacts[0].point.x - px) <= EPS
    L.frSetBodyPosition(b1, capi.Vector2(0, 0)); L.frSetBodyPosition(b2, capi.Vector2(1.01, 1.01))
    assert L.frComputeCollision(b1, b2, C.byref(col)) and col.count == 1
    d = np.float32(1.01) / np.sqrt(np.float32(1.01) ** 2 * 2)
    assert abs(col.direction.x - d) <= EPS and abs(col.direction.y - d) <= EPS
    assert abs(col.contacts[0].depth - (1.0 - float(np.sqrt(np.float32(0.01) ** 2 * 2)))) <= EPS


def test_reference_test_binary_relinked(gpu):
    """The reference's own unit-test PROGRAM (tests/src/*.c compiled against the reference's ferox.h)
    linked against libferox.so instead of the reference library (oracle/Makefile `relink`)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "ferox_tests_relinked.out")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ferox_tests_relinked.out not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Pass: 6, fail: 0" in out.stdout, out.stdout[-600:]


def test_known_answer_two_boxes(gpu):
    """two 2x1 boxes at (0,0) and (0.5,0.9)."""
    L = gpu
    g = np.load(os.path.join(GOLDEN, "known_answers.npz"))
    rect = L.frCreateRectangle(capi.Material(1.0, 0.5, 0.0), 2.0, 1.0)
    b1 = L.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.0, 0.0), rect)
    b2 = L.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.5, 0.9), rect)
    col = capi.Collision()
    assert L.frComputeCollision(b1, b2, C.byref(col))
    assert bytes(col) == g["two_box"].tobytes()


# ------------------------------------------------------------------------------------------------
# level 1 + 2: narrow phase, bit for bit, on the candidate pairs of settled scenes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("make,steps", [(lambda: scenes.mixed_pile(1500, width=60.0), 250), (lambda: scenes.straddle_origin(1500), 60),
                                        (lambda: scenes.box_stack(64, 6), 80)])
def test_narrowphase_bit_exact_on_candidates(gpu, make, steps):
    L = gpu
    sc = make_world(L, make(), fb.SOLVER_COLORED)
    L.frxStepWorldN(sc.world, DT, steps)
    ow = orc.OracleWorld.from_scene(sc)              # same positions, angles, shapes
    ow.step(1)                                       # gives the candidate list of this configuration
    cands = ow.candidates()
    ow2 = orc.OracleWorld.from_scene(sc)
    pairs = np.ascontiguousarray(cands[:, :2])
    out = (capi.Collision * len(pairs))()
    hit = np.zeros(len(pairs), np.int32)
    nh = L.frxComputeCollisions(sc.world, pairs.reshape(-1), len(pairs), C.byref(out), hit.ctypes.data_as(C.c_void_p))
    fb.check_error(L)
    got = np.frombuffer(out, dtype=orc.COLLISION_DTYPE)
    assert (hit == cands[:, 2]).all(), "hit/miss decisions differ"
    assert nh == int(cands[:, 2].sum()) and nh > len(pairs) // 10
    kinds = set()
    for i in np.nonzero(hit)[0]:
        _, want = ow2.collide(int(pairs[i, 0]), int(pairs[i, 1]))
        p = got[i]
        assert p["count"] == want["count"] and p["dir"].tobytes() == want["dir"].tobytes(), (i, p, want)
        for k in range(2):
            for f in ("id", "depth", "point"):
                assert p["contacts"][k][f].tobytes() == want["contacts"][k][f].tobytes(), (i, k, f, p, want)
        kinds.add(int(want["count"]) * 10 + (1 if want["contacts"][0]["id"] else 0))
    sc.release(); ow.close(); ow2.close()


# ------------------------------------------------------------------------------------------------
# level 1: broadphase candidate set == the reference's narrow-phase call set
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cell", [2.0, 1.0, 0.7])
def test_broadphase_pair_sequence_matches_reference(gpu, ref, cell):
    """The narrow-phase call SEQUENCE of the unmodified reference (lexicographic (i, j), hit flags)
    on a scene straddling x = 0 and y = 0, for three cell sizes, over several steps."""
    spec = scenes.straddle_origin(1200)
    spec.cell = cell
    L = gpu
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    rs = scenes.Scene(ref, spec)
    with RefCallLog(ref) as log:
        for s in range(4):
            log.reset()
            rs.step(1)
            L.frStepWorld(sc.world, DT)
            want, got = log.calls(rs), fb.candidate_pairs(L, sc.world)
            assert got.shape == want.shape and (got == want).all(), "step %d" % s
    assert fb.step_counters(L, sc.world)["overflow"] == 0
    # the production ordering emits the same SET (cell-major order instead of lexicographic)
    sc2 = make_world(L, spec, fb.SOLVER_COLORED)
    rs2 = scenes.Scene(ref, spec)
    with RefCallLog(ref) as log:
        rs2.step(1)
        L.frStepWorld(sc2.world, DT)
        assert sorted(map(tuple, fb.candidate_pairs(L, sc2.world))) == sorted(map(tuple, log.calls(rs2)))
    sc.release(); rs.release(); sc2.release(); rs2.release()


def test_broadphase_large_static_body_and_negative_cells(gpu):
    """A ground spanning thousands of cells (expanded by a whole block) and bodies on both sides of
    the truncate-toward-zero 'cell 0' produce exactly the oracle's pair set and entry count K."""
    spec = scenes.straddle_origin(900)
    spec.shapes[0] = scenes.ShapeSpec("rect", (6000.0, 2.0))
    L = gpu
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    L.frxStepWorldN(sc.world, DT, 5)
    ow = orc.OracleWorld.from_scene(sc)
    L.frStepWorld(sc.world, DT)
    ow.step(1)
    got, want = fb.candidate_pairs(L, sc.world), ow.candidates()
    assert sorted(map(tuple, got)) == sorted(map(tuple, want))
    c = fb.step_counters(L, sc.world)
    assert c["cellEntries"] == ow.counters()["K"] and c["cellEntries"] > 3000
    sc.release(); ow.close()


# ------------------------------------------------------------------------------------------------
# level 3a: reference-order solver mode is bit-exact, step after step
# ------------------------------------------------------------------------------------------------
SERIAL_SCENES = [
    ("box_stack_32x4", lambda: scenes.box_stack(32, 4), 120),
    ("straddle_origin_300", lambda: scenes.straddle_origin(300), 80),
    ("mixed_pile_400", lambda: scenes.mixed_pile(400, width=40.0), 150),
    ("small_world_5", lambda: scenes.small_world(5), 150),
]


@pytest.mark.parametrize("name,make,steps", SERIAL_SCENES, ids=[s[0] for s in SERIAL_SCENES])
def test_serial_mode_matches_golden_reference_run(gpu, name, make, steps):
    """Whole trajectories against the frozen outputs of the unmodified reference: body state, the
    last step's candidate SEQUENCE and the contact cache in array order, all bit for bit."""
    L = gpu
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = make_world(L, make(), fb.SOLVER_SERIAL)
    st = fb.body_states(L, sc.world)
    for k in ("pos", "angle", "sincos", "vel", "omega", "aabb"):
        assert_bits_equal(st[k], g["init_" + k], "initial " + k)
    L.frxStepWorldN(sc.world, DT, int(g["steps"]))
    fb.check_error(L)
    st = fb.body_states(L, sc.world)
    for k in ("pos", "angle", "sincos", "vel", "omega", "aabb"):
        assert_bits_equal(st[k], g["final_" + k], "final " + k)
    assert (fb.candidate_pairs(L, sc.world) == g["last_candidates"]).all()
    m = fb.manifolds(L, sc.world)
    assert_caches_equal(g["last_cache"], m[m["count"] > 0], name)
    sc.release()


def test_serial_mode_config1_600_steps_vs_live_reference(gpu, ref):
    """BASELINE config 1 (1,024 boxes on static ground, dt = 1/60, 10 iterations, 600 steps) next to
    the compiled reference: identical bits at the end, and identical KE / contact counts."""
    L = gpu
    spec = scenes.box_stack(128, 8)
    rs = scenes.Scene(ref, spec)
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    for chunk in range(6):
        rs.step(100)
        L.frxStepWorldN(sc.world, DT, 100)
        assert_states_equal(rs.snapshot(), fb.body_states(L, sc.world), "config 1 after %d steps" % (100 * (chunk + 1)))
    c = fb.step_counters(L, sc.world)
    assert c["manifolds"] == 1024 and c["overflow"] == 0
    fb.check_error(L)
    sc.release(); rs.release()


def test_serial_mode_eviction_rule_with_world_clock(gpu):
    """Stale entries are evicted by the timestamp rule (src/world.c:222-235) when the world clock
    moves, and kept when it does not (frStepWorld never advances it)."""
    L = gpu
    spec = scenes.mixed_pile(300, width=30.0)
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    ow = orc.OracleWorld.from_scene(sc)
    t = 0.0
    for s in range(120):
        if s % 7 == 6:
            t += 0.05                               # a "frame" longer than dt: everything stale goes
        L.frxSetWorldTimestamp(sc.world, t)
        ow.set_timestamp(t)
        L.frStepWorld(sc.world, DT)
        ow.step(1)
    assert_states_equal(fb.body_states(L, sc.world), ow.state(), "eviction run")
    sc.release(); ow.close()


# ------------------------------------------------------------------------------------------------
# level 3b: graph-coloured (production) mode
# ------------------------------------------------------------------------------------------------
def test_colored_mode_is_exact_when_manifolds_are_independent(gpu):
    """Separate boxes resting on static ground share no dynamic body: any Gauss-Seidel order gives
    the reference's bits, so the coloured solver must equal the oracle exactly."""
    L = gpu
    spec = scenes.box_stack(96, 1)
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    ow = orc.OracleWorld.from_scene(sc)
    L.frxStepWorldN(sc.world, DT, 200)
    ow.step(200)
    assert_states_equal(fb.body_states(L, sc.world), ow.state(), "independent manifolds")
    m, oc = fb.manifolds(L, sc.world), ow.cache()
    key = lambda a: np.argsort(a["a"].astype(np.int64) * 100000 + a["b"])
    assert_caches_equal(m[key(m)], oc[key(oc)], "cache (sorted by pair)")
    assert fb.step_counters(L, sc.world)["colors"] == 1
    sc.release(); ow.close()


def test_colored_mode_single_step_within_ordering_tolerance(gpu):
    """From bit-identical state (reached in reference-order mode) one coloured step differs from the
    reference-order step only by the Gauss-Seidel ordering effect.  The yardstick is MEASURED here, on the
    reference's own arithmetic: a second oracle takes the same step with its sweeps re-ordered colour-major
    (oracle/step_oracle.c `solve_order`: same contacts, same sweeps, another visiting order -- what SURVEY App. A.13
    probed), a third in a random order.  The GPU step must stay within 1.5x of the larger re-ordering effect (rms and max |dv|, max |dx|); pair set
    and contact geometry stay bit-exact (they are computed before the solver runs)."""
    L = gpu
    spec = scenes.mixed_pile(2000, width=70.0)
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    ow = orc.OracleWorld.from_scene(sc)
    ow2, ow3 = orc.OracleWorld.from_scene(sc), orc.OracleWorld.from_scene(sc)
    L.frxStepWorldN(sc.world, DT, 300)
    for o in (ow, ow2, ow3):
        o.step(300)
    assert_states_equal(fb.body_states(L, sc.world), ow.state(), "serial prefix")
    L.frxSetWorldSolverMode(sc.world, fb.SOLVER_COLORED)
    L.frStepWorld(sc.world, DT)
    ow.step(1)
    ow2.set_solve_order("colored")
    ow2.step(1)
    ow3.set_solve_order("random", 11)
    ow3.step(1)
    got, want = fb.body_states(L, sc.world), ow.state()
    assert sorted(map(tuple, fb.candidate_pairs(L, sc.world))) == sorted(map(tuple, ow.candidates()))
    m, oc = fb.manifolds(L, sc.world), ow.cache()
    key = lambda a: np.argsort(a["a"].astype(np.int64) * 100000 + a["b"])
    m, oc = m[key(m)], oc[key(oc)]
    assert len(m) == len(oc)
    for f in ("a", "b", "count", "dir"):
        assert m[f].tobytes() == oc[f].tobytes(), f
    for f in ("id", "depth", "point"):
        assert np.ascontiguousarray(m["contacts"][f]).tobytes() == np.ascontiguousarray(oc["contacts"][f]).tobytes(), f

    def effect(a):
        dv = np.linalg.norm(a["vel"] - want["vel"], axis=1)
        dx = np.linalg.norm(a["pos"] - want["pos"], axis=1)
        return float(np.sqrt((dv ** 2).mean())), float(dv.max()), float(dx.max())
    (g_rms, g_max, g_dx) = effect(got)
    o_rms, o_max, o_dx = (max(x) for x in zip(effect(ow2.state()), effect(ow3.state())))   # colour-major and random re-orderings
    assert o_rms > 0.0, "the yardstick step did not differ from the reference order at all"
    assert g_rms <= 1.5 * o_rms and g_max <= 1.5 * o_max + 0.05 and g_dx <= 1.5 * o_dx + 1e-3, \
        ("gpu", g_rms, g_max, g_dx, "re-ordered reference", o_rms, o_max, o_dx)
    c = fb.step_counters(L, sc.world)
    assert 2 <= c["colors"] <= 24
    sc.release(); ow.close(); ow2.close(); ow3.close()


def test_colored_mode_long_run_statistics(gpu):
    """1,000 steps (means of the last 200).  north_star asks for energy and penetration statistics within 1 %.  How
    far a pure RE-ORDERING of the reference's own sweeps moves each statistic is measured in this test: two more
    oracles run the same scene colour-major and in a fresh random order every step (oracle/step_oracle.c
    `solve_order`).  [measured, CPU] on this scene the re-ordered reference itself differs from the reference by
    ~0.3-0.5 % in potential energy, 1.5-2 % in mean penetration, 2-2.5 % in contact count and 25-50 % in residual
    kinetic energy (the chaotic jitter of a "resting" pile, App. A.13).  Asserted: potential energy within 1 %
    outright; mean penetration and contact count within max(1 %, 1.5 x the largest re-ordering spread measured
    here); kinetic energy within 3 x the largest re-ordered value."""
    L = gpu
    spec = scenes.mixed_pile(2500, width=80.0)
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    ows = []
    for mode in ("reference", "colored", "random"):
        ow = orc.OracleWorld.from_scene(sc)
        ow.set_solve_order(mode, 7)
        ows.append(ow)
    masses = np.array([L.frGetBodyMass(b) for b in sc.bodies], np.float64)
    runs = [dict(pe=[], pen=[], cnt=[], ke=[]) for _ in range(4)]       # gpu, reference order, colour-major, random
    for s in range(1000):
        L.frStepWorld(sc.world, DT)
        for ow in ows:
            ow.step(1)
        if s >= 800:
            views = [(fb.body_states(L, sc.world), fb.manifolds(L, sc.world))] + [(ow.state(), ow.cache()) for ow in ows]
            for r, (st, cache) in zip(runs, views):
                r["pe"].append(-(masses * 9.8 * st["pos"][:, 1]).sum())
                r["ke"].append(0.5 * (masses * (st["vel"].astype(np.float64) ** 2).sum(1)).sum())
                d = np.concatenate([cache["contacts"]["depth"][:, 0], cache["contacts"]["depth"][cache["count"] > 1, 1]])
                r["pen"].append(d.mean())
                r["cnt"].append(int(cache["count"].sum()))
    fb.check_error(L)
    mean = lambda j, k: float(np.mean(runs[j][k]))
    rel = lambda j, k: abs(mean(j, k) - mean(1, k)) / abs(mean(1, k))
    spread = {k: max(rel(2, k), rel(3, k)) for k in ("pe", "pen", "cnt")}
    assert rel(0, "pe") <= 0.01, ("potential energy", rel(0, "pe"), spread)
    assert rel(0, "pen") <= max(0.01, 1.5 * spread["pen"]), ("mean penetration", rel(0, "pen"), spread)
    assert rel(0, "cnt") <= max(0.01, 1.5 * spread["cnt"]), ("contact count", rel(0, "cnt"), spread)
    assert mean(0, "ke") <= 3.0 * max(mean(1, "ke"), mean(2, "ke"), mean(3, "ke")) + 1.0
    sc.release()
    for ow in ows:
        ow.close()


def test_colored_mode_is_deterministic(gpu):
    L = gpu
    runs = []
    for _ in range(2):
        sc = make_world(L, scenes.mixed_pile(1500, width=60.0), fb.SOLVER_COLORED)
        L.frxStepWorldN(sc.world, DT, 150)
        runs.append(fb.body_states(L, sc.world))
        sc.release()
    assert_states_equal(runs[0], runs[1], "two runs")


# ------------------------------------------------------------------------------------------------
# public per-pair / per-body entry points (they run the same device functions)
# ------------------------------------------------------------------------------------------------
def test_resolve_and_accumulated_impulses_match_reference(gpu, ref):
    L = gpu
    mk = lambda lib: scenes.Scene(lib, scenes.box_stack(2, 1), prime=False)
    ra, pa = mk(ref), mk(L)
    rng = np.random.default_rng(1)
    for trial in range(300):
        for lib, sc in ((ref, ra), (L, pa)):
            rs = np.random.default_rng(trial)
            for b in (sc.bodies[1], sc.bodies[2]):
                v = rs.normal(size=2) * rs.choice([0, 1, 1e-3])
                lib.frSetBodyVelocity(b, capi.Vector2(float(np.float32(v[0])), float(np.float32(v[1]))))
                lib.frSetBodyAngularVelocity(b, float(np.float32(rs.normal() * rs.choice([0, 1]))))
        col = random_collision(rng)
        c1, c2 = capi.Collision.from_buffer_copy(bytes(col)), capi.Collision.from_buffer_copy(bytes(col))
        if trial % 2:
            ref.frResolveCollision(ra.bodies[1], ra.bodies[2], C.byref(c1), 60.0)
            L.frResolveCollision(pa.bodies[1], pa.bodies[2], C.byref(c2), 60.0)
        else:
            ref.frApplyAccumulatedImpulses(ra.bodies[1], ra.bodies[2], C.byref(c1))
            L.frApplyAccumulatedImpulses(pa.bodies[1], pa.bodies[2], C.byref(c2))
        assert bytes(c1) == bytes(c2), trial
        a, b = scenes.snapshot(ref, ra.bodies[1:3]), scenes.snapshot(L, pa.bodies[1:3])
        assert a["vel"].tobytes() == b["vel"].tobytes() and a["omega"].tobytes() == b["omega"].tobytes(), trial
    fb.check_error(L)


def test_integrate_entry_points_match_reference(gpu, ref):
    L = gpu
    rng = np.random.default_rng(2)
    for trial in range(60):
        vals = rng.normal(size=8)
        out = []
        for lib in (ref, L):
            s = lib.frCreateRectangle(capi.Material(1.0, 0.5, 0.0), 1.5, 0.8) if trial % 2 else lib.frCreateCircle(capi.Material(1.0, 0.5, 0.0), 0.6)
            b = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(float(np.float32(vals[0])), float(np.float32(vals[1]))), s)
            lib.frSetBodyAngle(b, float(np.float32(vals[2] * 3)))
            lib.frSetBodyVelocity(b, capi.Vector2(float(np.float32(vals[3])), float(np.float32(vals[4]))))
            lib.frSetBodyAngularVelocity(b, float(np.float32(vals[5] * 4)))
            lib.frApplyForceToBody(b, capi.Vector2(0.3, 0.1), capi.Vector2(float(np.float32(vals[6])), float(np.float32(vals[7]))))
            lib.frIntegrateForBodyVelocity(b, DT)
            lib.frIntegrateForBodyPosition(b, DT)
            out.append(scenes.snapshot(lib, [b]))
        assert_states_equal(out[0], out[1], "trial %d" % trial)


def test_device_sincos_matches_host_libm_on_integrator_range(gpu):
    """Every float in [pi, 3*pi] (the range frNormalizeAngle produces, ~13.5 M values) through the
    device fx_sincosf against the host's sinf/cosf of the same glibc, bit for bit."""
    L = gpu
    libm = C.CDLL("libm.so.6")
    lo = np.float32(np.pi).view(np.uint32)
    hi = np.float32(3 * np.pi).view(np.uint32)
    a = np.arange(int(lo) - 64, int(hi) + 64, dtype=np.uint32).view(np.float32)
    s, c = np.empty_like(a), np.empty_like(a)
    assert L.frxDeviceSinCos(np.ascontiguousarray(a), len(a), s, c) == len(a)
    # vectorised host evaluation through numpy would use another libm path: call glibc directly
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import subprocess, tempfile
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "h.c")
    open(src, "w").write("#include <math.h>\nvoid f(const float*a,long n,float*s,float*c){for(long i=0;i<n;i++){s[i]=sinf(a[i]);c[i]=cosf(a[i]);}}\n")
    so = os.path.join(tmp, "h.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    H = C.CDLL(so)
    hs, hc = np.empty_like(a), np.empty_like(a)
    H.f(a.ctypes.data_as(C.c_void_p), C.c_long(len(a)), hs.ctypes.data_as(C.c_void_p), hc.ctypes.data_as(C.c_void_p))
    assert (s.view(np.uint32) == hs.view(np.uint32)).all() and (c.view(np.uint32) == hc.view(np.uint32)).all()
    _ = libm, ROOT


# ------------------------------------------------------------------------------------------------
# boundary behaviour: deferred add/remove, handlers, capacity
# ------------------------------------------------------------------------------------------------
def test_deferred_add_remove_and_renumbering_match_reference(gpu, ref):
    """Bodies enter/leave in the post-step of the NEXT step, removal is swap-with-last
    (src/world.c:450-479); same indices, same trajectories as the reference afterwards."""
    L = gpu
    spec = scenes.mixed_pile(120, width=24.0)
    worlds = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        assert lib.frGetBodyCountInWorld(sc.world) == 0
        lib.frStepWorld(sc.world, DT)
        assert lib.frGetBodyCountInWorld(sc.world) == spec.n
        worlds.append(sc)
    for s in range(90):
        for lib, sc in zip((ref, L), worlds):
            if s in (20, 21, 40):
                victim = sc.bodies[10 + s]
                assert lib.frRemoveBodyFromWorld(sc.world, victim)
                assert lib.frIsBodyInWorld(sc.world, victim)            # still a member until the post-step
            if s == 30:
                lib.frAddBodyToWorld(sc.world, sc.bodies[30])            # re-add a removed body
            lib.frStepWorld(sc.world, DT)
        order = [[sc.index_of[lib.frGetBodyInWorld(sc.world, i)] for i in range(lib.frGetBodyCountInWorld(sc.world))]
                 for lib, sc in zip((ref, L), worlds)]
        assert order[0] == order[1], "membership / renumbering at step %d" % s
    members = [b for b in worlds[0].bodies if ref.frIsBodyInWorld(worlds[0].world, b)]
    idx = [worlds[0].index_of[b] for b in members]
    a = scenes.snapshot(ref, [worlds[0].bodies[i] for i in idx])
    b = scenes.snapshot(L, [worlds[1].bodies[i] for i in idx])
    # the reference keeps solving cache entries of removed bodies (dangling pointers, App. A.7); we
    # drop them, so only near-equality can be asked once a removed body had contacts
    assert np.abs(a["pos"] - b["pos"]).max() < 0.2
    fb.check_error(L)


def test_collision_handlers_see_the_cache_and_edits_persist(gpu):
    L = gpu
    spec = scenes.box_stack(8, 2)
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    seen = {"pre": 0, "post": 0, "pairs": set()}

    def pre(key, value):
        seen["pre"] += 1
        c = value.contents
        assert c.count > 0
        c.friction = 0.125                           # an edit must persist (src/world.c:366-367)

    def post(key, value):
        seen["post"] += 1
        seen["pairs"].add((sc.index_of.get(key.first), sc.index_of.get(key.second)))
        assert abs(value.contents.friction - 0.125) < 1e-7

    h = capi.CollisionHandler(capi.CollisionEventFunc(pre), capi.CollisionEventFunc(post))
    L.frSetWorldCollisionHandler(sc.world, h)
    for _ in range(40):
        L.frStepWorld(sc.world, DT)
    fb.check_error(L)
    L.frSetWorldCollisionHandler(sc.world, capi.CollisionHandler(capi.CollisionEventFunc(), capi.CollisionEventFunc()))
    L.frStepWorld(sc.world, DT)
    m = fb.manifolds(L, sc.world)
    assert seen["pre"] > 0 and seen["post"] > 0 and len(m) > 0
    assert (np.abs(m["friction"] - 0.125) < 1e-7).all()
    assert all(a is not None and b is not None and a < b for a, b in seen["pairs"])
    sc.release()


def test_mid_simulation_setters_and_getters(gpu, ref):
    """Setters between steps (teleport, velocity kick, type change) and lazy getters."""
    L = gpu
    spec = scenes.mixed_pile(200, width=30.0)
    ws = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        lib.frStepWorld(sc.world, DT)
        ws.append(sc)
    for s in range(60):
        for lib, sc in zip((ref, L), ws):
            if s == 10:
                lib.frSetBodyVelocity(sc.bodies[50], capi.Vector2(3.0, -8.0))
                lib.frSetBodyAngularVelocity(sc.bodies[51], 5.0)
            if s == 20:
                p = lib.frGetBodyPosition(sc.bodies[60])
                lib.frSetBodyPosition(sc.bodies[60], capi.Vector2(p.x + 0.5, p.y - 3.0))
                lib.frSetBodyAngle(sc.bodies[61], 1.25)
            if s == 30:
                lib.frSetBodyType(sc.bodies[70], capi.BODY_STATIC)
                lib.frApplyImpulseToBody(sc.bodies[71], lib.frGetBodyPosition(sc.bodies[71]), capi.Vector2(1.0, -4.0))
            if s == 35:
                lib.frSetBodyGravityScale(sc.bodies[72], 0.25)
                lib.frApplyForceToBody(sc.bodies[73], lib.frGetBodyPosition(sc.bodies[73]), capi.Vector2(20.0, 0.0))
            lib.frStepWorld(sc.world, DT)
    assert_states_equal(ws[0].snapshot(), ws[1].snapshot(), "after setters")
    fb.check_error(L)


def test_world_raycast_on_device_state(gpu, ref):
    L = gpu
    spec = scenes.mixed_pile(150, width=26.0)
    hits = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        lib.frStepWorld(sc.world, DT)
        sc.step(50)
        got = []
        cb = capi.RaycastQueryFunc(lambda hit, ctx: got.append((sc.index_of.get(hit.body), round(hit.distance, 5))))
        lib.frComputeWorldRaycast(sc.world, capi.Ray(capi.Vector2(1.0, -30.0), capi.Vector2(0.2, 1.0), 60.0), cb, None)
        hits.append(sorted(got))
    assert hits[0] == hits[1] and len(hits[0]) > 0


def test_capacity_growth_and_counters_at_20k_bodies(gpu):
    L = gpu
    spec = scenes.mixed_pile(20000)
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    L.frxStepWorldN(sc.world, DT, 400)
    c = fb.step_counters(L, sc.world)
    fb.check_error(L)
    assert c["overflow"] == 0 and c["bodies"] == spec.n
    assert 1.5 * spec.n < c["cellEntries"] < 4 * spec.n and c["manifolds"] > spec.n
    ow = orc.OracleWorld.from_scene(sc)
    L.frStepWorld(sc.world, DT)
    ow.step(1)
    assert sorted(map(tuple, fb.candidate_pairs(L, sc.world))) == sorted(map(tuple, ow.candidates()))
    sc.release(); ow.close()


# ------------------------------------------------------------------------------------------------
# batched independent worlds (BASELINE config 4): one launch set, per-world results unchanged
# ------------------------------------------------------------------------------------------------
def test_batched_worlds_equal_separate_worlds_bit_for_bit(gpu):
    """24 small worlds (different seeds) stepped as ONE batch in reference-order mode: every
    sub-world's bodies end up with exactly the bits the CPU oracle gives that world alone, and no
    candidate pair ever crosses a sub-world boundary."""
    L = gpu
    specs = [scenes.small_world(k, n_dynamic=60) for k in range(24)]
    oracles = []
    for spec in specs:
        sc = scenes.Scene(L, spec, prime=False)          # host-side set-up only: the oracle's input
        oracles.append(orc.OracleWorld.from_scene(sc))
    batch = scenes.BatchScene(L, specs)
    L.frxSetWorldSolverMode(batch.world, fb.SOLVER_SERIAL)
    L.frStepWorld(batch.world, DT)
    assert L.frGetBodyCountInWorld(batch.world) == sum(s.n for s in specs)
    steps = 120
    L.frxStepWorldN(batch.world, DT, steps)
    fb.check_error(L)
    st = fb.body_states(L, batch.world)
    for k, (ow, (a, b)) in enumerate(zip(oracles, batch.ranges)):
        ow.step(steps)
        want = ow.state()
        for f in ("pos", "angle", "sincos", "vel", "omega", "aabb"):
            assert_bits_equal(st[f][a:b], want[f], "world %d %s" % (k, f))
    cands = fb.candidate_pairs(L, batch.world)
    owner = np.zeros(len(batch.bodies), np.int32)
    for k, (a, b) in enumerate(batch.ranges):
        owner[a:b] = k
    assert (owner[cands[:, 0]] == owner[cands[:, 1]]).all()
    assert len(cands) == sum(len(ow.candidates()) for ow in oracles)
    c = fb.step_counters(L, batch.world)
    assert c["overflow"] == 0
    batch.release()


def test_batched_worlds_colored_mode_runs_many_worlds(gpu):
    L = gpu
    specs = [scenes.small_world(k) for k in range(256)]
    batch = scenes.BatchScene(L, specs)
    L.frStepWorld(batch.world, DT)
    L.frxStepWorldN(batch.world, DT, 300)
    fb.check_error(L)
    st = fb.body_states(L, batch.world)
    c = fb.step_counters(L, batch.world)
    assert np.isfinite(st["pos"]).all() and c["overflow"] == 0 and c["bodies"] == 256 * 256
    assert c["manifolds"] > 256 * 100
    batch.release()


def test_state_views_equal_the_copying_getter(gpu):
    """frxViewBodyStates exposes the pinned host mirror itself: same numbers as frxGetBodyStates and the
    per-body getters, refreshed after every step."""
    L = gpu
    spec = scenes.mixed_pile(2000, width=80.0)
    sc = scenes.Scene(L, spec, prime=False)
    L.frStepWorld(sc.world, DT)
    for _ in range(3):
        L.frxStepWorldN(sc.world, DT, 40)
        v = fb.body_state_views(L, sc.world)
        st = fb.body_states(L, sc.world)
        assert np.array_equal(v["posrot"][:, :2], st["pos"]) and np.array_equal(v["posrot"][:, 2:], st["sincos"])
        assert np.array_equal(v["vel"][:, :2], st["vel"]) and np.array_equal(v["vel"][:, 2], st["omega"])
        assert np.array_equal(v["angle"], st["angle"]) and np.array_equal(v["aabb"], st["aabb"])
        p = L.frGetBodyPosition(sc.bodies[7])
        assert (p.x, p.y) == (float(v["posrot"][7, 0]), float(v["posrot"][7, 1]))
    # a view that asks for some arrays downloads only those; the rest follow when a getter needs them
    import ctypes as C
    L.frxStepWorldN(sc.world, DT, 5)
    pp = C.POINTER(C.c_float)()
    n = L.frxViewBodyStates(sc.world, C.byref(pp), None, None, None)
    pos_only = np.ctypeslib.as_array(pp, shape=(n, 4)).copy()
    box = L.frGetBodyAABB(sc.bodies[11])                      # stale array: must be fetched now
    st = fb.body_states(L, sc.world)
    assert np.array_equal(pos_only[:, :2], st["pos"])
    assert (box.x, box.y, box.width, box.height) == tuple(float(x) for x in st["aabb"][11])
    v = L.frGetBodyVelocity(sc.bodies[11])
    assert (v.x, v.y) == tuple(float(x) for x in st["vel"][11])
    fb.check_error(L)
    sc.release()


def test_small_and_large_world_solver_builds_agree_bit_for_bit(gpu):
    """k_solve_colored exists in two builds (fx_solve.cu): one block of up to 128 registers per SM for small
    worlds, two blocks of 64 registers for large ones, with different grid sizes.  Same coloured order,
    so the state after 200 steps must be identical; FEROX_SOLVE_WIDE pins the build, hence one
    subprocess each."""
    import hashlib, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import hashlib, numpy as np, ferox_b200 as fb\n"
            "from ferox_b200 import scenes\n"
            "L = fb.lib(); L.frxSetDevice(0)\n"
            "sc = scenes.Scene(L, scenes.mixed_pile(6000, width=150.0), prime=False)\n"
            "L.frStepWorld(sc.world, scenes.DT); L.frxStepWorldN(sc.world, scenes.DT, 200)\n"
            "st = fb.body_states(L, sc.world); fb.check_error(L)\n"
            "print('HASH', hashlib.sha1(b''.join(st[k].tobytes() for k in ('pos', 'angle', 'vel', 'omega'))).hexdigest(),"
            " fb.step_counters(L, sc.world)['manifolds'])\n")
    out = []
    for wide in ("0", "1"):
        env = dict(os.environ, FEROX_SOLVE_WIDE=wide, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out.append([l for l in r.stdout.splitlines() if l.startswith("HASH")][-1])
    assert out[0] == out[1], out


def test_mapped_forces_equal_apply_force_calls(gpu):
    """frxMapBodyForces hands out the pinned accumulators themselves: adding (fx, fy, torque) there must
    move the world exactly like frxApplyForces / frApplyForceToBody, and the view reads zero again
    after the step."""
    import ctypes as C
    L = gpu
    out = []
    for mapped in (False, True):
        spec = scenes.mixed_pile(1500, width=60.0)
        sc = scenes.Scene(L, spec, prime=False)
        L.frStepWorld(sc.world, DT)
        n = L.frGetBodyCountInWorld(sc.world)
        f = np.zeros((n, 2), np.float32); f[:, 0] = 0.3; f[::3, 1] = -0.2
        tq = np.linspace(-0.1, 0.1, n).astype(np.float32)
        for _ in range(30):
            if mapped:
                p = C.POINTER(C.c_float)()
                assert L.frxMapBodyForces(sc.world, C.byref(p)) == n
                fv = np.ctypeslib.as_array(p, shape=(n, 4))
                assert not fv.any()
                movable = fb.body_state_views(L, sc.world)["vel"][:, 3] > 0
                fv[movable, 0] += f[movable, 0]; fv[movable, 1] += f[movable, 1]; fv[movable, 2] += tq[movable]
            else:
                L.frxApplyForces(sc.world, f.ctypes.data_as(C.c_void_p), tq.ctypes.data_as(C.c_void_p), n)
            L.frStepWorld(sc.world, DT)
        fb.check_error(L)
        out.append(fb.body_states(L, sc.world))
        sc.release()
    for k in ("pos", "angle", "vel", "omega"):
        assert np.array_equal(out[0][k].view(np.uint32), out[1][k].view(np.uint32)), k


def test_update_world_accumulator_matches_reference_with_a_fake_clock(gpu, ref):
    """frUpdateWorld (src/world.c:268-285): the first call latches the clock, later calls run
    floor(accumulator / dt) steps and stamp contacts with the world clock (which arms the eviction rule,
    :222-235).  A preloaded deterministic frGetCurrentTime drives the reference and the product through
    the same 40 calls (0, 1 or 2 steps each); reference-order mode must end bit-identical."""
    import hashlib, os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(tempfile.mkdtemp(), "libfake_clock.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(root, "tests", "c", "fake_clock.c")])
    code = ("import sys, hashlib, numpy as np\n"
            "from ferox_b200 import capi, scenes\n"
            "import ferox_b200 as fb\n"
            "kind = sys.argv[1]\n"
            "L = fb.lib() if kind == 'gpu' else capi.load_abi(sys.argv[2])\n"
            "if kind == 'gpu': L.frxSetDevice(0)\n"
            "sc = scenes.Scene(L, scenes.mixed_pile(300, width=30.0), prime=False)\n"
            "if kind == 'gpu': L.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)\n"
            "for k in range(40): L.frUpdateWorld(sc.world, scenes.DT)\n"
            "st = scenes.snapshot(L, sc.bodies)\n"
            "print('HASH', hashlib.sha1(b''.join(st[k].tobytes() for k in ('pos', 'angle', 'vel', 'omega'))).hexdigest(),"
            " L.frGetBodyCountInWorld(sc.world), float(st['pos'][:, 1].mean()))\n")
    from tests.conftest import ref_path
    out = []
    for kind in ("ref", "gpu"):
        env = dict(os.environ, LD_PRELOAD=so, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code, kind, ref_path(8192)], capture_output=True, text=True, timeout=300, cwd=root, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out.append([l for l in r.stdout.splitlines() if l.startswith("HASH")][-1])
    assert out[0] == out[1], out
    assert int(out[0].split()[2]) == 303          # the bodies did enter the world (a call ran at least one step)


def test_update_world_replays_the_captured_graph_across_calls(gpu):
    """frUpdateWorld moves the world clock on every call (src/world.c:274,281).  The per-step scalars (clock, dt,
    gravity) live in a device block, not in the captured kernel arguments, so the CUDA graph of the launch
    sequence survives: under the same deterministic clock as above nearly every step of 120 calls is a replay."""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(tempfile.mkdtemp(), "libfake_clock.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(root, "tests", "c", "fake_clock.c")])
    code = ("import ferox_b200 as fb\n"
            "from ferox_b200 import scenes\n"
            "L = fb.lib(); L.frxSetDevice(0)\n"
            "sc = scenes.Scene(L, scenes.mixed_pile(6000, width=150.0), prime=False)\n"
            "for k in range(120): L.frUpdateWorld(sc.world, scenes.DT)\n"
            "c = fb.step_counters(L, sc.world); fb.check_error(L)\n"
            "print('GRAPH', c['steps'], L.frxGetGraphReplays(sc.world), c['manifolds'])\n")
    env = dict(os.environ, LD_PRELOAD=so, PYTHONPATH=root)
    env.pop("FEROX_GRAPH", None)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    steps, replays, manifolds = (int(x) for x in [l for l in r.stdout.splitlines() if l.startswith("GRAPH")][-1].split()[1:])
    assert steps >= 100 and manifolds > 0
    # the first steps size the buffers and wait for two identical parameter sets; growth re-captures a few times
    assert replays >= steps - 40, (steps, replays)


def test_mutating_handlers_forces_and_membership_match_reference(gpu, ref):
    """The SAME collision handlers on the unmodified reference and on the product (reference-order mode), doing
    what real programs do (examples/src/cows.c:375-393, melon.c:283-311): pre-step edits that must persist or
    take effect (count = 0 disables a manifold for the step, a friction edit sticks through the cache,
    src/world.c:209-214,366-367), forces applied before the step / inside the pre-step sweep (they act,
    :216-220) / inside the post-step sweep (they are wiped, :481-482), frGetBodyForce read in both sweeps, and
    bodies removed and added from inside a callback (:450-479).  The callback logs (order, pairs, counts, depths,
    impulses, forces seen) and every body state must agree bit for bit, step by step."""
    L = gpu
    base = scenes.mixed_pile(150, width=26.0)
    spec = scenes.SceneSpec(name="handlers", cell=base.cell, shapes=list(base.shapes))
    far = np.array([[5.0, -300.0], [9.0, -300.0]], np.float32)        # two bodies that never touch anything
    scenes._finish(spec, np.concatenate([base.body_type, [capi.BODY_DYNAMIC] * 2]), np.concatenate([base.body_shape, [2, 2]]),
                   np.concatenate([base.pos, far]), np.concatenate([base.angle, [0.0, 0.0]]))
    runs = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        lib.frStepWorld(sc.world, DT)
        spare = lib.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(13.0, -250.0), sc.shapes[2])
        sc.bodies.append(spare); sc.index_of[spare] = len(sc.bodies) - 1
        log, state = [], {"step": 0}

        def pre(key, value, lib=lib, sc=sc, log=log, state=state):
            c = value.contents
            a, b = sc.index_of[key.first], sc.index_of[key.second]
            f = lib.frGetBodyForce(key.second)
            log.append(("pre", state["step"], a, b, c.count, c.friction, c.contacts[0].depth, c.contacts[0].cache.normalScalar, f.x, f.y))
            if (a + b) % 7 == 0:
                c.count = 0
            if (a + b) % 5 == 0:
                c.friction = 0.125
            if (a + b) % 11 == 0:
                lib.frApplyForceToBody(key.second, lib.frGetBodyPosition(key.second), capi.Vector2(3.0, -2.0))

        def post(key, value, lib=lib, sc=sc, log=log, state=state):
            c = value.contents
            a, b = sc.index_of[key.first], sc.index_of[key.second]
            f = lib.frGetBodyForce(key.second)
            log.append(("post", state["step"], a, b, c.count, c.friction, c.contacts[0].cache.normalScalar,
                        c.contacts[0].cache.tangentScalar, c.contacts[0].cache.normalMass, f.x, f.y, lib.frGetBodyTorque(key.second)))
            lib.frApplyForceToBody(key.first, lib.frGetBodyPosition(key.first), capi.Vector2(50.0, 0.0))     # wiped by the post-step
            if state["step"] == 20 and "removed" not in state:
                state["removed"] = lib.frRemoveBodyFromWorld(sc.world, sc.bodies[-2])
            if state["step"] == 30 and "added" not in state:
                state["added"] = lib.frAddBodyToWorld(sc.world, sc.bodies[-1])

        h = capi.CollisionHandler(capi.CollisionEventFunc(pre), capi.CollisionEventFunc(post))
        lib.frSetWorldCollisionHandler(sc.world, h)
        snaps = []
        for s in range(60):
            state["step"] = s
            if s % 9 == 4:
                lib.frApplyForceToBody(sc.bodies[40], lib.frGetBodyPosition(sc.bodies[40]), capi.Vector2(-6.0, 1.5))
            lib.frStepWorld(sc.world, DT)
            members = [sc.index_of[lib.frGetBodyInWorld(sc.world, i)] for i in range(lib.frGetBodyCountInWorld(sc.world))]
            snaps.append((members, scenes.snapshot(lib, sc.bodies)))
        runs.append((log, snaps, h))
    fb.check_error(L)
    (rlog, rsnaps, _), (glog, gsnaps, _) = runs
    assert len(rlog) == len(glog) and len(rlog) > 2000
    f32 = lambda t: tuple(np.float32(x).tobytes() if isinstance(x, float) else x for x in t)
    for k, (x, y) in enumerate(zip(rlog, glog)):
        assert f32(x) == f32(y), (k, x, y)
    for s, ((rm, rst), (gm, gst)) in enumerate(zip(rsnaps, gsnaps)):
        assert rm == gm, "membership at step %d" % s
        assert_states_equal(rst, gst, "step %d" % s)
    assert any(e[0] == "pre" and e[4] > 0 and abs(e[5] - 0.125) < 1e-7 for e in rlog[len(rlog) // 2:])   # the friction edit stuck


def test_world_raycast_batch_matches_reference_for_thousands_of_rays(gpu, ref):
    """frxRaycastBatch (device sweep over the world's state, fx_bodies.cu k_raycast_batch) against the unmodified
    reference's frComputeWorldRaycast, ray by ray: the same bodies in the same (ascending-index) order with the
    same points, normals, distances and inside flags, bit for bit -- 3,000 rays through a settled mixed pile,
    long and short, from outside and from inside bodies (the polygon branch only sees crossings within one unit of
    the origin, src/collision.c:592-600, so many rays start next to a body)."""
    import ctypes as C
    L = gpu
    spec = scenes.mixed_pile(400, width=40.0)
    worlds = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        lib.frStepWorld(sc.world, DT)
        sc.step(120)
        worlds.append(sc)
    rs, gs = worlds
    assert_states_equal(rs.snapshot(), gs.snapshot(), "settled pile")
    st = gs.snapshot()
    rng = np.random.default_rng(5)
    n_rays = 3000
    rays = (capi.Ray * n_rays)()
    for k in range(n_rays):
        if k % 3 == 0:                               # from the box around the pile, any direction, long
            o = (rng.uniform(-5, 45), rng.uniform(-30, 5))
            dist = rng.uniform(5, 90)
        else:                                        # next to (or inside) a body, short
            c = st["pos"][rng.integers(3, spec.n)]
            o = (c[0] + rng.normal() * 0.6, c[1] + rng.normal() * 0.6)
            dist = rng.choice([0.3, 1.0, 2.5, 20.0])
        a = rng.uniform(0, 2 * np.pi)
        scale = rng.choice([1.0, 0.25, 7.0])         # un-normalised directions are normalised by the callee
        rays[k] = capi.Ray(capi.Vector2(float(np.float32(o[0])), float(np.float32(o[1]))),
                           capi.Vector2(float(np.float32(np.cos(a) * scale)), float(np.float32(np.sin(a) * scale))), float(np.float32(dist)))
    max_hits = 64
    hits = (capi.RaycastHit * (n_rays * max_hits))()
    counts = np.zeros(n_rays, np.int32)
    total = L.frxRaycastBatch(gs.world, C.byref(rays), n_rays, C.byref(hits), counts, max_hits)
    fb.check_error(L)
    assert total == int(counts.sum()) and counts.max() <= max_hits
    f32 = lambda x: np.float32(x).tobytes()
    n_hits = 0
    for k in range(n_rays):
        want = []
        cb = capi.RaycastQueryFunc(lambda h, ctx: want.append((rs.index_of[h.body], f32(h.point.x), f32(h.point.y), f32(h.normal.x), f32(h.normal.y),
                                                               f32(h.distance), bool(h.inside))))
        ref.frComputeWorldRaycast(rs.world, rays[k], cb, None)
        got = []
        for j in range(counts[k]):
            h = hits[k * max_hits + j]
            got.append((gs.index_of[h.body], f32(h.point.x), f32(h.point.y), f32(h.normal.x), f32(h.normal.y), f32(h.distance), bool(h.inside)))
        assert got == want, (k, got, want)
        n_hits += len(want)
    assert n_hits > 500, n_hits
    # frComputeWorldRaycast of the product is the same path, one ray at a time
    for k in range(0, n_rays, 97):
        got = []
        cb = capi.RaycastQueryFunc(lambda h, ctx: got.append(gs.index_of[h.body]))
        L.frComputeWorldRaycast(gs.world, rays[k], cb, None)
        assert len(got) == counts[k]
    rs.release(); gs.release()


def test_point_query_batch_matches_reference_contains_point(gpu, ref):
    """frxPointQueryBatch (device sweep) against the unmodified reference's frBodyContainsPoint evaluated body by body:
    for 1,500 points in and around a settled mixed pile the same ascending list of containing bodies."""
    import ctypes as C
    L = gpu
    spec = scenes.mixed_pile(400, width=40.0)
    worlds = []
    for lib in (ref, L):
        sc = scenes.Scene(lib, spec, prime=False)
        if lib is L:
            lib.frxSetWorldSolverMode(sc.world, fb.SOLVER_SERIAL)
        lib.frStepWorld(sc.world, DT)
        sc.step(120)
        worlds.append(sc)
    rs, gs = worlds
    st = gs.snapshot()
    assert_states_equal(rs.snapshot(), st, "settled pile")
    rng = np.random.default_rng(9)
    n = 1500
    pts = np.empty((n, 2), np.float32)
    for k in range(n):
        c = st["pos"][rng.integers(0, spec.n)]
        pts[k] = (c[0] + rng.normal() * 0.5, c[1] + rng.normal() * 0.5)
    max_per = 8
    idx = np.full(n * max_per, -1, np.int32)
    counts = np.zeros(n, np.int32)
    total = L.frxPointQueryBatch(gs.world, pts.ctypes.data_as(C.c_void_p), n, idx, counts, max_per)
    fb.check_error(L)
    members = [ref.frGetBodyInWorld(rs.world, i) for i in range(ref.frGetBodyCountInWorld(rs.world))]
    found = 0
    for k in range(n):
        p = capi.Vector2(float(pts[k, 0]), float(pts[k, 1]))
        want = [i for i, b in enumerate(members) if ref.frBodyContainsPoint(b, p)]
        got = idx[k * max_per:k * max_per + counts[k]].tolist()
        assert got == want, (k, got, want)
        found += len(want)
    assert total == found and found > 400, (total, found)
    # the host entry point is the same arithmetic
    for k in range(0, n, 50):
        p = capi.Vector2(float(pts[k, 0]), float(pts[k, 1]))
        for i in idx[k * max_per:k * max_per + counts[k]]:
            assert L.frBodyContainsPoint(L.frGetBodyInWorld(gs.world, int(i)), p)
    rs.release(); gs.release()
