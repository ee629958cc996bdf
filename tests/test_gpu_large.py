"""GPU parity at the sizes and on the kernel variants the BASELINE configs actually run.

Worlds of >= 400k manifolds switch to another set of kernel instantiations (fx_world.cpp `fill_dev_params`:
contact-major solver records, k_emit_pairs<4>, k_narrowphase_tiled<4>, k_carry_over<8>, k_solve_colored<2>).
Two things are pinned here:
  * FEROX_FORCE_LARGE=1 selects that set at ANY size, and the golden / reference-sequence / reference-order
    tests are repeated under it (bit-exact, as for the small-world set);
  * full-size worlds against oracle/step_oracle.c: config 2 at 65,536 bodies (candidate sequence, manifold
    geometry, >= 20 reference-order steps, all bit for bit) and one step of a 262,144-body pile -- large
    enough to select the large-world set by itself -- (pair set, contact counts / ids / points / depths).
"""
import os

import numpy as np
import pytest

import ferox_b200 as fb
from ferox_b200 import capi, scenes
from oracle import orc
from tests.helpers import RefCallLog, assert_bits_equal, assert_caches_equal, assert_states_equal

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DT = scenes.DT


def make_world(L, spec, mode):
    sc = scenes.Scene(L, spec, prime=False)
    L.frxSetWorldSolverMode(sc.world, mode)
    L.frStepWorld(sc.world, DT)
    assert L.frGetBodyCountInWorld(sc.world) == spec.n
    fb.check_error(L)
    return sc


def settled_spec(L, spec, steps):
    """`spec` stepped `steps` times by the product (production mode); returns a NEW spec holding the settled positions,
    angles and velocities, so that fresh worlds (empty contact caches on every side) can start from a deep pile."""
    import copy
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    L.frxStepWorldN(sc.world, DT, steps)
    fb.check_error(L)
    st = fb.body_states(L, sc.world)
    sc.release()
    out = copy.copy(spec)
    out.pos, out.angle, out.vel, out.omega = st["pos"].copy(), st["angle"].copy(), st["vel"].copy(), st["omega"].copy()
    out.angle[spec.body_type == capi.BODY_STATIC] = 0.0      # statics were created unrotated; keep their AABBs as created
    return out


def pair_order(a, b):
    return np.lexsort((np.asarray(b), np.asarray(a)))


def assert_geometry_equal(m, oc, what):
    """manifolds of the product vs cache entries of the oracle, both sorted by pair: everything the narrow phase
    computes (count, normal, feature ids, points, depths) bit for bit"""
    m, oc = m[pair_order(m["a"], m["b"])], oc[pair_order(oc["a"], oc["b"])]
    assert len(m) == len(oc), (what, len(m), len(oc))
    for f in ("a", "b", "count", "dir"):
        assert m[f].tobytes() == oc[f].tobytes(), "%s: %s differs" % (what, f)
    one = m["count"] == 1
    for f in ("id", "depth", "point"):
        x, y = np.ascontiguousarray(m["contacts"][f]), np.ascontiguousarray(oc["contacts"][f])
        assert x.tobytes() == y.tobytes(), "%s: contact %s differs" % (what, f)
    assert one.sum() + (m["count"] == 2).sum() == len(m)


# ------------------------------------------------------------------------------------------------
# the large-world kernel set, forced at small sizes: same bits as the reference
# ------------------------------------------------------------------------------------------------
SERIAL_SCENES = [
    ("box_stack_32x4", lambda: scenes.box_stack(32, 4)),
    ("straddle_origin_300", lambda: scenes.straddle_origin(300)),
    ("mixed_pile_400", lambda: scenes.mixed_pile(400, width=40.0)),
    ("small_world_5", lambda: scenes.small_world(5)),
]


@pytest.mark.parametrize("name,make", SERIAL_SCENES, ids=[s[0] for s in SERIAL_SCENES])
def test_forced_large_variants_match_golden_reference_run(gpu, monkeypatch, name, make):
    """tests/test_gpu_parity.py::test_serial_mode_matches_golden_reference_run on the large-world kernel set:
    k_emit_pairs<4>, k_narrowphase_tiled<4> and k_carry_over<8> produce the candidate SEQUENCE, the contact cache
    (array order) and the trajectories of the unmodified reference, bit for bit."""
    monkeypatch.setenv("FEROX_FORCE_LARGE", "1")
    L = gpu
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = make_world(L, make(), fb.SOLVER_SERIAL)
    L.frxStepWorldN(sc.world, DT, int(g["steps"]))
    fb.check_error(L)
    st = fb.body_states(L, sc.world)
    for k in ("pos", "angle", "sincos", "vel", "omega", "aabb"):
        assert_bits_equal(st[k], g["final_" + k], "final " + k)
    assert (fb.candidate_pairs(L, sc.world) == g["last_candidates"]).all()
    m = fb.manifolds(L, sc.world)
    assert_caches_equal(g["last_cache"], m[m["count"] > 0], name)
    sc.release()


@pytest.mark.parametrize("cell", [2.0, 0.7])
def test_forced_large_variants_pair_sequence_matches_reference(gpu, ref, monkeypatch, cell):
    monkeypatch.setenv("FEROX_FORCE_LARGE", "1")
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
    sc.release(); rs.release()


def test_forced_large_variants_colored_mode_same_bits_as_small_variants(gpu, monkeypatch):
    """Production (graph-coloured) mode: contact-major records + k_solve_colored<2> + the wide tiles visit the same
    manifolds in the same coloured order as the small-world set, so 300 steps of a 6,000-body pile end in identical
    bits, manifold for manifold."""
    L = gpu
    out = []
    for large in ("0", "1"):
        monkeypatch.setenv("FEROX_FORCE_LARGE", large)
        sc = make_world(L, scenes.mixed_pile(6000, width=150.0), fb.SOLVER_COLORED)
        L.frxStepWorldN(sc.world, DT, 300)
        fb.check_error(L)
        out.append((fb.body_states(L, sc.world), fb.manifolds(L, sc.world), fb.step_counters(L, sc.world)))
        sc.release()
    assert_states_equal(out[0][0], out[1][0], "forced-large vs default")
    assert_caches_equal(out[0][1], out[1][1], "manifolds", timestamps=True)
    assert out[0][2]["colors"] == out[1][2]["colors"] and out[0][2]["manifolds"] > 6000


@pytest.mark.parametrize("blocks", ["1", "3"])
def test_lean_records_with_several_manifolds_per_thread_same_bits(gpu, monkeypatch, blocks):
    """Large grid-solved worlds sweep LEAN records (normal + lever arms; tangent and unit perpendiculars recomputed every
    sweep), and only 1 M-body worlds give a thread more than one manifold per colour phase.  Squeezing the grid to 1 / 3
    blocks (FEROX_SOLVE_BLOCKS, read per world) gives every thread of a 6,000-body pile up to four, and 300 steps must end
    in the bits of the default small-world build (full records, one manifold per thread), manifold for manifold."""
    L = gpu
    out = []
    for env in ({"FEROX_FORCE_LARGE": "0", "FEROX_SOLVE_BLOCKS": "0"}, {"FEROX_FORCE_LARGE": "1", "FEROX_SOLVE_BLOCKS": blocks}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sc = make_world(L, scenes.mixed_pile(6000, width=150.0), fb.SOLVER_COLORED)
        L.frxStepWorldN(sc.world, DT, 300)
        fb.check_error(L)
        out.append((fb.body_states(L, sc.world), fb.manifolds(L, sc.world), fb.step_counters(L, sc.world)))
        sc.release()
    assert_states_equal(out[0][0], out[1][0], "lean records, squeezed grid vs default")
    assert_caches_equal(out[0][1], out[1][1], "manifolds", timestamps=True)
    assert out[0][2]["colors"] == out[1][2]["colors"] and out[0][2]["manifolds"] > 6000


# ------------------------------------------------------------------------------------------------
# full-size worlds against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("large", ["0", "1"])
def test_config2_full_size_reference_order_vs_oracle(gpu, monkeypatch, large):
    """BASELINE config 2 at its full size (65,536 mixed bodies + 3 walls), started from the pile the production mode
    settles into: 20 steps in reference order against oracle/step_oracle.c -- the candidate SEQUENCE of every
    step, the contact cache (array order, every field) and the body state, bit for bit; once on each kernel set."""
    monkeypatch.setenv("FEROX_FORCE_LARGE", large)
    L = gpu
    spec = settled_spec(L, scenes.mixed_pile(65536), 400)
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    ow = orc.OracleWorld.from_scene(sc)
    for s in range(20):
        L.frStepWorld(sc.world, DT)
        ow.step(1)
        if s in (0, 1, 19):
            got, want = fb.candidate_pairs(L, sc.world), ow.candidates()
            assert got.shape == want.shape and (got == want).all(), "candidate sequence, step %d" % s
            m = fb.manifolds(L, sc.world)
            assert_caches_equal(ow.cache(), m, "cache, step %d" % s)
    fb.check_error(L)
    c = fb.step_counters(L, sc.world)
    oc = ow.counters()
    assert c["overflow"] == 0 and c["cellEntries"] == oc["K"] and c["candidates"] == oc["C"] and c["manifolds"] == oc["M"] \
        and c["contacts"] == oc["Q"], (c, oc)
    assert c["manifolds"] > 80000                       # a deep pile, not a falling lattice
    assert_states_equal(fb.body_states(L, sc.world), ow.state(), "config 2, 20 reference-order steps")
    sc.release(); ow.close()


def test_262k_bodies_one_step_pairs_and_contacts_vs_oracle(gpu):
    """A 262,144-body pile (>= 400k manifolds: the large-world kernel set is what runs, unforced).  One production
    step from the settled state against the oracle: the candidate pair SET with its hit flags, and for every
    manifold the count, normal, feature ids, contact points and depths, bit for bit; counters K / C equal."""
    L = gpu
    spec = scenes.mixed_pile(262144, shared_shapes=True)
    sc = make_world(L, spec, fb.SOLVER_COLORED)
    L.frxStepWorldN(sc.world, DT, 300)
    fb.check_error(L)
    assert fb.step_counters(L, sc.world)["manifolds"] >= 400000
    ow = orc.OracleWorld.from_scene(sc)
    L.frStepWorld(sc.world, DT)
    ow.step(1)
    got, want = fb.candidate_pairs(L, sc.world), ow.candidates()
    assert got.shape == want.shape
    got, want = got[pair_order(got[:, 0], got[:, 1])], want[pair_order(want[:, 0], want[:, 1])]
    assert (got == want).all(), "candidate set / hit flags"
    c, oc = fb.step_counters(L, sc.world), ow.counters()
    assert c["overflow"] == 0 and c["cellEntries"] == oc["K"] and c["candidates"] == oc["C"], (c, oc)
    # the oracle started with an empty cache: its entries are exactly this step's hits; ours also carries the
    # entries no candidate touched (App. A.7), which have no counterpart there
    m = fb.manifolds(L, sc.world)
    hits = want[want[:, 2] == 1]
    key = lambda a, b: a.astype(np.int64) * (1 << 32) + b
    m = m[np.isin(key(m["a"], m["b"]), key(hits[:, 0], hits[:, 1]))]
    assert_geometry_equal(m, ow.cache(), "262k bodies")
    sc.release(); ow.close()


def test_config3_column_recipe_overruns_the_reference_solver_too(gpu):
    """SURVEY.md section 8d's config 3 (column 2,048 rows high, lattice pitch 1.05) on a 16-column slice, next to the
    oracle: the reference's own arithmetic does not hold such a column up -- 10 Gauss-Seidel sweeps per step cannot
    carry the weight of 2,000 rows, the column telescopes into the floor in free fall (mean speed = g t) and
    later bursts -- and the GPU step reproduces that trajectory BIT FOR BIT in reference order.  bench.py therefore
    reports the full-size recipe as it is (`config3`) and a shallow 1 M-circle bed beside it (`config3_shallow`)."""
    L = gpu
    spec = scenes.column_collapse(16 * 2048, cols=16)
    sc = make_world(L, spec, fb.SOLVER_SERIAL)
    ow = orc.OracleWorld.from_scene(sc)
    L.frxStepWorldN(sc.world, DT, 300)
    ow.step(300)
    fb.check_error(L)
    got, want = fb.body_states(L, sc.world), ow.state()
    assert_states_equal(got, want, "column slice, 300 reference-order steps")
    speed = np.linalg.norm(want["vel"][2:], axis=1)
    assert speed.mean() > 0.8 * 9.8 * 300 * DT             # still falling freely through itself after 5 s
    assert fb.step_counters(L, sc.world)["manifolds"] < spec.n // 2
    sc.release(); ow.close()
