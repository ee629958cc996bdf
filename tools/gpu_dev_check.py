"""Developer diagnostic (not a pytest): run the product on a GPU next to the oracle and print where
the first difference shows up.  Usage: python tests/gpu_dev_check.py [quick]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferox_b200 as fb  # noqa: E402
from ferox_b200 import capi, scenes  # noqa: E402
from oracle import orc  # noqa: E402

L = fb.lib()


def bits_equal(a, b):
    return a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()


def compare_states(tag, a, b):
    ok = True
    for k in a:
        if not bits_equal(a[k], b[k]):
            d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            print("  %s: %s differs: max abs %.3e at %s (n bad %d)" % (tag, k, d.max(), np.unravel_index(d.argmax(), d.shape), int((a[k].view(np.uint32) != b[k].view(np.uint32)).sum())))
            ok = False
    return ok


def check_scene(spec, steps, mode, capacity=None):
    t0 = time.time()
    w_cap = capacity or max(2048, spec.n + 16)
    os.environ["FEROX_MAX_OBJECT_COUNT"] = str(w_cap)
    sc = scenes.Scene(L, spec, prime=False)
    L.frxSetWorldSolverMode(sc.world, mode)
    L.frStepWorld(sc.world, scenes.DT)  # priming step: bodies enter the world
    fb.check_error(L)
    print("[%s] mode=%s n=%d in_world=%d" % (spec.name, "serial" if mode else "colored", spec.n, L.frGetBodyCountInWorld(sc.world)))
    ow = orc.OracleWorld.from_scene(sc)
    ok_all = True
    for s in range(steps):
        L.frStepWorld(sc.world, scenes.DT)
        ow.step(1)
        fb.check_error(L)
        pc = fb.candidate_pairs(L, sc.world)
        oc = ow.candidates()
        if mode == fb.SOLVER_SERIAL:
            same = pc.shape == oc.shape and (pc == oc).all()
        else:
            same = pc.shape == oc.shape and (np.array(sorted(map(tuple, pc))) == np.array(sorted(map(tuple, oc)))).all() if len(pc) else pc.shape == oc.shape
        if not same:
            print("  step %d: candidate %s differs: product %d oracle %d" % (s, "sequence" if mode else "set", len(pc), len(oc)))
            sp, so = set(map(tuple, pc)), set(map(tuple, oc))
            print("    only product:", sorted(sp - so)[:5], "only oracle:", sorted(so - sp)[:5])
            ok_all = False
            break
        if mode == fb.SOLVER_SERIAL:
            st = fb.body_states(L, sc.world)
            if not compare_states("step %d" % s, st, ow.state()):
                pm, om = fb.manifolds(L, sc.world), ow.cache()
                print("    manifolds product %d oracle %d" % (len(pm), len(om)))
                k = min(len(pm), len(om))
                for f in ("a", "b", "count"):
                    if not (pm[f][:k] == om[f][:k]).all():
                        i = int(np.argmax(pm[f][:k] != om[f][:k]))
                        print("    first manifold field %s mismatch at %d: %s vs %s" % (f, i, pm[i], om[i]))
                        break
                ok_all = False
                break
    st = fb.body_states(L, sc.world)
    os_ = ow.state()
    dv = np.abs(st["vel"] - os_["vel"]).max() if spec.n else 0.0
    dp = np.abs(st["pos"] - os_["pos"]).max() if spec.n else 0.0
    ctr = fb.step_counters(L, sc.world)
    print("  %s after %d steps: max|dv|=%.3e max|dx|=%.3e counters=%s oracle=%s (%.1fs)" % (
        "OK" if ok_all else "MISMATCH", steps, dv, dp, ctr, ow.counters(), time.time() - t0))
    sc.release()
    ow.close()
    return ok_all


def check_narrowphase(spec, n_pairs=20000, seed=3):
    sc = scenes.Scene(L, spec, prime=False)
    L.frStepWorld(sc.world, scenes.DT)
    ow = orc.OracleWorld.from_scene(sc)
    rng = np.random.default_rng(seed)
    n = spec.n
    # neighbours in the lattice overlap often: pair each body with a few nearby indices
    a = rng.integers(0, n, n_pairs)
    b = np.clip(a + rng.integers(-40, 41, n_pairs), 0, n - 1)
    keep = a != b
    pairs = np.stack([a[keep], b[keep]], 1).astype(np.int32)
    out = (capi.Collision * len(pairs))()
    hit = np.zeros(len(pairs), np.int32)
    nh = L.frxComputeCollisions(sc.world, np.ascontiguousarray(pairs.reshape(-1)), len(pairs), C.byref(out), hit.ctypes.data_as(C.c_void_p))
    fb.check_error(L)
    arr = np.frombuffer(out, dtype=orc.COLLISION_DTYPE)
    bad = 0
    for i, (x, y) in enumerate(pairs):
        h, col = ow.collide(int(x), int(y))
        if h != bool(hit[i]):
            bad += 1
            if bad < 5:
                print("  hit mismatch pair", x, y, h, hit[i])
            continue
        if h:
            p = arr[i]
            same = (p["count"] == col["count"] and p["dir"].tobytes() == col["dir"].tobytes())
            for k in range(2):
                for f in ("id", "depth", "point"):
                    same = same and p["contacts"][k][f].tobytes() == col["contacts"][k][f].tobytes()
            if not same:
                bad += 1
                if bad < 5:
                    print("  manifold mismatch pair", x, y, "\n   product", p, "\n   oracle ", col)
    print("[narrowphase %s] %d pairs, %d hits, %d mismatches" % (spec.name, len(pairs), nh, bad))
    sc.release()
    ow.close()
    return bad == 0


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ok = True
    a = np.linspace(np.pi, 3 * np.pi, 200001).astype(np.float32)
    s, c = np.empty_like(a), np.empty_like(a)
    L.frxDeviceSinCos(a, len(a), s, c)
    fb.check_error(L)
    hs = np.array([np.float32(np.sin(np.float64(x))) for x in a[:10]])
    print("device sincos sample:", s[:3], c[:3], "numpy f64->f32:", hs[:3])
    ok &= check_narrowphase(scenes.straddle_origin(1500))
    ok &= check_narrowphase(scenes.mixed_pile(3000, width=80.0))
    ok &= check_scene(scenes.box_stack(16, 4), 60, fb.SOLVER_SERIAL)
    ok &= check_scene(scenes.straddle_origin(400), 40, fb.SOLVER_SERIAL)
    ok &= check_scene(scenes.mixed_pile(600, width=40.0), 60, fb.SOLVER_SERIAL)
    ok &= check_scene(scenes.small_world(1), 60, fb.SOLVER_SERIAL)
    ok &= check_scene(scenes.box_stack(16, 4), 60, fb.SOLVER_COLORED)
    ok &= check_scene(scenes.mixed_pile(600, width=40.0), 60, fb.SOLVER_COLORED)
    if not quick:
        ok &= check_scene(scenes.mixed_pile(20000, width=256.0), 20, fb.SOLVER_COLORED)
    print("ALL OK" if ok else "SOME CHECKS FAILED")
