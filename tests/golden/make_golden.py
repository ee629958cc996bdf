"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libferox_ref_8192.so,
compiled from /root/reference/src by oracle/Makefile) in THIS container.  The vectors travel with the
repo; /root/reference does not.  Run:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ferox_b200 import capi, scenes  # noqa: E402

SCENES = {
    "box_stack_32x4": (lambda: scenes.box_stack(32, 4), 120),
    "straddle_origin_300": (lambda: scenes.straddle_origin(300), 80),
    "mixed_pile_400": (lambda: scenes.mixed_pile(400, width=40.0), 150),
    "small_world_5": (lambda: scenes.small_world(5), 150),
}


def records_to_array(recs):
    from oracle import orc
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


def main():
    ref = capi.load_abi(os.path.join(ROOT, "oracle", "_ref", "libferox_ref_8192.so"))
    ref.refshim_count.restype = C.c_size_t

    class Call(C.Structure):
        _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("hit", C.c_int)]
    ref.refshim_calls.restype = C.POINTER(Call)
    for name, (make, steps) in SCENES.items():
        spec = make()
        sc = scenes.Scene(ref, spec)
        init = sc.snapshot()
        rec = scenes.ManifoldRecorder(sc)
        ref.refshim_enable(1)
        for _ in range(steps):
            ref.refshim_reset()
            rec.take()
            sc.step(1)
        n = ref.refshim_count()
        p = ref.refshim_calls()
        cands = np.array([(sc.index_of[p[i].a], sc.index_of[p[i].b], p[i].hit) for i in range(n)], np.int32).reshape(-1, 3)
        ref.refshim_enable(0)
        final = sc.snapshot()
        cache = records_to_array(rec.take())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), steps=steps, last_candidates=cands, last_cache=cache,
                            **{"init_" + k: v for k, v in init.items()}, **{"final_" + k: v for k, v in final.items()})
        print(name, "bodies", spec.n, "steps", steps, "candidates", len(cands), "manifolds", len(cache))
        rec.remove()
        sc.release()

    # known-answer vectors of SURVEY.md section 8c, re-measured on the reference here
    mat = capi.Material(1.0, 0.5, 0.0)
    rect = ref.frCreateRectangle(mat, 2.0, 1.0)
    verts = ref.frGetPolygonVertices(rect).contents
    normals = ref.frGetPolygonNormals(rect).contents
    b1 = ref.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.0, 0.0), rect)
    b2 = ref.frCreateBodyFromShape(capi.BODY_DYNAMIC, capi.Vector2(0.5, 0.9), rect)
    col = capi.Collision()
    hit = ref.frComputeCollision(b1, b2, C.byref(col))
    b3 = ref.frCreateBody(capi.BODY_DYNAMIC, capi.Vector2(0.0, 0.0))
    ref.frSetBodyAngle(b3, 0.1)
    np.savez_compressed(
        os.path.join(HERE, "known_answers.npz"),
        rect_verts=np.array([(verts.data[i].x, verts.data[i].y) for i in range(verts.count)], np.float32),
        rect_normals=np.array([(normals.data[i].x, normals.data[i].y) for i in range(normals.count)], np.float32),
        rect_area=np.float32(ref.frGetShapeArea(rect)), rect_mass=np.float32(ref.frGetShapeMass(rect)),
        rect_inertia=np.float32(ref.frGetShapeInertia(rect)),
        two_box_hit=np.int32(hit), two_box=np.frombuffer(bytes(col), dtype=np.uint8).copy(),
        angle_0p1=np.float32(ref.frGetBodyAngle(b3)))
    print("known answers: area", ref.frGetShapeArea(rect), "inertia", ref.frGetShapeInertia(rect), "hit", hit,
          "count", col.count, "ids", hex(col.contacts[0].id), hex(col.contacts[1].id), "angle", ref.frGetBodyAngle(b3))


if __name__ == "__main__":
    main()
