// fx_narrow.cuh -- narrow phase device code (kernel K5): circle/circle, circle/polygon and
// polygon/polygon SAT + clipping.  Replaces reference src/collision.c:114-135, 228-558, 628-724.
//
// Execution shape: circle/circle pairs are one thread each.  Every pair that involves a polygon is
// evaluated by a whole warp: the <= 8 x 8 (axis, vertex) dot products of a SAT direction sit one
// (axis, vertex-pair) per lane and the two arg-max searches (support vertex per axis, then the axis
// of least penetration) are shuffle reductions.  The reductions keep the reference's tie rule
// ("first maximum wins": strict `<` in index order, src/collision.c:697,718) by preferring the lower
// index among equal values, so feature ids stay bit-exact.  The short scalar tail (edge selection,
// clipping) is computed redundantly by all lanes, i.e. without divergence.
#pragma once
#include "fx_device.h"
#include "fx_scan.cuh"

struct ArgMax {
    float v;
    int i;
};

FX_D ArgMax am_make(float v, int i, bool valid) {
    ArgMax a;
    bool ok = valid && (v == v);   // a NaN never wins the reference's `max < x` test either
    a.v = ok ? v : -INFINITY;
    a.i = ok ? i : 0x7fff;
    return a;
}
FX_D ArgMax am_best(ArgMax a, ArgMax b) {
    bool take_b = (b.v > a.v) || (b.v == a.v && b.i < a.i);
    return take_b ? b : a;
}
FX_D ArgMax am_shfl_xor(ArgMax a, int d) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, d);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, d);
    return am_best(a, b);
}

// circle vs circle, src/collision.c:264-298 (one thread)
FX_D void collide_circles(float r1, Xf t1, float r2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    V2 d = vsub(t2.p, t1.p);
    float rsum = r1 + r2;
    float m2 = vlen2(d);
    if (rsum * rsum < m2) return;
    float mag = sqrtf(m2);
    if (mag <= 0.0f) { d.x = 0.0f; d.y = mag = FLT_EPSILON; }
    m.dir = vscale(d, 1.0f / mag);
    m.p[0] = xf_apply(vscale(m.dir, r1), t1);
    m.depth[0] = rsum - mag;
    m.p[1] = m.p[0];
    m.depth[1] = m.depth[0];
    m.id[0] = m.id[1] = 0;
    m.count = 1;
}

// circle vs polygon in either order, src/collision.c:305-449.  Warp-cooperative: all 32 lanes call
// this with identical arguments and receive identical results.
FX_D void warp_collide_circle_poly(const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    const bool first_is_circle = (s1->type == FX_SHAPE_CIRCLE);
    const ShapeDev *circle = first_is_circle ? s1 : s2, *poly = first_is_circle ? s2 : s1;
    const Xf ct = first_is_circle ? t1 : t2, pt = first_is_circle ? t2 : t1;
    const int nv = poly->nv;
    const unsigned int lane = lane_id();

    V2 c = xf_unrotate(vsub(ct.p, pt.p), pt);          // circle centre in polygon space (:328-330)
    const float radius = circle->radius;
    float dot = -INFINITY;
    if ((int) lane < nv) dot = vdot(poly->normals[lane], vsub(c, poly->verts[lane]));
    if (__any_sync(0xffffffffu, (int) lane < nv && dot > radius)) return;   // separating face (:345)
    ArgMax best = am_make(dot, (int) lane, (int) lane < nv);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) best = am_shfl_xor(best, d);
    if (best.i >= nv) return;                                             // no vertices (:350)
    const int mi = best.i;
    const float max_dot = best.v;

    V2 delta = vsub(t2.p, t1.p);
    if (max_dot < 0.0f) {                               // centre inside the polygon (:358-375)
        m.dir = vneg(xf_rotate(poly->normals[mi], pt));
        if (vdot(delta, m.dir) < 0.0f) m.dir = vneg(m.dir);
        m.p[0] = vadd(ct.p, vscale(m.dir, radius));
        m.depth[0] = radius - max_dot;
    } else {
        V2 p1 = (mi > 0) ? poly->verts[mi - 1] : poly->verts[nv - 1];
        V2 p2 = poly->verts[mi];
        V2 edge = vsub(p2, p1);
        V2 p1c = vsub(c, p1), p2c = vsub(c, p2);
        float d1 = vdot(p1c, edge), d2 = vdot(p2c, vneg(edge));
        if (d1 < 0.0f || d2 < 0.0f) {                   // vertex region (:395-424)
            V2 dv = (d1 < 0.0f) ? p1c : p2c;
            float m2 = vlen2(dv);
            if (radius * radius < m2) return;
            float mag = sqrtf(m2);
            if (mag <= 0.0f) mag = FLT_EPSILON;
            m.dir = vscale(xf_rotate(vneg(dv), pt), 1.0f / mag);
            if (vdot(delta, m.dir) < 0.0f) m.dir = vneg(m.dir);
            m.p[0] = xf_apply(vscale(m.dir, radius), ct);
            m.depth[0] = radius - mag;
        } else {                                        // face region (:425-445)
            m.dir = vneg(xf_rotate(poly->normals[mi], pt));
            if (vdot(delta, m.dir) < 0.0f) m.dir = vneg(m.dir);
            m.p[0] = vadd(ct.p, vscale(m.dir, radius));
            m.depth[0] = radius - max_dot;
        }
    }
    m.p[1] = m.p[0];
    m.depth[1] = m.depth[0];
    m.id[0] = m.id[1] = 0;
    m.count = 1;
}

// axis of least penetration of sa's faces against sb (src/collision.c:669-705): lane = axis*4 + q,
// q covers vertices q and q+4 of sb.  Returns the face index (or -1) and its signed depth.
FX_D int warp_separating_axis(const ShapeDev *sa, Xf ta, const ShapeDev *sb, Xf tb, float *depth) {
    const unsigned int lane = lane_id();
    const int ax = (int) (lane >> 2), q = (int) (lane & 3u);
    const int na = sa->nv, nb = sb->nv;
    const bool ax_ok = ax < na;
    V2 nrm = v2(0.0f, 0.0f), vert = v2(0.0f, 0.0f);
    if (ax_ok) {
        vert = xf_apply(sa->verts[ax], ta);
        nrm = xf_rotate(sa->normals[ax], ta);
    }
    V2 v = xf_unrotate(vneg(nrm), tb);                 // support direction in sb's frame (:715)
    ArgMax sup = am_make(0.0f, 0, false);
    if (q < nb) sup = am_best(sup, am_make(vdot(sb->verts[q], v), q, true));
    if (q + 4 < nb) sup = am_best(sup, am_make(vdot(sb->verts[q + 4], v), q + 4, true));
    sup = am_shfl_xor(sup, 1);
    sup = am_shfl_xor(sup, 2);
    float d = -INFINITY;
    if (ax_ok && sup.i < nb) {
        V2 sp = xf_apply(sb->verts[sup.i], tb);
        d = vdot(nrm, vsub(sp, vert));
    }
    ArgMax best = am_make(d, ax, ax_ok && sup.i < nb);
    best = am_shfl_xor(best, 4);
    best = am_shfl_xor(best, 8);
    best = am_shfl_xor(best, 16);
    if (best.i >= na) { *depth = FLT_MAX; return -1; }
    *depth = best.v;
    return best.i;
}

struct EdgeDev {
    V2 d0, d1, d2;
    int i0, i1;
};

// src/collision.c:628-666 once the support index is known (scalar, computed by every lane)
FX_D EdgeDev contact_edge_from_support(const ShapeDev *s, Xf t, V2 v, int si) {
    const int nv = s->nv;
    int prev = (si == 0) ? nv - 1 : si - 1;
    int next = (si == nv - 1) ? 0 : si + 1;
    V2 pe = vunit(vsub(s->verts[si], s->verts[prev]));
    V2 ne = vunit(vsub(s->verts[si], s->verts[next]));
    V2 vl = xf_unrotate(v, t);
    V2 sv = xf_apply(s->verts[si], t);
    EdgeDev e;
    if (vdot(pe, vl) < vdot(ne, vl)) {
        e.d0 = xf_apply(s->verts[prev], t); e.d1 = sv; e.d2 = sv; e.i0 = prev; e.i1 = si;
    } else {
        e.d0 = sv; e.d1 = xf_apply(s->verts[next], t); e.d2 = sv; e.i0 = si; e.i1 = next;
    }
    return e;
}

// src/collision.c:228-257; the indices are NOT swapped when the points are (:249-252)
FX_D bool clip_edge(EdgeDev &e, V2 v, float dot) {
    float a = vdot(e.d0, v) - dot;
    float b = vdot(e.d1, v) - dot;
    if (a >= 0.0f && b >= 0.0f) return true;
    V2 ev = vsub(e.d1, e.d0);
    V2 mid = vadd(e.d0, vscale(ev, (a / (a - b))));
    if (a > 0.0f && b < 0.0f) { e.d1 = mid; return true; }
    if (a < 0.0f && b > 0.0f) { e.d0 = e.d1; e.d1 = mid; return true; }
    return false;
}

// polygon vs polygon, src/collision.c:456-558 (warp-cooperative, uniform arguments and result)
FX_D void warp_collide_polys(const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 1;
    float dep1, dep2;
    int i1 = warp_separating_axis(s1, t1, s2, t2, &dep1);
    if (dep1 >= 0.0f) return;
    int i2 = warp_separating_axis(s2, t2, s1, t1, &dep2);
    if (dep2 >= 0.0f) return;

    V2 dir = (dep1 > dep2) ? xf_rotate(s1->normals[i1], t1) : xf_rotate(s2->normals[i2], t2);
    V2 delta = vsub(t2.p, t1.p);
    if (vdot(delta, dir) < 0.0f) dir = vneg(dir);

    // both support searches at once: lanes 0-7 scan s1 along dir, lanes 8-15 scan s2 along -dir
    const unsigned int lane = lane_id();
    const int grp = (int) (lane >> 3), k = (int) (lane & 7u);
    const ShapeDev *ss = (grp == 0) ? s1 : s2;
    const Xf ts = (grp == 0) ? t1 : t2;
    V2 vl = xf_unrotate((grp == 0) ? dir : vneg(dir), ts);
    ArgMax sup = am_make((grp < 2 && k < ss->nv) ? vdot(ss->verts[k], vl) : 0.0f, k, grp < 2 && k < ss->nv);
    sup = am_shfl_xor(sup, 1);
    sup = am_shfl_xor(sup, 2);
    sup = am_shfl_xor(sup, 4);
    int si1 = __shfl_sync(0xffffffffu, sup.i, 0), si2 = __shfl_sync(0xffffffffu, sup.i, 8);
    if (si1 >= s1->nv || si2 >= s2->nv) return;

    EdgeDev e1 = contact_edge_from_support(s1, t1, dir, si1);
    EdgeDev e2 = contact_edge_from_support(s2, t2, vneg(dir), si2);
    EdgeDev ref = e1, inc = e2;
    float ed1 = vdot(vsub(e1.d1, e1.d0), dir);
    float ed2 = vdot(vsub(e2.d1, e2.d0), dir);
    int flipped = 0;
    if (fabsf(ed1) > fabsf(ed2)) { ref = e2; inc = e1; flipped = 1; }      // :497-502

    V2 rv = vunit(vsub(ref.d1, ref.d0));
    float rd1 = vdot(ref.d0, rv), rd2 = vdot(ref.d1, rv);
    if (!clip_edge(inc, rv, rd1)) return;
    if (!clip_edge(inc, vneg(rv), -rd2)) return;

    V2 rn = vperp_right(rv);
    float max_depth = vdot(ref.d2, rn);
    float depth1 = vdot(inc.d0, rn) - max_depth;
    float depth2 = vdot(inc.d1, rn) - max_depth;

    m.dir = dir;
    unsigned int mask = ((unsigned int) flipped << 16) | ((unsigned int) ref.i0 << 8);   // :524-528
    int id0 = (int) (mask | (unsigned int) inc.i0), id1 = (int) (mask | (unsigned int) inc.i1);
    if (depth1 < 0.0f) {
        m.id[0] = m.id[1] = id1;
        m.p[0] = m.p[1] = inc.d1;
        m.depth[0] = m.depth[1] = depth2;
        m.count = 1;
    } else if (depth2 < 0.0f) {
        m.id[0] = id0;
        m.p[0] = inc.d0;
        m.depth[0] = depth1;
        m.id[1] = m.id[0]; m.p[1] = m.p[0]; m.depth[1] = m.depth[0];
        m.count = 1;
    } else {
        m.id[0] = id0; m.p[0] = inc.d0; m.depth[0] = depth1;
        m.id[1] = id1; m.p[1] = inc.d1; m.depth[1] = depth2;
        m.count = 2;
    }
}

// Dispatch for one pair evaluated by a whole warp (uniform arguments), src/collision.c:114-135.
FX_D void warp_collide(const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    if (s1 == nullptr || s2 == nullptr) return;
    const int ty1 = s1->type, ty2 = s2->type;
    if (ty1 == FX_SHAPE_CIRCLE && ty2 == FX_SHAPE_CIRCLE) collide_circles(s1->radius, t1, s2->radius, t2, m);
    else if ((ty1 == FX_SHAPE_CIRCLE && ty2 == FX_SHAPE_POLYGON) || (ty1 == FX_SHAPE_POLYGON && ty2 == FX_SHAPE_CIRCLE))
        warp_collide_circle_poly(s1, t1, s2, t2, m);
    else if (ty1 == FX_SHAPE_POLYGON && ty2 == FX_SHAPE_POLYGON) warp_collide_polys(s1, t1, s2, t2, m);
}
