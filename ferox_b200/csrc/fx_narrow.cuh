// fx_narrow.cuh -- narrow phase device code (kernel K5): circle/circle, circle/polygon and
// polygon/polygon SAT + clipping.  Replaces reference src/collision.c:114-135, 228-558, 628-724.
//
// Execution shape: circle/circle pairs are one thread each.  Every pair that involves a polygon is
// evaluated cooperatively by a group of 8 lanes, ONE LANE PER POLYGON AXIS (FR_GEOMETRY_MAX_VERTEX_COUNT
// = 8), four pairs per warp: each lane walks the <= 8 vertices of the other polygon for its axis (the
// support search, sequential so "first maximum wins" holds by construction, src/collision.c:718) and
// the axis of least penetration is a 3-step shuffle arg-max across the group that prefers the lower
// index among equal depths (the reference's strict `<`, src/collision.c:697), so feature ids stay
// bit-exact.  The short scalar tail (edge selection, clipping) is computed redundantly by the 8
// lanes.  [measured, profiles/r01: a whole warp per pair (32 lanes, 2 vertices per lane) issued
// ~4x more warp instructions per pair and ran the 65k-body narrow phase in 248 us.]
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
FX_D ArgMax am_shfl_xor(unsigned int mask, ArgMax a, int d) {
    ArgMax b;
    b.v = __shfl_xor_sync(mask, a.v, d);
    b.i = __shfl_xor_sync(mask, a.i, d);
    return am_best(a, b);
}
// arg-max over the 8 lanes of a group (mask names exactly those lanes)
FX_D ArgMax am_group8(unsigned int mask, ArgMax a) {
    a = am_shfl_xor(mask, a, 1);
    a = am_shfl_xor(mask, a, 2);
    a = am_shfl_xor(mask, a, 4);
    return a;
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

// circle vs polygon in either order, src/collision.c:305-449.  Group-cooperative: the 8 lanes named
// by `gmask` call this with identical arguments and receive identical results; k = lane within group.
FX_D void grp_collide_circle_poly(unsigned int gmask, int k, const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    const bool first_is_circle = (s1->type == FX_SHAPE_CIRCLE);
    const ShapeDev *circle = first_is_circle ? s1 : s2, *poly = first_is_circle ? s2 : s1;
    const Xf ct = first_is_circle ? t1 : t2, pt = first_is_circle ? t2 : t1;
    const int nv = poly->nv;

    V2 c = xf_unrotate(vsub(ct.p, pt.p), pt);          // circle centre in polygon space (:328-330)
    const float radius = circle->radius;
    float dot = -INFINITY;
    if (k < nv) dot = vdot(poly->normals[k], vsub(c, poly->verts[k]));
    if (__ballot_sync(gmask, k < nv && dot > radius) & gmask) return;        // separating face (:345)
    ArgMax best = am_group8(gmask, am_make(dot, k, k < nv));
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

// farthest vertex of s along v (first maximum wins), src/collision.c:708-724: sequential, one lane
FX_D int support_index(const ShapeDev *s, Xf t, V2 v) {
    float max_dot = -FLT_MAX;
    int max_i = -1;
    const V2 vl = xf_unrotate(v, t);
    const int nv = s->nv;
    for (int i = 0; i < nv; ++i) {
        float dot = vdot(s->verts[i], vl);
        if (max_dot < dot) { max_dot = dot; max_i = i; }
    }
    return max_i;
}

// axis of least penetration of sa's faces against sb (src/collision.c:669-705): lane k owns axis k.
// Returns the face index (or -1) and its signed depth.
FX_D int grp_separating_axis(unsigned int gmask, int k, const ShapeDev *sa, Xf ta, const ShapeDev *sb, Xf tb, float *depth) {
    const int na = sa->nv;
    float d = -INFINITY;
    bool ok = false;
    if (k < na) {
        V2 vert = xf_apply(sa->verts[k], ta);
        V2 nrm = xf_rotate(sa->normals[k], ta);
        int si = support_index(sb, tb, vneg(nrm));
        if (si >= 0) {
            V2 sp = xf_apply(sb->verts[si], tb);
            d = vdot(nrm, vsub(sp, vert));
            ok = true;
        }
    }
    ArgMax best = am_group8(gmask, am_make(d, k, ok));
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

// polygon vs polygon, src/collision.c:456-558 (group-cooperative, uniform arguments and result)
FX_D void grp_collide_polys(unsigned int gmask, int k, const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 1;
    float dep1, dep2;
    int i1 = grp_separating_axis(gmask, k, s1, t1, s2, t2, &dep1);
    if (dep1 >= 0.0f) return;
    int i2 = grp_separating_axis(gmask, k, s2, t2, s1, t1, &dep2);
    if (dep2 >= 0.0f) return;

    V2 dir = (dep1 > dep2) ? xf_rotate(s1->normals[i1], t1) : xf_rotate(s2->normals[i2], t2);
    V2 delta = vsub(t2.p, t1.p);
    if (vdot(delta, dir) < 0.0f) dir = vneg(dir);

    // the two support searches of frGetContactEdge: lane k scores vertex k, shuffle arg-max
    V2 vl1 = xf_unrotate(dir, t1), vl2 = xf_unrotate(vneg(dir), t2);
    ArgMax a1 = am_group8(gmask, am_make(k < s1->nv ? vdot(s1->verts[k], vl1) : 0.0f, k, k < s1->nv));
    ArgMax a2 = am_group8(gmask, am_make(k < s2->nv ? vdot(s2->verts[k], vl2) : 0.0f, k, k < s2->nv));
    const int si1 = a1.i, si2 = a2.i;
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

// Dispatch for one pair evaluated by a group of 8 lanes (uniform arguments), src/collision.c:114-135.
FX_D void grp_collide(unsigned int gmask, int k, const ShapeDev *s1, Xf t1, const ShapeDev *s2, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    const int ty1 = s1->type, ty2 = s2->type;
    if (ty1 == FX_SHAPE_CIRCLE && ty2 == FX_SHAPE_CIRCLE) collide_circles(s1->radius, t1, s2->radius, t2, m);
    else if ((ty1 == FX_SHAPE_CIRCLE && ty2 == FX_SHAPE_POLYGON) || (ty1 == FX_SHAPE_POLYGON && ty2 == FX_SHAPE_CIRCLE))
        grp_collide_circle_poly(gmask, k, s1, t1, s2, t2, m);
    else if (ty1 == FX_SHAPE_POLYGON && ty2 == FX_SHAPE_POLYGON) grp_collide_polys(gmask, k, s1, t1, s2, t2, m);
}

// ---------------------------------------------------------------------------------------------
// One lane per pair.  The polygon data of the lane's two shapes sits in shared memory as
// [shape][component][vertex][lane] (conflict-free: consecutive lanes hit consecutive banks); all 32
// lanes run the reference's loops verbatim (src/collision.c:305-724), so ties, early-outs and
// feature ids are the reference's by construction.  The long dependent chains of a pair (sqrt/div in
// the edge normalisations, clipping) overlap across the 32 pairs of a warp instead of being paid
// once per pair as in the group-cooperative variant above.  [measured on B200, 65,536-body config 2:
// warp per pair 248 us, 8 lanes per pair 186 us, lane per pair: see profiles/]
// ---------------------------------------------------------------------------------------------
struct LaneShapes {
    float vx[2][FX_MAX_VERTS][32], vy[2][FX_MAX_VERTS][32];   // vertices
    float nx[2][FX_MAX_VERTS][32], ny[2][FX_MAX_VERTS][32];   // normals
};

struct LanePoly {          // what one lane knows about one of its polygons
    const LaneShapes *sh;
    int which, lane, nv;
    FX_D V2 vert(int i) const { return v2(sh->vx[which][i][lane], sh->vy[which][i][lane]); }
    FX_D V2 normal(int i) const { return v2(sh->nx[which][i][lane], sh->ny[which][i][lane]); }
};

// polygon read straight from the shape table (read-only for the whole kernel: L1-cacheable loads)
struct GlobalPoly {
    const ShapeDev *s;
    int nv;
    FX_D V2 vert(int i) const {
        float2 v = __ldg(reinterpret_cast<const float2 *>(&s->verts[i]));
        return v2(v.x, v.y);
    }
    FX_D V2 normal(int i) const {
        float2 v = __ldg(reinterpret_cast<const float2 *>(&s->normals[i]));
        return v2(v.x, v.y);
    }
};

FX_D void lane_stage_poly(LaneShapes *sh, int which, int lane, const ShapeDev *src, int nv) {
    const float4 *g = reinterpret_cast<const float4 *>(src);
    const int q = (nv + 1) >> 1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < q) {
            float4 v = g[2 + k], n = g[6 + k];
            sh->vx[which][2 * k][lane] = v.x; sh->vy[which][2 * k][lane] = v.y;
            sh->vx[which][2 * k + 1][lane] = v.z; sh->vy[which][2 * k + 1][lane] = v.w;
            sh->nx[which][2 * k][lane] = n.x; sh->ny[which][2 * k][lane] = n.y;
            sh->nx[which][2 * k + 1][lane] = n.z; sh->ny[which][2 * k + 1][lane] = n.w;
        }
}

template <class P>
FX_D int lane_support_index(const P &s, Xf t, V2 v) {      // src/collision.c:708-724
    float max_dot = -FLT_MAX;
    int max_i = -1;
    const V2 vl = xf_unrotate(v, t);
    for (int i = 0; i < s.nv; ++i) {
        float dot = vdot(s.vert(i), vl);
        if (max_dot < dot) { max_dot = dot; max_i = i; }
    }
    return max_i;
}

template <class P>
FX_D int lane_separating_axis(const P &sa, Xf ta, const P &sb, Xf tb, float *depth) {   // :669-705
    float max_depth = -FLT_MAX;
    int max_i = -1;
    for (int i = 0; i < sa.nv; ++i) {
        V2 vert = xf_apply(sa.vert(i), ta);
        V2 nrm = xf_rotate(sa.normal(i), ta);
        int si = lane_support_index(sb, tb, vneg(nrm));
        if (si < 0) return si;
        V2 sp = xf_apply(sb.vert(si), tb);
        float d = vdot(nrm, vsub(sp, vert));
        if (max_depth < d) { max_depth = d; max_i = i; }
    }
    *depth = max_depth;
    return max_i;
}

template <class P>
FX_D EdgeDev lane_contact_edge(const P &s, Xf t, V2 v) {    // :628-666
    const int si = lane_support_index(s, t, v), nv = s.nv;
    int prev = (si == 0) ? nv - 1 : si - 1;
    int next = (si == nv - 1) ? 0 : si + 1;
    V2 pe = vunit(vsub(s.vert(si), s.vert(prev)));
    V2 ne = vunit(vsub(s.vert(si), s.vert(next)));
    V2 vl = xf_unrotate(v, t);
    V2 sv = xf_apply(s.vert(si), t);
    EdgeDev e;
    if (vdot(pe, vl) < vdot(ne, vl)) {
        e.d0 = xf_apply(s.vert(prev), t); e.d1 = sv; e.d2 = sv; e.i0 = prev; e.i1 = si;
    } else {
        e.d0 = sv; e.d1 = xf_apply(s.vert(next), t); e.d2 = sv; e.i0 = si; e.i1 = next;
    }
    return e;
}

template <class P>
FX_D void lane_collide_polys(const P &s1, Xf t1, const P &s2, Xf t2, Manifold &m) {   // :456-558
    m.count = 0;
    m.has_ids = 1;
    if (s1.nv <= 0 || s2.nv <= 0) return;
    float dep1 = FLT_MAX, dep2 = FLT_MAX;
    int i1 = lane_separating_axis(s1, t1, s2, t2, &dep1);
    if (dep1 >= 0.0f) return;
    int i2 = lane_separating_axis(s2, t2, s1, t1, &dep2);
    if (dep2 >= 0.0f) return;

    V2 dir = (dep1 > dep2) ? xf_rotate(s1.normal(i1), t1) : xf_rotate(s2.normal(i2), t2);
    V2 delta = vsub(t2.p, t1.p);
    if (vdot(delta, dir) < 0.0f) dir = vneg(dir);

    EdgeDev e1 = lane_contact_edge(s1, t1, dir);
    EdgeDev e2 = lane_contact_edge(s2, t2, vneg(dir));
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
        m.id[0] = m.id[1] = id0;
        m.p[0] = m.p[1] = inc.d0;
        m.depth[0] = m.depth[1] = depth1;
        m.count = 1;
    } else {
        m.id[0] = id0; m.p[0] = inc.d0; m.depth[0] = depth1;
        m.id[1] = id1; m.p[1] = inc.d1; m.depth[1] = depth2;
        m.count = 2;
    }
}

// circle (radius) vs polygon; `circle_first` tells which body of the pair is the circle  (:305-449)
template <class P>
FX_D void lane_collide_circle_poly(bool circle_first, float radius, const P &poly, Xf t1, Xf t2, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    const Xf ct = circle_first ? t1 : t2, pt = circle_first ? t2 : t1;
    const int nv = poly.nv;
    V2 c = xf_unrotate(vsub(ct.p, pt.p), pt);
    float max_dot = -FLT_MAX;
    int mi = -1;
    for (int i = 0; i < nv; ++i) {                      // :340-348
        float dot = vdot(poly.normal(i), vsub(c, poly.vert(i)));
        if (dot > radius) return;
        if (max_dot < dot) { max_dot = dot; mi = i; }
    }
    if (mi < 0) return;
    V2 delta = vsub(t2.p, t1.p);
    if (max_dot < 0.0f) {
        m.dir = vneg(xf_rotate(poly.normal(mi), pt));
        if (vdot(delta, m.dir) < 0.0f) m.dir = vneg(m.dir);
        m.p[0] = vadd(ct.p, vscale(m.dir, radius));
        m.depth[0] = radius - max_dot;
    } else {
        V2 p1 = (mi > 0) ? poly.vert(mi - 1) : poly.vert(nv - 1);
        V2 p2 = poly.vert(mi);
        V2 edge = vsub(p2, p1);
        V2 p1c = vsub(c, p1), p2c = vsub(c, p2);
        float d1 = vdot(p1c, edge), d2 = vdot(p2c, vneg(edge));
        if (d1 < 0.0f || d2 < 0.0f) {
            V2 dv = (d1 < 0.0f) ? p1c : p2c;
            float m2 = vlen2(dv);
            if (radius * radius < m2) return;
            float mag = sqrtf(m2);
            if (mag <= 0.0f) mag = FLT_EPSILON;
            m.dir = vscale(xf_rotate(vneg(dv), pt), 1.0f / mag);
            if (vdot(delta, m.dir) < 0.0f) m.dir = vneg(m.dir);
            m.p[0] = xf_apply(vscale(m.dir, radius), ct);
            m.depth[0] = radius - mag;
        } else {
            m.dir = vneg(xf_rotate(poly.normal(mi), pt));
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
