// fx_raycast.cuh -- one ray against one body, shared by the host entry point (frComputeRaycast) and the device
// world-query kernel (fx_bodies.cu k_raycast_batch), so that both produce the same bits.
//
// Restates reference src/collision.c:138-220 (frComputeRaycast) and :561-625 (circle/line and line/line
// intersection), quirks included: the polygon branch accepts an edge crossing only when the ray parameter t is in
// [0, 1] -- with the direction normalised that is ONE UNIT of ray length, whatever maxDistance says (:592-600) -- and
// never writes `distance`; the circle branch reports `inside` for a negative distance.
#pragma once
#include "fx_math.cuh"

// frRaycastHit (include/ferox.h:186-199) with the body as a 64-bit word: a world index on the device, the frBody
// pointer once the host has patched it
struct RayHitPod {
    long long body;
    V2 point, normal;
    float distance;
    unsigned char inside, pad[3];
};
static_assert(sizeof(RayHitPod) == 32, "frRaycastHit layout");

FX_HD bool ray_circle_line(V2 center, float radius, V2 origin, V2 dir, float *distance) {
    V2 oc = vsub(center, origin);
    float dot = vdot(oc, dir);
    float h2 = vlen2(oc) - (dot * dot);
    float b2 = (radius * radius) - h2;
    *distance = dot - sqrtf(b2);
    return (dot >= 0.0f && b2 >= 0.0f);
}

FX_HD bool ray_lines(V2 o1, V2 d1, V2 o2, V2 d2, float *distance) {
    float rxs = vcross(d1, d2);
    V2 qp = vsub(o2, o1);
    float qpxs = vcross(qp, d2), qpxr = vcross(qp, d1);
    if (rxs != 0.0f) {
        float inv = 1.0f / rxs;
        float t = qpxs * inv, u = qpxr * inv;
        if ((t >= 0.0f && t <= 1.0f) && (u >= 0.0f && u <= 1.0f)) { *distance = t; return true; }
        return false;
    }
    if (qpxr != 0.0f) return false;
    float rdr = vdot(d1, d1), sdr = vdot(d2, d1);
    float inv = 1.0f / rdr;
    float qpdr = vdot(qp, d1);
    float k, t0 = qpdr * inv, t1 = t0 + sdr * inv;
    if (sdr < 0.0f) { k = t0; t0 = t1; t1 = k; }
    if ((t0 < 0.0f && t1 == 0.0f) || (t0 == 1.0f && t1 > 1.0f)) { *distance = (t0 == 1.0f); return true; }
    return (t1 >= 0.0f && t0 <= 1.0f);
}

// `dir` is already normalised (src/collision.c:141).  `hit` arrives as the caller initialised it and only the
// fields the reference writes are touched.  type: 1 circle, 2 polygon (FX_SHAPE_*).
FX_HD bool ray_vs_shape(int type, float radius, int nv, const V2 *verts, Xf t, V2 origin, V2 dir, float max_distance, RayHitPod &hit) {
    float distance = FLT_MAX;
    if (type == 1) {
        (void) ray_circle_line(t.p, radius, origin, dir, &distance);
        const bool result = (distance >= 0.0f) && (distance <= max_distance);
        hit.point = vadd(origin, vscale(dir, distance));
        hit.normal = vperp_left(vsub(origin, hit.point));
        hit.distance = distance;
        hit.inside = (distance < 0.0f) ? 1 : 0;
        return result;
    }
    if (type == 2) {
        int crossings = 0;
        float best = FLT_MAX;
        for (int j = nv - 1, i = 0; i < nv; j = i, ++i) {
            V2 a = xf_apply(verts[i], t), c = xf_apply(verts[j], t);
            V2 edge = vsub(a, c);
            const bool ok = ray_lines(origin, dir, c, edge, &distance);
            if (ok && distance <= max_distance) {
                if (best > distance) {
                    best = distance;
                    hit.point = vadd(origin, vscale(dir, best));
                    hit.normal = vperp_left(edge);
                }
                ++crossings;
            }
        }
        const bool inside = (crossings & 1) != 0;
        hit.inside = inside ? 1 : 0;
        return (!inside && crossings > 0);
    }
    return false;
}

// frBodyContainsPoint (src/rigid_body.c:261-289): a circle by squared distance; a polygon by the parity of the edge
// crossings of a ray cast along +x from the point (through frComputeRaycast, quirks included)
FX_HD bool shape_contains_point(int type, float radius, int nv, const V2 *verts, Xf t, V2 point) {
    if (type == 1) {
        const float dx = point.x - t.p.x, dy = point.y - t.p.y;
        return (dx * dx) + (dy * dy) <= radius * radius;
    }
    if (type == 2) {
        RayHitPod hit;
        hit.body = 0; hit.point = v2(0.f, 0.f); hit.normal = v2(0.f, 0.f); hit.distance = 0.0f; hit.inside = 0;
        (void) ray_vs_shape(type, radius, nv, verts, t, point, vunit(v2(1.0f, 0.0f)), FLT_MAX, hit);
        return hit.inside != 0;
    }
    return false;
}
