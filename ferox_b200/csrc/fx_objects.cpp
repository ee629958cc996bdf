// fx_objects.cpp -- host-side set-up objects of the ferox.h API: shapes, bodies, the stand-alone
// spatial hash, ray casts and the clock.  None of this is on the world-step hot path (SURVEY.md
// section 2: "setup-time host C; tables uploaded to device"); the arithmetic follows the reference
// so that the tables the kernels consume (hull order, normals, area, mass, inertia, AABB) are
// bit-identical to what the reference would have built.  Compiled with -ffp-contract=off.
//
// Replaces reference src/geometry.c (shapes), src/rigid_body.c:67-330 + 556-591 (body object,
// accessors, force/impulse accumulators, mass), src/broad_phase.c (stand-alone hash),
// src/collision.c:138-220, 561-625 (ray casts) and src/timer.c.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>

#include "fx_host.h"
#include "fx_raycast.cuh"

static uint64_t g_shape_epoch = 1;
uint64_t fx_shape_epoch() { return g_shape_epoch; }
void fx_shape_touch() { ++g_shape_epoch; }

static inline V2 V(frVector2 a) { return v2(a.x, a.y); }
static inline frVector2 F(V2 a) { frVector2 r; r.x = a.x; r.y = a.y; return r; }
static inline Xf to_xf(frTransform t) { Xf x; x.p = V(t.position); x.sn = t.rotation.sin_; x.cs = t.rotation.cos_; return x; }

// ================================================================================================
// shapes (reference src/geometry.c)
// ================================================================================================

// frVector2CounterClockwise, reference include/ferox.h:698-715
static inline int ccw(frVector2 a, frVector2 b, frVector2 c) {
    float lhs = (b.y - a.y) * (c.x - a.x);
    float rhs = (c.y - a.y) * (b.x - a.x);
    return (lhs > rhs) - (lhs < rhs);
}
static inline float dist2(frVector2 a, frVector2 b) {   // include/ferox.h:638-640
    return (b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y);
}

// Gift wrapping with the reference's start vertex and tie rules (src/geometry.c:361-413): start at
// the first vertex with minimal x; wrap by picking, among all other points, the most clockwise one
// as seen from the current vertex (collinear: the farther one); stop when the start comes back.
static void wrap_hull(const frVertices *in, frVertices *out) {
    if (in == nullptr || out == nullptr || in->count < 3) return;
    const int n = in->count;
    int start = 0;
    for (int i = 1; i < n; ++i)
        if (in->data[start].x > in->data[i].x) start = i;
    out->count = 0;
    out->data[out->count++] = in->data[start];
    int cur = start, next = cur;
    for (;;) {
        for (int i = 0; i < n; ++i)
            if (i != cur) { next = i; break; }
        for (int i = 0; i < n; ++i) {
            if (i == cur || i == next) continue;
            int dir = ccw(in->data[cur], in->data[i], in->data[next]);
            if (dir < 0) continue;
            float to_i = dist2(in->data[cur], in->data[i]);
            float to_next = dist2(in->data[cur], in->data[next]);
            if (dir != 0 || to_i > to_next) next = i;
        }
        if (next == start) break;
        cur = next;
        if (out->count >= FR_GEOMETRY_MAX_VERTEX_COUNT) break;   // defensive: never overrun the POD
        out->data[out->count++] = in->data[next];
    }
}

extern "C" void frSetPolygonVertices(frShape *s, const frVertices *vertices) {
    if (s == nullptr || vertices == nullptr || vertices->count <= 0) return;
    frVertices hull;
    std::memset(&hull, 0, sizeof hull);
    frVertices in = *vertices;
    if (in.count > FR_GEOMETRY_MAX_VERTEX_COUNT) in.count = FR_GEOMETRY_MAX_VERTEX_COUNT;
    wrap_hull(&in, &hull);
    s->vertices.count = hull.count;
    s->normals.count = hull.count;
    for (int i = 0; i < hull.count; ++i) s->vertices.data[i] = hull.data[i];
    // normal i belongs to the edge (i-1 -> i): unit left perpendicular (src/geometry.c:329-334)
    for (int j = hull.count - 1, i = 0; i < hull.count; j = i, ++i)
        s->normals.data[i] = F(vperp_left(vsub(V(s->vertices.data[i]), V(s->vertices.data[j]))));
    float twice = 0.0f;                                      // fan triangulation, :337-352
    for (int i = 0; i < s->vertices.count - 1; ++i) {
        float t = vcross(vsub(V(s->vertices.data[i]), V(s->vertices.data[0])),
                         vsub(V(s->vertices.data[i + 1]), V(s->vertices.data[0])));
        twice += t;
    }
    s->area = fabsf(0.5f * twice);
    fx_shape_touch();
}

extern "C" void frSetCircleRadius(frShape *s, float radius) {
    if (s == nullptr || s->type != FR_SHAPE_CIRCLE) return;
    s->radius = radius;
    s->area = (float) (3.14159265358979323846 * (double) (radius * radius));   // double product, :291
    fx_shape_touch();
}

static void set_rectangle(frShape *s, float width, float height) {
    float hw = 0.5f * width, hh = 0.5f * height;
    frVertices v;
    std::memset(&v, 0, sizeof v);
    v.data[0].x = -hw; v.data[0].y = -hh;
    v.data[1].x = -hw; v.data[1].y = hh;
    v.data[2].x = hw;  v.data[2].y = hh;
    v.data[3].x = hw;  v.data[3].y = -hh;
    v.count = 4;
    frSetPolygonVertices(s, &v);
}

static frShape *new_shape(frShapeType type, frMaterial material) {
    frShape *s = (frShape *) std::calloc(1, sizeof(frShape));
    if (s == nullptr) return nullptr;
    s->type = type;
    s->material = material;
    return s;
}

extern "C" frShape *frCreateCircle(frMaterial material, float radius) {
    if (radius <= 0.0f) return nullptr;
    frShape *s = new_shape(FR_SHAPE_CIRCLE, material);
    frSetCircleRadius(s, radius);
    return s;
}
extern "C" frShape *frCreateRectangle(frMaterial material, float width, float height) {
    if (width <= 0.0f || height <= 0.0f) return nullptr;
    frShape *s = new_shape(FR_SHAPE_POLYGON, material);
    set_rectangle(s, width, height);
    return s;
}
extern "C" frShape *frCreatePolygon(frMaterial material, const frVertices *vertices) {
    if (vertices == nullptr || vertices->count <= 0) return nullptr;
    frShape *s = new_shape(FR_SHAPE_POLYGON, material);
    frSetPolygonVertices(s, vertices);
    return s;
}
extern "C" void frReleaseShape(frShape *s) { std::free(s); }
extern "C" void frSetRectangleDimensions(frShape *s, float width, float height) {
    if (s == nullptr || width <= 0.0f || height <= 0.0f) return;
    set_rectangle(s, width, height);
}

extern "C" frShapeType frGetShapeType(const frShape *s) { return s ? s->type : FR_SHAPE_UNKNOWN; }
extern "C" frMaterial frGetShapeMaterial(const frShape *s) {
    frMaterial z = { 0.f, 0.f, 0.f };
    return s ? s->material : z;
}
extern "C" float frGetShapeDensity(const frShape *s) { return s ? s->material.density : 0.0f; }
extern "C" float frGetShapeFriction(const frShape *s) { return s ? s->material.friction : 0.0f; }
extern "C" float frGetShapeRestitution(const frShape *s) { return s ? s->material.restitution : 0.0f; }
extern "C" float frGetShapeArea(const frShape *s) { return s ? s->area : 0.0f; }
extern "C" float frGetShapeMass(const frShape *s) { return s ? s->material.density * s->area : 0.0f; }

// src/geometry.c:152-180; note the polygon formula yields inertia / area
extern "C" float frGetShapeInertia(const frShape *s) {
    if (s == nullptr || s->material.density <= 0.0f) return 0.0f;
    if (s->type == FR_SHAPE_CIRCLE) return 0.5f * frGetShapeMass(s) * (s->radius * s->radius);
    if (s->type == FR_SHAPE_POLYGON) {
        float num = 0.0f, den = 0.0f;
        const int n = s->vertices.count;
        for (int j = n - 1, i = 0; i < n; j = i, ++i) {
            V2 a = V(s->vertices.data[j]), b = V(s->vertices.data[i]);
            float cr = vcross(a, b), ds = (vdot(a, a) + vdot(a, b) + vdot(b, b));
            num += (cr * ds);
            den += cr;
        }
        return s->material.density * (num / (6.0f * den));
    }
    return 0.0f;
}

frAABB fx_shape_aabb(const frShape *s, frTransform tx) {     // src/geometry.c:183-216
    frAABB r = { 0.f, 0.f, 0.f, 0.f };
    if (s == nullptr) return r;
    if (s->type == FR_SHAPE_CIRCLE) {
        r.x = tx.position.x - s->radius;
        r.y = tx.position.y - s->radius;
        r.width = r.height = 2.0f * s->radius;
    } else if (s->type == FR_SHAPE_POLYGON) {
        float lox = FLT_MAX, loy = FLT_MAX, hix = -FLT_MAX, hiy = -FLT_MAX;
        Xf t = to_xf(tx);
        for (int i = 0; i < s->vertices.count; ++i) {
            V2 v = xf_apply(V(s->vertices.data[i]), t);
            if (lox > v.x) lox = v.x;
            if (loy > v.y) loy = v.y;
            if (hix < v.x) hix = v.x;
            if (hiy < v.y) hiy = v.y;
        }
        r.x = lox; r.y = loy; r.width = hix - lox; r.height = hiy - loy;
    }
    return r;
}
extern "C" frAABB frGetShapeAABB(const frShape *s, frTransform tx) { return fx_shape_aabb(s, tx); }

extern "C" float frGetCircleRadius(const frShape *s) { return (frGetShapeType(s) == FR_SHAPE_CIRCLE) ? s->radius : 0.0f; }
extern "C" frVector2 frGetPolygonVertex(const frShape *s, int i) {
    frVector2 z = { 0.f, 0.f };
    if (frGetShapeType(s) != FR_SHAPE_POLYGON || i < 0 || i >= s->vertices.count) return z;
    return s->vertices.data[i];
}
extern "C" const frVertices *frGetPolygonVertices(const frShape *s) {
    return (frGetShapeType(s) == FR_SHAPE_POLYGON) ? &s->vertices : nullptr;
}
extern "C" frVector2 frGetPolygonNormal(const frShape *s, int i) {
    frVector2 z = { 0.f, 0.f };
    if (frGetShapeType(s) != FR_SHAPE_POLYGON || i < 0 || i >= s->normals.count) return z;
    return s->normals.data[i];
}
extern "C" const frVertices *frGetPolygonNormals(const frShape *s) {
    return (frGetShapeType(s) == FR_SHAPE_POLYGON) ? &s->normals : nullptr;
}
extern "C" void frSetShapeType(frShape *s, frShapeType type) { if (s) { s->type = type; fx_shape_touch(); } }
extern "C" void frSetShapeMaterial(frShape *s, frMaterial m) { if (s) { s->material = m; fx_shape_touch(); } }
extern "C" void frSetShapeDensity(frShape *s, float d) { if (s) { s->material.density = d; fx_shape_touch(); } }
extern "C" void frSetShapeFriction(frShape *s, float f) { if (s) { s->material.friction = f; fx_shape_touch(); } }
extern "C" void frSetShapeRestitution(frShape *s, float r) { if (s) { s->material.restitution = r; fx_shape_touch(); } }

// ================================================================================================
// bodies (reference src/rigid_body.c:67-330, 556-591)
// ================================================================================================
static inline float bits_of(frBodyType type, frBodyFlags flags) {
    uint32_t u = ((uint32_t) type & 0xffu) | (((uint32_t) flags & 0xffffu) << 8);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct RowRef {
    float4 *posrot, *vel, *mprop, *force, *aabb;
    float *angle;
};

static RowRef row_of(frBody *b) {
    RowRef r;
    if (b->world != nullptr) {
        frWorld *w = b->world;
        fx_world_pull(w);
        const int i = b->index;
        r.posrot = &w->h.posrot[i]; r.vel = &w->h.vel[i]; r.mprop = &w->h.mprop[i];
        r.force = &w->h.force[i]; r.aabb = &w->h.aabb[i]; r.angle = &w->h.angle[i];
    } else {
        r.posrot = &b->loc.posrot; r.vel = &b->loc.vel; r.mprop = &b->loc.mprop;
        r.force = &b->loc.force; r.aabb = &b->loc.aabb; r.angle = &b->loc.angle;
    }
    return r;
}
static inline void mark(frBody *b, unsigned bits) { if (b->world) fx_world_mark_body(b->world, b->index, bits); }

static frTransform tx_of(const RowRef &r) {
    frTransform t;
    t.position.x = r.posrot->x; t.position.y = r.posrot->y;
    t.rotation.sin_ = r.posrot->z; t.rotation.cos_ = r.posrot->w;
    t.angle = *r.angle;
    return t;
}
static void store_aabb(const RowRef &r, frAABB a) { *r.aabb = make_float4(a.x, a.y, a.width, a.height); }

static void push_mass_props(frBody *b) {
    RowRef r = row_of(b);
    r.vel->w = b->inv_mass;
    *r.mprop = make_float4(b->mass, b->inv_inertia, b->gravity_scale, bits_of(b->type, b->flags));
    mark(b, FX_DIRTY_VEL | FX_DIRTY_MPROP);
}

// frComputeBodyMass, src/rigid_body.c:556-586
void fx_body_recompute_mass(frBody *b) {
    b->mass = b->inv_mass = 0.0f;
    b->inertia = b->inv_inertia = 0.0f;
    RowRef r = row_of(b);
    switch (b->type) {
        case FR_BODY_STATIC:
            r.vel->x = r.vel->y = r.vel->z = 0.0f;
            break;
        case FR_BODY_DYNAMIC:
            if (!(b->flags & FR_FLAG_INFINITE_MASS)) {
                b->mass = frGetShapeMass(b->shape);
                if (b->mass > 0.0f) b->inv_mass = 1.0f / b->mass;
            }
            if (!(b->flags & FR_FLAG_INFINITE_INERTIA)) {
                b->inertia = frGetShapeInertia(b->shape);
                if (b->inertia > 0.0f) b->inv_inertia = 1.0f / b->inertia;
            }
            break;
        default:
            break;
    }
    push_mass_props(b);
}

extern "C" frBody *frCreateBody(frBodyType type, frVector2 position) {
    if (type < FR_BODY_STATIC || type > FR_BODY_DYNAMIC) return nullptr;
    frBody *b = (frBody *) std::calloc(1, sizeof(frBody));
    if (b == nullptr) return nullptr;
    b->type = type;
    b->gravity_scale = 1.0f;
    b->index = -1;
    b->loc.posrot = make_float4(position.x, position.y, 0.0f, 1.0f);
    b->loc.mprop = make_float4(0.f, 0.f, 1.0f, bits_of(type, 0));
    return b;
}
extern "C" frBody *frCreateBodyFromShape(frBodyType type, frVector2 position, frShape *s) {
    frBody *b = frCreateBody(type, position);
    frSetBodyShape(b, s);
    return b;
}
extern "C" void frReleaseBody(frBody *b) { std::free(b); }

extern "C" frBodyType frGetBodyType(const frBody *b) { return b ? b->type : FR_BODY_UNKNOWN; }
extern "C" frBodyFlags frGetBodyFlags(const frBody *b) { return b ? b->flags : 0u; }
extern "C" frShape *frGetBodyShape(const frBody *b) { return b ? b->shape : nullptr; }
extern "C" frTransform frGetBodyTransform(const frBody *b) {
    frTransform z;
    std::memset(&z, 0, sizeof z);
    return b ? tx_of(row_of(const_cast<frBody *>(b))) : z;
}
extern "C" frVector2 frGetBodyPosition(const frBody *b) { return frGetBodyTransform(b).position; }
extern "C" float frGetBodyAngle(const frBody *b) { return b ? *row_of(const_cast<frBody *>(b)).angle : 0.0f; }
extern "C" float frGetBodyMass(const frBody *b) { return b ? b->mass : 0.0f; }
extern "C" float frGetBodyInverseMass(const frBody *b) { return b ? b->inv_mass : 0.0f; }
extern "C" float frGetBodyInertia(const frBody *b) { return b ? b->inertia : 0.0f; }
extern "C" float frGetBodyInverseInertia(const frBody *b) { return b ? b->inv_inertia : 0.0f; }
extern "C" float frGetBodyGravityScale(const frBody *b) { return b ? b->gravity_scale : 0.0f; }
extern "C" frVector2 frGetBodyVelocity(const frBody *b) {
    frVector2 z = { 0.f, 0.f };
    if (!b) return z;
    RowRef r = row_of(const_cast<frBody *>(b));
    z.x = r.vel->x; z.y = r.vel->y;
    return z;
}
extern "C" float frGetBodyAngularVelocity(const frBody *b) { return b ? row_of(const_cast<frBody *>(b)).vel->z : 0.0f; }
extern "C" frVector2 frGetBodyForce(const frBody *b) {
    frVector2 z = { 0.f, 0.f };
    if (!b) return z;
    RowRef r = row_of(const_cast<frBody *>(b));
    z.x = r.force->x; z.y = r.force->y;
    return z;
}
extern "C" float frGetBodyTorque(const frBody *b) { return b ? row_of(const_cast<frBody *>(b)).force->z : 0.0f; }
extern "C" frAABB frGetBodyAABB(const frBody *b) {
    frAABB z = { 0.f, 0.f, 0.f, 0.f };
    if (!b || !b->shape) return z;
    float4 a = *row_of(const_cast<frBody *>(b)).aabb;
    z.x = a.x; z.y = a.y; z.width = a.z; z.height = a.w;
    return z;
}
extern "C" void *frGetBodyUserData(const frBody *b) { return b ? b->ctx : nullptr; }

extern "C" void frSetBodyType(frBody *b, frBodyType type) {
    if (!b) return;
    b->type = type;
    fx_body_recompute_mass(b);
}
extern "C" void frSetBodyFlags(frBody *b, frBodyFlags flags) {
    if (!b) return;
    b->flags = flags;
    fx_body_recompute_mass(b);
}
extern "C" void frSetBodyShape(frBody *b, frShape *s) {
    if (!b) return;
    b->shape = s;
    RowRef r = row_of(b);
    store_aabb(r, fx_shape_aabb(s, tx_of(r)));      // zero AABB when s is NULL, rigid_body.c:210
    mark(b, FX_DIRTY_AABB | FX_DIRTY_SHAPE);
    fx_body_recompute_mass(b);
}
// src/rigid_body.c:216-222: the AABB is shifted by the new ABSOLUTE position
extern "C" void frSetBodyPosition(frBody *b, frVector2 position) {
    if (!b) return;
    RowRef r = row_of(b);
    r.posrot->x = position.x;
    r.posrot->y = position.y;
    r.aabb->x += position.x;
    r.aabb->y += position.y;
    mark(b, FX_DIRTY_POSROT | FX_DIRTY_AABB);
}
// src/rigid_body.c:225-238
extern "C" void frSetBodyAngle(frBody *b, float angle) {
    if (!b) return;
    RowRef r = row_of(b);
    if (*r.angle == angle) return;
    *r.angle = fx_normalize_angle(angle);
    r.posrot->z = sinf(*r.angle);
    r.posrot->w = cosf(*r.angle);
    store_aabb(r, fx_shape_aabb(b->shape, tx_of(r)));
    mark(b, FX_DIRTY_POSROT | FX_DIRTY_ANGLE | FX_DIRTY_AABB);
}
// Declared at ferox.h:476 and defined nowhere in the reference.  Defined here as: adopt the
// transform verbatim (caller-supplied sin/cos included) and refresh the AABB.
extern "C" void frSetBodyTransform(frBody *b, frTransform tx) {
    if (!b) return;
    RowRef r = row_of(b);
    *r.posrot = make_float4(tx.position.x, tx.position.y, tx.rotation.sin_, tx.rotation.cos_);
    *r.angle = tx.angle;
    store_aabb(r, fx_shape_aabb(b->shape, tx));
    mark(b, FX_DIRTY_POSROT | FX_DIRTY_ANGLE | FX_DIRTY_AABB);
}
extern "C" void frSetBodyGravityScale(frBody *b, float scale) {
    if (!b) return;
    b->gravity_scale = scale;
    push_mass_props(b);
}
extern "C" void frSetBodyVelocity(frBody *b, frVector2 v) {
    if (!b) return;
    RowRef r = row_of(b);
    r.vel->x = v.x; r.vel->y = v.y;
    mark(b, FX_DIRTY_VEL);
}
extern "C" void frSetBodyAngularVelocity(frBody *b, float av) {
    if (!b) return;
    row_of(b).vel->z = av;
    mark(b, FX_DIRTY_VEL);
}
extern "C" void frSetBodyUserData(frBody *b, void *ctx) { if (b) b->ctx = ctx; }

extern "C" void frClearBodyForces(frBody *b) {
    if (!b) return;
    *row_of(b).force = make_float4(0.f, 0.f, 0.f, 0.f);
    mark(b, FX_DIRTY_FORCE);
}
// src/rigid_body.c:299-306
extern "C" void frApplyForceToBody(frBody *b, frVector2 point, frVector2 force) {
    if (!b || b->inv_mass <= 0.0f) return;
    RowRef r = row_of(b);
    V2 local = vsub(V(point), v2(r.posrot->x, r.posrot->y));
    r.force->x = r.force->x + force.x;
    r.force->y = r.force->y + force.y;
    r.force->z += vcross(local, V(force));
    mark(b, FX_DIRTY_FORCE);
}
// src/rigid_body.c:309-316
extern "C" void frApplyGravityToBody(frBody *b, frVector2 g) {
    if (!b || b->mass <= 0.0f) return;
    RowRef r = row_of(b);
    float k = b->gravity_scale * b->mass;
    r.force->x = r.force->x + g.x * k;
    r.force->y = r.force->y + g.y * k;
    mark(b, FX_DIRTY_FORCE);
}
// src/rigid_body.c:319-330
extern "C" void frApplyImpulseToBody(frBody *b, frVector2 point, frVector2 impulse) {
    if (!b || b->inv_mass <= 0.0f) return;
    RowRef r = row_of(b);
    V2 local = vsub(V(point), v2(r.posrot->x, r.posrot->y));
    r.vel->x = r.vel->x + impulse.x * b->inv_mass;
    r.vel->y = r.vel->y + impulse.y * b->inv_mass;
    r.vel->z += b->inv_inertia * vcross(local, V(impulse));
    mark(b, FX_DIRTY_VEL);
}

// ================================================================================================
// ray casts and point queries (reference src/collision.c:138-220, 561-625; rigid_body.c:261-289)
// ================================================================================================
// (the arithmetic lives in fx_raycast.cuh, shared with the device world-query kernel)
extern "C" bool frComputeRaycast(const frBody *b, frRay ray, frRaycastHit *hit) {
    if (b == nullptr) return false;
    const V2 dir = vunit(V(ray.direction)), origin = V(ray.origin);
    const frShape *s = b->shape;
    const frTransform tx = frGetBodyTransform(b);
    const frShapeType type = frGetShapeType(s);
    if (type != FR_SHAPE_CIRCLE && type != FR_SHAPE_POLYGON) return false;
    RayHitPod pod;
    std::memset(&pod, 0, sizeof pod);
    if (hit != nullptr) { pod.point = V(hit->point); pod.normal = V(hit->normal); pod.distance = hit->distance; pod.inside = hit->inside ? 1 : 0; }
    V2 verts[FX_MAX_VERTS] = {};
    int nv = 0;
    if (type == FR_SHAPE_POLYGON) {
        nv = s->vertices.count < FX_MAX_VERTS ? s->vertices.count : FX_MAX_VERTS;
        for (int k = 0; k < nv; ++k) verts[k] = V(s->vertices.data[k]);
    }
    const bool result = ray_vs_shape((int) type, s->radius, nv, verts, to_xf(tx), origin, dir, ray.maxDistance, pod);
    if (hit != nullptr) {
        // (pod started as *hit and only the fields the reference writes were touched)
        hit->body = const_cast<frBody *>(b);
        hit->point = F(pod.point);
        hit->normal = F(pod.normal);
        hit->distance = pod.distance;
        hit->inside = pod.inside != 0;
    }
    return result;
}

extern "C" bool frBodyContainsPoint(const frBody *b, frVector2 point) {
    if (b == nullptr) return false;
    const frShape *s = b->shape;
    const frShapeType type = frGetShapeType(s);
    if (type != FR_SHAPE_CIRCLE && type != FR_SHAPE_POLYGON) return false;
    V2 verts[FX_MAX_VERTS] = {};
    int nv = 0;
    if (type == FR_SHAPE_POLYGON) {
        nv = s->vertices.count < FX_MAX_VERTS ? s->vertices.count : FX_MAX_VERTS;
        for (int k = 0; k < nv; ++k) verts[k] = V(s->vertices.data[k]);
    }
    return shape_contains_point((int) type, s->radius, nv, verts, to_xf(frGetBodyTransform(b)), V(point));
}

// ================================================================================================
// stand-alone spatial hash (reference src/broad_phase.c) -- query API, host side
// ================================================================================================
struct CellKey {
    int x, y;
    bool operator==(const CellKey &o) const { return x == o.x && y == o.y; }
};
struct CellKeyHash {
    size_t operator()(const CellKey &k) const {
        uint64_t z = ((uint64_t) (uint32_t) k.x << 32) | (uint32_t) k.y;
        z = (z ^ (z >> 31)) * 0x9E3779B97F4A7C15ull;
        return (size_t) (z ^ (z >> 29));
    }
};
struct frSpatialHash_ {
    float cell, inv_cell;
    std::unordered_map<CellKey, std::vector<int>, CellKeyHash> cells;
    std::vector<int> result;
};

extern "C" frSpatialHash *frCreateSpatialHash(float cellSize) {
    if (cellSize <= 0.0f) return nullptr;
    frSpatialHash *sh = new (std::nothrow) frSpatialHash_();
    if (!sh) return nullptr;
    sh->cell = cellSize;
    sh->inv_cell = 1.0f / cellSize;
    return sh;
}
extern "C" void frReleaseSpatialHash(frSpatialHash *sh) { delete sh; }
extern "C" void frClearSpatialHash(frSpatialHash *sh) {
    if (!sh) return;
    for (auto &kv : sh->cells) kv.second.clear();      // buckets persist, src/broad_phase.c:86-91
}
extern "C" float frGetSpatialHashCellSize(const frSpatialHash *sh) { return sh ? sh->cell : 0.0f; }
extern "C" void frInsertIntoSpatialHash(frSpatialHash *sh, frAABB key, int value) {
    if (!sh) return;
    int x0 = cell_coord(key.x, sh->inv_cell), y0 = cell_coord(key.y, sh->inv_cell);
    int x1 = cell_coord(key.x + key.width, sh->inv_cell), y1 = cell_coord(key.y + key.height, sh->inv_cell);
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) sh->cells[CellKey{ x, y }].push_back(value);
}
extern "C" void frQuerySpatialHash(frSpatialHash *sh, frAABB aabb, frHashQueryFunc func, void *userData) {
    if (!sh) return;
    int x0 = cell_coord(aabb.x, sh->inv_cell), y0 = cell_coord(aabb.y, sh->inv_cell);
    int x1 = cell_coord(aabb.x + aabb.width, sh->inv_cell), y1 = cell_coord(aabb.y + aabb.height, sh->inv_cell);
    sh->result.clear();
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            auto it = sh->cells.find(CellKey{ x, y });
            if (it == sh->cells.end()) continue;
            sh->result.insert(sh->result.end(), it->second.begin(), it->second.end());
        }
    // unique ids, ascending (the reference's byte-array pass, src/broad_phase.c:159-170)
    std::vector<int> &r = sh->result;
    if (r.size() > 1) {
        std::qsort(r.data(), r.size(), sizeof(int), [](const void *a, const void *b) {
            int x = *(const int *) a, y = *(const int *) b;
            return (x > y) - (x < y);
        });
        size_t k = 1;
        for (size_t i = 1; i < r.size(); ++i)
            if (r[i] != r[i - 1]) r[k++] = r[i];
        r.resize(k);
    }
    if (func == nullptr) return;
    std::vector<int> ids = r;   // the callback may re-enter the hash
    for (int id : ids) {
        frContextNode node;
        node.id = id;
        node.ctx = userData;
        (void) func(node);      // the bool result is ignored, src/broad_phase.c:176-178
    }
}

// ================================================================================================
// clock (reference src/timer.c): monotonic seconds since the first call, as float.  Weak so that a
// program can interpose its own frGetCurrentTime (SURVEY.md finding 1).
// ================================================================================================
extern "C" __attribute__((weak)) float frGetCurrentTime(void) {
    static bool initialised = false;
    static struct timespec t0;
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    if (!initialised) { initialised = true; t0 = t; }
    uint64_t ns = (uint64_t) (t.tv_sec - t0.tv_sec) * 1000000000ull + (uint64_t) t.tv_nsec - (uint64_t) t0.tv_nsec;
    return (float) ns * (1.0f / 1000000000.0f);
}
