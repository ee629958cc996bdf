/*
 * step_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference world step (c-krit/ferox v0.9.7,
 * frStepWorld, reference src/world.c:204-262) over flat arrays.  It exists to CHECK the CUDA path
 * and to serve as bench.py's `cpu_baseline` / `--impl reference` leg at sizes where the unmodified
 * reference is infeasible (its query dedup is O(FR_WORLD_MAX_OBJECT_COUNT) per body, reference
 * src/broad_phase.c:159-170).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load it.  The product library never links, imports or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py drives this file and the unmodified
 * reference (oracle/_ref/libferox_ref_*.so, compiled from /root/reference/src by oracle/Makefile)
 * over identical scenes and requires bit-identical body state, candidate-pair sequence and contact
 * cache (array order included) for hundreds of steps; tests/golden/ freezes reference outputs.
 *
 * What is restated, and what is deliberately different:
 *   - arithmetic: every float expression keeps the reference's operand order and rounding
 *     (compiled with -ffp-contract=off, no -ffast-math; libm sinf/cosf/sqrtf/floorf; fminf/fmaxf pinned, see pin_fminf).
 *   - broadphase: the reference inserts every body into a hash of cells, then per body gathers,
 *     de-duplicates through a byte array and visits ids ascending (src/broad_phase.c:99-179).  The
 *     visit ORDER and SET are what matter; here the per-body candidate list is produced with a
 *     sorted cell index + sort/unique, which yields the identical sequence in O(K log K).
 *   - contact cache: the reference keeps a stb_ds hash map whose entry array is insertion-ordered
 *     with swap-delete (src/external/stb_ds.h:1436-1448, 1506-1520); that array order IS the
 *     Gauss-Seidel order (src/world.c:237-249).  `cache_put`/`cache_del` reproduce exactly that
 *     array discipline (including the stale tail the eviction loop reads, src/world.c:222-235).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_VERTS 8          /* FR_GEOMETRY_MAX_VERTEX_COUNT, include/ferox.h:43-46 */
#define ORC_BAUMGARTE_FACTOR 0.2f /* include/ferox.h:53-56 */
#define ORC_BAUMGARTE_SLOP 0.01f  /* include/ferox.h:58-61 */
#define ORC_ITERATIONS 10         /* include/ferox.h:68-71 */

enum { SHAPE_UNKNOWN, SHAPE_CIRCLE, SHAPE_POLYGON };          /* include/ferox.h:204-208 */
enum { BODY_UNKNOWN, BODY_STATIC, BODY_KINEMATIC, BODY_DYNAMIC }; /* include/ferox.h:229-234 */

typedef struct { float x, y; } vec2;

typedef struct {
    int type;
    float radius;
    int nv;
    vec2 verts[ORC_MAX_VERTS], normals[ORC_MAX_VERTS];
    float friction, restitution;
} orc_shape;

typedef struct { vec2 pos; float sn, cs, angle; } orc_tx; /* frTransform, include/ferox.h:250-256 */

/* One contact / one manifold, field-for-field the public frContact / frCollision
   (include/ferox.h:163-183) so a cache dump can be compared with the reference's byte for byte. */
typedef struct {
    int id;
    float depth, timestamp;
    vec2 point;
    float normalMass, normalScalar, tangentMass, tangentScalar;
} orc_contact;

typedef struct {
    int count;
    vec2 direction;
    orc_contact contacts[2];
    float friction, restitution;
} orc_collision;

typedef struct { int a, b; orc_collision value; } orc_entry;

typedef struct { int a, b, hit; } orc_candidate;

typedef struct orc_world {
    int n, ns;
    float cell, inv_cell;
    vec2 gravity;
    float timestamp;
    orc_shape *shapes;
    /* per body */
    int *type, *shape;
    orc_tx *tx;
    vec2 *vel, *force;
    float *omega, *torque;
    float *mass, *inv_mass, *inertia, *inv_inertia, *gravity_scale;
    float *aabb; /* x,y,w,h */
    /* contact cache: dense entry array + open-addressing index (key -> array position) */
    orc_entry *entries;
    int n_entries, cap_entries;
    int64_t *ht_key;
    int *ht_val;
    int ht_cap, ht_used;
    /* last step's narrow-phase call log */
    orc_candidate *cand;
    int n_cand, cap_cand;
    /* counters: K cell entries, C candidates, M manifolds, Q contacts */
    long long K, Cn, M, Q;
    /* Gauss-Seidel visiting order of the solver loops (src/world.c:237-249).  0: the reference's (cache array
       order).  1: greedy colour-major (lowest colour free on both bodies the solver writes, taken in array
       order; then colour by colour).  2: a fresh seeded random permutation every step.  Modes 1 and 2 are NOT
       the reference: they exist so that tests can MEASURE, on the reference's own arithmetic, how much a pure
       re-ordering of the same sweeps moves the statistics -- the yardstick for the
       tolerance of the graph-coloured GPU solver. */
    int solve_order;
    uint64_t order_seed;
    int *perm;
    int cap_perm;
    int last_colors;
} orc_world;

/* ---------------------------------------------------------------------------------------------- */
/* vec2 helpers: same formulas as the reference's inline math, include/ferox.h:597-686            */
/* ---------------------------------------------------------------------------------------------- */
/* fminf/fmaxf exactly as the compiled reference evaluates them.  Two facts, both [probe]d here:
   (1) glibc 2.39 on x86-64 implements them around minss/maxss (sysdeps/x86_64/fpu/s_fminf.S,
       s_fmaxf.S): for ordered operands the SECOND argument wins ties, so fmaxf(+0,-0) = -0 and
       fminf(-0,+0) = +0; a NaN operand yields the other operand.
   (2) gcc treats the builtins as commutative and, at -O2, emits the reference's solver calls with
       these operand orders (disassembly of oracle/_ref/obj_NNNN/rigid_body.o, frResolveCollision):
           fminf(slop - depth, 0)  fmaxf(old + dn, 0)  fmaxf(-maxT, old + dt)  fminf(<that>, maxT)
       and src/world.c:398 as fminf(restitution1, restitution2).
   The explicit helpers below pin both, so signed zeros and NaNs come out exactly as the compiled
   reference produces them (they only matter for +-0 / NaN; ordinary values are unaffected). */
static inline float pin_fminf(float x, float y) {
    if (isnan(y)) return x;
    if (isnan(x)) return y;
    return (x < y) ? x : y;
}
static inline float pin_fmaxf(float x, float y) {
    if (isnan(y)) return x;
    if (isnan(x)) return y;
    return (x > y) ? x : y;
}

static inline vec2 v2(float x, float y) { vec2 r = { x, y }; return r; }
static inline vec2 vadd(vec2 a, vec2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline vec2 vsub(vec2 a, vec2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline vec2 vneg(vec2 a) { return v2(-a.x, -a.y); }
static inline vec2 vmul(vec2 a, float k) { return v2(a.x * k, a.y * k); }
static inline float vdot(vec2 a, vec2 b) { return (a.x * b.x) + (a.y * b.y); }
static inline float vcross(vec2 a, vec2 b) { return (a.x * b.y) - (a.y * b.x); }
static inline float vmag2(vec2 a) { return (a.x * a.x) + (a.y * a.y); }
static inline vec2 vnorm(vec2 a) {            /* ferox.h:648-653 */
    float m = sqrtf(vmag2(a));
    return (m > 0.0f) ? vmul(a, 1.0f / m) : a;
}
static inline vec2 vleft(vec2 a) { return vnorm(v2(-a.y, a.x)); }   /* ferox.h:656-658 */
static inline vec2 vright(vec2 a) { return vnorm(v2(a.y, -a.x)); }  /* ferox.h:661-663 */
static inline vec2 vrot(vec2 v, float angle) {                      /* ferox.h:666-672 */
    float s = sinf(angle), c = cosf(angle);
    return v2(v.x * c - v.y * s, v.x * s + v.y * c);
}
static inline vec2 vrot_tx(vec2 v, orc_tx t) {                      /* ferox.h:675-678 */
    return v2(v.x * t.cs - v.y * t.sn, v.x * t.sn + v.y * t.cs);
}
static inline vec2 vxform(vec2 v, orc_tx t) {                       /* ferox.h:681-686 */
    return v2(t.pos.x + (v.x * t.cs - v.y * t.sn), t.pos.y + (v.x * t.sn + v.y * t.cs));
}

/* ---------------------------------------------------------------------------------------------- */
/* narrow phase: reference src/collision.c                                                        */
/* ---------------------------------------------------------------------------------------------- */

/* circle vs circle, src/collision.c:264-298 */
static int collide_circles(const orc_shape *s1, orc_tx t1, const orc_shape *s2, orc_tx t2,
                           orc_collision *out) {
    vec2 d = vsub(t2.pos, t1.pos);
    float rsum = s1->radius + s2->radius;
    float m2 = vmag2(d);
    if (rsum * rsum < m2) return 0;
    float m = sqrtf(m2);
    if (m <= 0.0f) { d.x = 0.0f; d.y = m = FLT_EPSILON; }
    out->direction = vmul(d, 1.0f / m);
    out->contacts[0].point = vxform(vmul(out->direction, s1->radius), t1);
    out->contacts[0].depth = rsum - m;
    out->contacts[1] = out->contacts[0];
    out->count = 1;
    return 1;
}

/* circle vs polygon (either order), src/collision.c:305-449 */
static int collide_circle_poly(const orc_shape *s1, orc_tx t1, const orc_shape *s2, orc_tx t2,
                               orc_collision *out) {
    const orc_shape *circle, *poly;
    orc_tx ct, pt;
    if (s1->type == SHAPE_CIRCLE) { circle = s1; poly = s2; ct = t1; pt = t2; }
    else { circle = s2; poly = s1; ct = t2; pt = t1; }

    vec2 c = vrot(vsub(ct.pos, pt.pos), -pt.angle); /* circle centre in polygon space, :328-330 */
    float radius = circle->radius, max_dot = -FLT_MAX;
    int max_i = -1;
    for (int i = 0; i < poly->nv; i++) {            /* :340-348 */
        float dot = vdot(poly->normals[i], vsub(c, poly->verts[i]));
        if (dot > radius) return 0;
        if (max_dot < dot) { max_dot = dot; max_i = i; }
    }
    if (max_i < 0) return 0;

    vec2 delta = vsub(t2.pos, t1.pos);
    if (max_dot < 0.0f) {                           /* centre inside the polygon, :358-375 */
        out->direction = vneg(vrot_tx(poly->normals[max_i], pt));
        if (vdot(delta, out->direction) < 0.0f) out->direction = vneg(out->direction);
        out->contacts[0].point = vadd(ct.pos, vmul(out->direction, radius));
        out->contacts[0].depth = radius - max_dot;
    } else {
        vec2 p1 = (max_i > 0) ? poly->verts[max_i - 1] : poly->verts[poly->nv - 1];
        vec2 p2 = poly->verts[max_i];
        vec2 edge = vsub(p2, p1);
        vec2 p1c = vsub(c, p1), p2c = vsub(c, p2);
        float d1 = vdot(p1c, edge), d2 = vdot(p2c, vneg(edge));
        if (d1 < 0.0f || d2 < 0.0f) {               /* vertex region, :395-424 */
            vec2 dir = (d1 < 0.0f) ? p1c : p2c;
            float m2 = vmag2(dir);
            if (radius * radius < m2) return 0;
            float m = sqrtf(m2);
            if (m <= 0.0f) m = FLT_EPSILON;
            out->direction = vmul(vrot_tx(vneg(dir), pt), 1.0f / m);
            if (vdot(delta, out->direction) < 0.0f) out->direction = vneg(out->direction);
            out->contacts[0].point = vxform(vmul(out->direction, radius), ct);
            out->contacts[0].depth = radius - m;
        } else {                                    /* face region, :425-445 */
            out->direction = vneg(vrot_tx(poly->normals[max_i], pt));
            if (vdot(delta, out->direction) < 0.0f) out->direction = vneg(out->direction);
            out->contacts[0].point = vadd(ct.pos, vmul(out->direction, radius));
            out->contacts[0].depth = radius - max_dot;
        }
    }
    out->contacts[1] = out->contacts[0];
    out->count = 1;
    return 1;
}

/* farthest vertex along v (first maximum wins), src/collision.c:708-724 */
static int support_index(const orc_shape *s, orc_tx t, vec2 v) {
    float max_dot = -FLT_MAX;
    int max_i = -1;
    v = vrot(v, -t.angle);
    for (int i = 0; i < s->nv; i++) {
        float dot = vdot(s->verts[i], v);
        if (max_dot < dot) { max_dot = dot; max_i = i; }
    }
    return max_i;
}

/* axis of least penetration of s1's faces against s2, src/collision.c:669-705 */
static int separating_axis(const orc_shape *s1, orc_tx t1, const orc_shape *s2, orc_tx t2,
                           float *depth) {
    float max_depth = -FLT_MAX;
    int max_i = -1;
    for (int i = 0; i < s1->nv; i++) {
        vec2 vertex = vxform(s1->verts[i], t1);
        vec2 normal = vrot_tx(s1->normals[i], t1);
        int si = support_index(s2, t2, vneg(normal));
        if (si < 0) return si;
        vec2 sp = vxform(s2->verts[si], t2);
        float d = vdot(normal, vsub(sp, vertex));
        if (max_depth < d) { max_depth = d; max_i = i; }
    }
    *depth = max_depth;
    return max_i;
}

typedef struct { vec2 data[3]; int idx[2]; } orc_edge;

/* edge of s most perpendicular to v around the support vertex, src/collision.c:628-666 */
static orc_edge contact_edge(const orc_shape *s, orc_tx t, vec2 v) {
    int si = support_index(s, t, v);
    int prev = (si == 0) ? s->nv - 1 : si - 1;
    int next = (si == s->nv - 1) ? 0 : si + 1;
    vec2 pe = vnorm(vsub(s->verts[si], s->verts[prev]));
    vec2 ne = vnorm(vsub(s->verts[si], s->verts[next]));
    v = vrot(v, -t.angle);
    vec2 sv = vxform(s->verts[si], t);
    orc_edge e;
    if (vdot(pe, v) < vdot(ne, v)) {
        e.data[0] = vxform(s->verts[prev], t); e.data[1] = sv; e.data[2] = sv;
        e.idx[0] = prev; e.idx[1] = si;
    } else {
        e.data[0] = sv; e.data[1] = vxform(s->verts[next], t); e.data[2] = sv;
        e.idx[0] = si; e.idx[1] = next;
    }
    return e;
}

/* keep the part of e with dot(p, v) >= dot; indices are NOT swapped, src/collision.c:228-257 */
static int clip_edge(orc_edge *e, vec2 v, float dot) {
    float d1 = vdot(e->data[0], v) - dot;
    float d2 = vdot(e->data[1], v) - dot;
    if (d1 >= 0.0f && d2 >= 0.0f) return 1;
    vec2 ev = vsub(e->data[1], e->data[0]);
    vec2 mid = vadd(e->data[0], vmul(ev, (d1 / (d1 - d2))));
    if (d1 > 0.0f && d2 < 0.0f) { e->data[1] = mid; return 1; }
    if (d1 < 0.0f && d2 > 0.0f) { e->data[0] = e->data[1]; e->data[1] = mid; return 1; }
    return 0;
}

/* polygon vs polygon: SAT + clipping, src/collision.c:456-558 */
static int collide_polys(const orc_shape *s1, orc_tx t1, const orc_shape *s2, orc_tx t2,
                         orc_collision *out) {
    float dep1 = FLT_MAX, dep2 = FLT_MAX;
    int i1 = separating_axis(s1, t1, s2, t2, &dep1);
    if (dep1 >= 0.0f) return 0;
    int i2 = separating_axis(s2, t2, s1, t1, &dep2);
    if (dep2 >= 0.0f) return 0;

    vec2 dir = (dep1 > dep2) ? vrot_tx(s1->normals[i1], t1) : vrot_tx(s2->normals[i2], t2);
    vec2 delta = vsub(t2.pos, t1.pos);
    if (vdot(delta, dir) < 0.0f) dir = vneg(dir);

    orc_edge e1 = contact_edge(s1, t1, dir);
    orc_edge e2 = contact_edge(s2, t2, vneg(dir));
    orc_edge ref = e1, inc = e2;
    float ed1 = vdot(vsub(e1.data[1], e1.data[0]), dir);
    float ed2 = vdot(vsub(e2.data[1], e2.data[0]), dir);
    int flipped = 0;
    if (fabsf(ed1) > fabsf(ed2)) { ref = e2; inc = e1; flipped = 1; }   /* :497-502 */

    vec2 rv = vnorm(vsub(ref.data[1], ref.data[0]));
    float rd1 = vdot(ref.data[0], rv), rd2 = vdot(ref.data[1], rv);
    if (!clip_edge(&inc, rv, rd1)) return 0;
    if (!clip_edge(&inc, vneg(rv), -rd2)) return 0;

    vec2 rn = vright(rv);
    float max_depth = vdot(ref.data[2], rn);
    float depth1 = vdot(inc.data[0], rn) - max_depth;
    float depth2 = vdot(inc.data[1], rn) - max_depth;

    out->direction = dir;
    unsigned mask = ((unsigned) flipped << 16) | ((unsigned) ref.idx[0] << 8);   /* :524-528 */
    out->contacts[0].id = (int) (mask | (unsigned) inc.idx[0]);
    out->contacts[1].id = (int) (mask | (unsigned) inc.idx[1]);
    if (depth1 < 0.0f) {
        out->contacts[0].id = out->contacts[1].id;
        out->contacts[0].point = inc.data[1];
        out->contacts[0].depth = depth2;
        out->contacts[1] = out->contacts[0];
        out->count = 1;
    } else if (depth2 < 0.0f) {
        out->contacts[0].point = inc.data[0];
        out->contacts[0].depth = depth1;
        out->contacts[1] = out->contacts[0];
        out->count = 1;
    } else {
        out->contacts[0].point = inc.data[0];
        out->contacts[0].depth = depth1;
        out->contacts[1].point = inc.data[1];
        out->contacts[1].depth = depth2;
        out->count = 2;
    }
    return 1;
}

/* dispatch, src/collision.c:114-135; `out` is only written on a hit */
static int compute_collision(const orc_world *w, int a, int b, orc_collision *out) {
    if (w->shape[a] < 0 || w->shape[b] < 0) return 0;
    const orc_shape *s1 = &w->shapes[w->shape[a]], *s2 = &w->shapes[w->shape[b]];
    orc_tx t1 = w->tx[a], t2 = w->tx[b];
    if (s1->type == SHAPE_CIRCLE && s2->type == SHAPE_CIRCLE) return collide_circles(s1, t1, s2, t2, out);
    if ((s1->type == SHAPE_CIRCLE && s2->type == SHAPE_POLYGON)
        || (s1->type == SHAPE_POLYGON && s2->type == SHAPE_CIRCLE))
        return collide_circle_poly(s1, t1, s2, t2, out);
    if (s1->type == SHAPE_POLYGON && s2->type == SHAPE_POLYGON) return collide_polys(s1, t1, s2, t2, out);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* contact cache with stb_ds array discipline                                                     */
/* ---------------------------------------------------------------------------------------------- */
static inline int64_t pair_key(int a, int b) { return ((int64_t) a << 32) | (uint32_t) b; }
static inline uint32_t key_hash(int64_t k) {
    uint64_t z = (uint64_t) k * 0x9E3779B97F4A7C15ull;
    return (uint32_t) (z >> 29);
}

static void ht_rebuild(orc_world *w, int cap) {
    free(w->ht_key); free(w->ht_val);
    w->ht_cap = cap; w->ht_used = 0;
    w->ht_key = malloc(sizeof(int64_t) * cap);
    w->ht_val = malloc(sizeof(int) * cap);
    for (int i = 0; i < cap; i++) w->ht_key[i] = -1;
    for (int p = 0; p < w->n_entries; p++) {
        int64_t k = pair_key(w->entries[p].a, w->entries[p].b);
        uint32_t s = key_hash(k) & (cap - 1);
        while (w->ht_key[s] != -1) s = (s + 1) & (cap - 1);
        w->ht_key[s] = k; w->ht_val[s] = p; w->ht_used++;
    }
}

static int ht_find_slot(const orc_world *w, int64_t k) {
    if (w->ht_cap == 0) return -1;
    uint32_t s = key_hash(k) & (w->ht_cap - 1);
    while (w->ht_key[s] != -1) {
        if (w->ht_key[s] == k) return (int) s;
        s = (s + 1) & (w->ht_cap - 1);
    }
    return -1;
}

static orc_entry *cache_get(orc_world *w, int a, int b) {
    int s = ht_find_slot(w, pair_key(a, b));
    return (s < 0) ? NULL : &w->entries[w->ht_val[s]];
}

/* hmputs: overwrite in place, or append at the end (stb_ds.h:1436-1448) */
static void cache_put(orc_world *w, int a, int b, const orc_collision *c) {
    orc_entry *e = cache_get(w, a, b);
    if (e != NULL) { e->value = *c; return; }
    if (w->n_entries + 1 >= w->cap_entries) {
        w->cap_entries = w->cap_entries ? w->cap_entries * 2 : 1024;
        w->entries = realloc(w->entries, sizeof(orc_entry) * w->cap_entries);
    }
    if ((w->ht_used + 1) * 2 > w->ht_cap) ht_rebuild(w, w->ht_cap ? w->ht_cap * 2 : 2048);
    int p = w->n_entries++;
    w->entries[p].a = a; w->entries[p].b = b; w->entries[p].value = *c;
    int64_t k = pair_key(a, b);
    uint32_t s = key_hash(k) & (w->ht_cap - 1);
    while (w->ht_key[s] != -1) s = (s + 1) & (w->ht_cap - 1);
    w->ht_key[s] = k; w->ht_val[s] = p; w->ht_used++;
}

/* hmdel: swap-delete; the vacated tail slot keeps its old bytes (stb_ds.h:1506-1520) */
static void cache_del(orc_world *w, int a, int b) {
    int s = ht_find_slot(w, pair_key(a, b));
    if (s < 0) return;
    int old_index = w->ht_val[s], final_index = w->n_entries - 1;
    /* backward-shift deletion keeps the probe sequences intact */
    uint32_t mask = w->ht_cap - 1, hole = (uint32_t) s, j = (uint32_t) s;
    for (;;) {
        j = (j + 1) & mask;
        if (w->ht_key[j] == -1) break;
        uint32_t home = key_hash(w->ht_key[j]) & mask;
        if (((j - home) & mask) >= ((j - hole) & mask)) {
            w->ht_key[hole] = w->ht_key[j]; w->ht_val[hole] = w->ht_val[j]; hole = j;
        }
    }
    w->ht_key[hole] = -1; w->ht_used--;
    if (old_index != final_index) {
        w->entries[old_index] = w->entries[final_index];
        int s2 = ht_find_slot(w, pair_key(w->entries[old_index].a, w->entries[old_index].b));
        w->ht_val[s2] = old_index;
    }
    w->n_entries--;
}

/* ---------------------------------------------------------------------------------------------- */
/* broadphase: candidate sequence of the reference (src/world.c:432-444, src/broad_phase.c)       */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t cell; int body; } cell_entry;

static int cmp_cell_entry(const void *pa, const void *pb) {
    const cell_entry *a = pa, *b = pb;
    if (a->cell != b->cell) return (a->cell < b->cell) ? -1 : 1;
    return (a->body > b->body) - (a->body < b->body);
}
static int cmp_int(const void *pa, const void *pb) {
    int a = *(const int *) pa, b = *(const int *) pb;
    return (a > b) - (a < b);
}
static inline int64_t cell_key(int x, int y) { return ((int64_t) y << 32) | (uint32_t) x; }

static inline void cell_rect(const orc_world *w, int i, int *x0, int *y0, int *x1, int *y1) {
    const float *a = &w->aabb[4 * i];
    /* C float->int conversion truncates toward zero, src/broad_phase.c:104-108 */
    *x0 = (int) (a[0] * w->inv_cell);
    *y0 = (int) (a[1] * w->inv_cell);
    *x1 = (int) ((a[0] + a[2]) * w->inv_cell);
    *y1 = (int) ((a[1] + a[3]) * w->inv_cell);
}

/* the pre-step callback body, src/world.c:328-410 */
static void on_candidate(orc_world *w, int a, int b) {
    if (w->inv_mass[a] + w->inv_mass[b] <= 0.0f) return;               /* :340-341 */
    orc_collision col;
    memset(&col, 0, sizeof col);
    int hit = compute_collision(w, a, b, &col);
    if (w->n_cand == w->cap_cand) {
        w->cap_cand = w->cap_cand ? w->cap_cand * 2 : 4096;
        w->cand = realloc(w->cand, sizeof(orc_candidate) * w->cap_cand);
    }
    w->cand[w->n_cand].a = a; w->cand[w->n_cand].b = b; w->cand[w->n_cand].hit = hit; w->n_cand++;
    if (!hit) { cache_del(w, a, b); return; }                            /* :347-356 */
    orc_entry *old = cache_get(w, a, b);
    for (int i = 0; i < col.count; i++) col.contacts[i].timestamp = w->timestamp;   /* :362-363 */
    if (old != NULL) {                                                   /* :365-394 */
        col.friction = old->value.friction;
        col.restitution = old->value.restitution;
        for (int i = 0; i < col.count; i++) {
            int oi = -1;
            for (int j = 0; j < old->value.count; j++)
                if (col.contacts[i].id == old->value.contacts[j].id) { oi = j; break; }
            if (oi < 0) continue;
            col.contacts[i].normalScalar = old->value.contacts[oi].normalScalar;
            col.contacts[i].tangentScalar = old->value.contacts[oi].tangentScalar;
        }
    } else {                                                             /* :395-404 */
        const orc_shape *s1 = &w->shapes[w->shape[a]], *s2 = &w->shapes[w->shape[b]];
        col.friction = 0.5f * (s1->friction + s2->friction);
        col.restitution = pin_fminf(s1->restitution, s2->restitution);
        if (col.friction < 0.0f) col.friction = 0.0f;
        if (col.restitution < 0.0f) col.restitution = 0.0f;
    }
    cache_put(w, a, b, &col);                                            /* :406-407 */
}

static void pre_step(orc_world *w) {
    int n = w->n;
    w->n_cand = 0;
    /* count + emit (cell, body) entries: src/broad_phase.c:99-127 */
    long long total = 0;
    for (int i = 0; i < n; i++) {
        int x0, y0, x1, y1;
        cell_rect(w, i, &x0, &y0, &x1, &y1);
        if (x1 >= x0 && y1 >= y0) total += (long long) (x1 - x0 + 1) * (y1 - y0 + 1);
    }
    w->K = total;
    cell_entry *ent = malloc(sizeof(cell_entry) * (size_t) (total ? total : 1));
    long long k = 0;
    for (int i = 0; i < n; i++) {
        int x0, y0, x1, y1;
        cell_rect(w, i, &x0, &y0, &x1, &y1);
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) { ent[k].cell = cell_key(x, y); ent[k].body = i; k++; }
    }
    qsort(ent, (size_t) total, sizeof(cell_entry), cmp_cell_entry);
    /* unique cell directory for binary search */
    long long ncell = 0;
    for (long long e = 0; e < total; e++) if (e == 0 || ent[e].cell != ent[e - 1].cell) ncell++;
    int64_t *dir_key = malloc(sizeof(int64_t) * (size_t) (ncell ? ncell : 1));
    long long *dir_off = malloc(sizeof(long long) * (size_t) (ncell + 1));
    ncell = 0;
    for (long long e = 0; e < total; e++)
        if (e == 0 || ent[e].cell != ent[e - 1].cell) { dir_key[ncell] = ent[e].cell; dir_off[ncell] = e; ncell++; }
    dir_off[ncell] = total;

    int *tmp = NULL;
    size_t tmp_cap = 0;
    /* query every body in index order; visit unique ids ascending: src/broad_phase.c:130-179 */
    for (int i = 0; i < n; i++) {
        int x0, y0, x1, y1;
        cell_rect(w, i, &x0, &y0, &x1, &y1);
        size_t m = 0;
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                int64_t key = cell_key(x, y);
                long long lo = 0, hi = ncell;
                while (lo < hi) { long long mid = (lo + hi) >> 1; if (dir_key[mid] < key) lo = mid + 1; else hi = mid; }
                if (lo >= ncell || dir_key[lo] != key) continue;
                for (long long e = dir_off[lo]; e < dir_off[lo + 1]; e++) {
                    int j = ent[e].body;
                    if (j <= i) continue;                                /* src/world.c:333 */
                    if (m == tmp_cap) { tmp_cap = tmp_cap ? tmp_cap * 2 : 256; tmp = realloc(tmp, sizeof(int) * tmp_cap); }
                    tmp[m++] = j;
                }
            }
        if (m == 0) continue;
        qsort(tmp, m, sizeof(int), cmp_int);
        for (size_t q = 0; q < m; q++) {
            if (q > 0 && tmp[q] == tmp[q - 1]) continue;
            on_candidate(w, i, tmp[q]);
        }
    }
    free(tmp); free(dir_key); free(dir_off); free(ent);
    w->Cn = w->n_cand;
}

/* ---------------------------------------------------------------------------------------------- */
/* solver + integration: reference src/rigid_body.c                                               */
/* ---------------------------------------------------------------------------------------------- */

/* frApplyAccumulatedImpulses, src/rigid_body.c:333-411 */
static void warm_start(orc_world *w, orc_entry *e) {
    int a = e->a, b = e->b;
    orc_collision *c = &e->value;
    if (w->inv_mass[a] + w->inv_mass[b] <= 0.0f) {
        if (w->type[a] == BODY_STATIC) { w->vel[a].x = w->vel[a].y = w->omega[a] = 0.0f; }
        if (w->type[b] == BODY_STATIC) { w->vel[b].x = w->vel[b].y = w->omega[b] = 0.0f; }
        return;
    }
    vec2 tangent = vright(c->direction);
    for (int i = 0; i < c->count; i++) {
        orc_contact *ct = &c->contacts[i];
        vec2 r1 = vsub(ct->point, w->tx[a].pos), r2 = vsub(ct->point, w->tx[b].pos);
        float c1 = vcross(r1, c->direction), c2 = vcross(r2, c->direction);
        float nm = (w->inv_mass[a] + w->inv_mass[b]) + w->inv_inertia[a] * (c1 * c1)
                   + w->inv_inertia[b] * (c2 * c2);
        ct->normalMass = 1.0f / nm;
        vec2 acc_n = vmul(c->direction, ct->normalScalar);
        c1 = vcross(r1, tangent); c2 = vcross(r2, tangent);
        float tm = (w->inv_mass[a] + w->inv_mass[b]) + w->inv_inertia[a] * (c1 * c1)
                   + w->inv_inertia[b] * (c2 * c2);
        ct->tangentMass = 1.0f / tm;
        vec2 acc_t = vmul(tangent, ct->tangentScalar);
        vec2 acc = vadd(acc_n, acc_t);
        w->vel[a] = vsub(w->vel[a], vmul(acc, w->inv_mass[a]));
        w->omega[a] -= w->inv_inertia[a] * vcross(r1, acc);
        w->vel[b] = vadd(w->vel[b], vmul(acc, w->inv_mass[b]));
        w->omega[b] += w->inv_inertia[b] * vcross(r2, acc);
    }
}

/* frResolveCollision, src/rigid_body.c:446-551 */
static void resolve(orc_world *w, orc_entry *e, float inv_dt) {
    int a = e->a, b = e->b;
    orc_collision *c = &e->value;
    if (w->inv_mass[a] + w->inv_mass[b] <= 0.0f || inv_dt <= 0.0f) return;
    vec2 tangent = vright(c->direction);
    for (int i = 0; i < c->count; i++) {
        orc_contact *ct = &c->contacts[i];
        vec2 r1 = vsub(ct->point, w->tx[a].pos), r2 = vsub(ct->point, w->tx[b].pos);
        vec2 n1 = vleft(r1), n2 = vleft(r2);      /* UNIT perpendiculars, :465-466 */
        vec2 rv = vsub(vadd(w->vel[b], vmul(n2, w->omega[b])), vadd(w->vel[a], vmul(n1, w->omega[a])));
        float vn = vdot(rv, c->direction);
        float bias = -(ORC_BAUMGARTE_FACTOR * inv_dt) * pin_fminf(-ct->depth + ORC_BAUMGARTE_SLOP, 0.0f);
        float ns = ((-(1.0f + c->restitution) * vn) + bias) * ct->normalMass;
        {
            float old = ct->normalScalar;
            ct->normalScalar = pin_fmaxf(old + ns, 0.0f);
            ns = ct->normalScalar - old;
        }
        vec2 n_imp = vmul(c->direction, ns);
        /* recomputed BEFORE the normal impulse is applied, :501-507 */
        rv = vsub(vadd(w->vel[b], vmul(n2, w->omega[b])), vadd(w->vel[a], vmul(n1, w->omega[a])));
        float ts = -vdot(rv, tangent) * ct->tangentMass;
        {
            float max_t = fabsf(c->friction * ct->normalScalar);
            float old = ct->tangentScalar;
            ct->tangentScalar = pin_fminf(pin_fmaxf(-max_t, old + ts), max_t);
            ts = ct->tangentScalar - old;
        }
        vec2 t_imp = vmul(tangent, ts);
        vec2 tot = vadd(n_imp, t_imp);
        w->vel[a] = vsub(w->vel[a], vmul(tot, w->inv_mass[a]));
        w->omega[a] -= w->inv_inertia[a] * vcross(r1, tot);
        w->vel[b] = vadd(w->vel[b], vmul(tot, w->inv_mass[b]));
        w->omega[b] += w->inv_inertia[b] * vcross(r2, tot);
    }
}

/* frGetShapeAABB, src/geometry.c:183-216 */
static void shape_aabb(const orc_world *w, int i, float *out) {
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    if (w->shape[i] < 0) return;
    const orc_shape *s = &w->shapes[w->shape[i]];
    orc_tx t = w->tx[i];
    if (s->type == SHAPE_CIRCLE) {
        out[0] = t.pos.x - s->radius; out[1] = t.pos.y - s->radius;
        out[2] = out[3] = 2.0f * s->radius;
    } else if (s->type == SHAPE_POLYGON) {
        vec2 lo = { FLT_MAX, FLT_MAX }, hi = { -FLT_MAX, -FLT_MAX };
        for (int k = 0; k < s->nv; k++) {
            vec2 v = vxform(s->verts[k], t);
            if (lo.x > v.x) lo.x = v.x;
            if (lo.y > v.y) lo.y = v.y;
            if (hi.x < v.x) hi.x = v.x;
            if (hi.y < v.y) hi.y = v.y;
        }
        out[0] = lo.x; out[1] = lo.y; out[2] = hi.x - lo.x; out[3] = hi.y - lo.y;
    }
}

/* frNormalizeAngle, src/rigid_body.c:589-591: the inner sum and product are evaluated in double
   (M_PI is a double constant), floorf rounds its argument to float first. */
static const float ORC_TWO_PI = 2.0f * M_PI, ORC_INV_TWO_PI = 1.0f / (2.0f * M_PI); /* :54 */
static inline float normalize_angle(float angle) {
    return angle - (ORC_TWO_PI * floorf((angle + -M_PI) * ORC_INV_TWO_PI));
}

/* frIntegrateForBodyPosition -> frSetBodyAngle, src/rigid_body.c:433-443, 225-238 */
static void integrate_position(orc_world *w, int i, float dt) {
    if (w->type[i] == BODY_STATIC || dt <= 0.0f) return;
    w->tx[i].pos.x += w->vel[i].x * dt;
    w->tx[i].pos.y += w->vel[i].y * dt;
    if (w->omega[i] != 0.0f) {
        float angle = w->tx[i].angle + (w->omega[i] * dt);
        if (w->tx[i].angle != angle) {
            w->tx[i].angle = normalize_angle(angle);
            w->tx[i].sn = sinf(w->tx[i].angle);
            w->tx[i].cs = cosf(w->tx[i].angle);
        }
    }
    shape_aabb(w, i, &w->aabb[4 * i]);
}

/* ---------------------------------------------------------------------------------------------- */
/* public C surface (bound from Python with ctypes)                                               */
/* ---------------------------------------------------------------------------------------------- */
orc_world *orc_create(int n, int ns, float cell, float gx, float gy) {
    orc_world *w = calloc(1, sizeof *w);
    w->n = n; w->ns = ns; w->cell = cell; w->inv_cell = 1.0f / cell;   /* src/broad_phase.c:62-63 */
    w->gravity = v2(gx, gy);
    w->shapes = calloc((size_t) (ns ? ns : 1), sizeof(orc_shape));
    size_t m = (size_t) (n ? n : 1);
    w->type = calloc(m, sizeof(int)); w->shape = calloc(m, sizeof(int));
    w->tx = calloc(m, sizeof(orc_tx)); w->vel = calloc(m, sizeof(vec2)); w->force = calloc(m, sizeof(vec2));
    w->omega = calloc(m, sizeof(float)); w->torque = calloc(m, sizeof(float));
    w->mass = calloc(m, sizeof(float)); w->inv_mass = calloc(m, sizeof(float));
    w->inertia = calloc(m, sizeof(float)); w->inv_inertia = calloc(m, sizeof(float));
    w->gravity_scale = calloc(m, sizeof(float)); w->aabb = calloc(4 * m, sizeof(float));
    return w;
}

void orc_destroy(orc_world *w) {
    if (!w) return;
    free(w->shapes); free(w->type); free(w->shape); free(w->tx); free(w->vel); free(w->force);
    free(w->omega); free(w->torque); free(w->mass); free(w->inv_mass); free(w->inertia);
    free(w->inv_inertia); free(w->gravity_scale); free(w->aabb); free(w->entries);
    free(w->ht_key); free(w->ht_val); free(w->cand); free(w->perm); free(w);
}

void orc_set_shape(orc_world *w, int s, int type, float radius, int nv, const float *verts,
                   const float *normals, float friction, float restitution) {
    orc_shape *sh = &w->shapes[s];
    sh->type = type; sh->radius = radius; sh->nv = nv; sh->friction = friction; sh->restitution = restitution;
    for (int k = 0; k < nv && k < ORC_MAX_VERTS; k++) {
        sh->verts[k] = v2(verts[2 * k], verts[2 * k + 1]);
        sh->normals[k] = v2(normals[2 * k], normals[2 * k + 1]);
    }
}

void orc_set_bodies(orc_world *w, const int *type, const int *shape, const float *pos,
                    const float *angle, const float *sincos, const float *vel, const float *omega,
                    const float *mass4, const float *gravity_scale, const float *aabb) {
    for (int i = 0; i < w->n; i++) {
        w->type[i] = type[i]; w->shape[i] = shape[i];
        w->tx[i].pos = v2(pos[2 * i], pos[2 * i + 1]);
        w->tx[i].angle = angle[i]; w->tx[i].sn = sincos[2 * i]; w->tx[i].cs = sincos[2 * i + 1];
        w->vel[i] = v2(vel[2 * i], vel[2 * i + 1]); w->omega[i] = omega[i];
        w->mass[i] = mass4[4 * i]; w->inv_mass[i] = mass4[4 * i + 1];
        w->inertia[i] = mass4[4 * i + 2]; w->inv_inertia[i] = mass4[4 * i + 3];
        w->gravity_scale[i] = gravity_scale[i];
        memcpy(&w->aabb[4 * i], &aabb[4 * i], 4 * sizeof(float));
    }
}

void orc_set_forces(orc_world *w, const float *force, const float *torque) {
    for (int i = 0; i < w->n; i++) { w->force[i] = v2(force[2 * i], force[2 * i + 1]); w->torque[i] = torque[i]; }
}

void orc_set_timestamp(orc_world *w, float ts) { w->timestamp = ts; }
void orc_set_solve_order(orc_world *w, int mode, unsigned long long seed) { w->solve_order = mode; w->order_seed = seed; }
int orc_last_colors(const orc_world *w) { return w->last_colors; }

/* visiting order of the solver loops for solve_order != 0 (NULL: array order, the reference's) */
static const int *solve_permutation(orc_world *w) {
    const int m = w->n_entries;
    if (w->solve_order == 0 || m == 0) return NULL;
    if (w->cap_perm < m) { w->cap_perm = m + m / 2 + 16; w->perm = realloc(w->perm, sizeof(int) * (size_t) w->cap_perm); }
    if (w->solve_order == 2) {
        for (int j = 0; j < m; j++) w->perm[j] = j;
        for (int j = m - 1; j > 0; j--) {                                  /* Fisher-Yates on splitmix64 */
            uint64_t z = (w->order_seed += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            int k = (int) (z % (uint64_t) (j + 1)), t = w->perm[j];
            w->perm[j] = w->perm[k]; w->perm[k] = t;
        }
        return w->perm;
    }
    /* greedy colouring: only bodies the solver writes (invMass > 0 or invInertia > 0) constrain it (App. A.13) */
    uint64_t *used = calloc((size_t) w->n, sizeof(uint64_t));
    int *col = malloc(sizeof(int) * (size_t) m), cnt[65] = { 0 }, ncol = 0;
    for (int j = 0; j < m; j++) {
        const int a = w->entries[j].a, b = w->entries[j].b;
        const int wa = w->inv_mass[a] > 0.0f || w->inv_inertia[a] > 0.0f, wb = w->inv_mass[b] > 0.0f || w->inv_inertia[b] > 0.0f;
        uint64_t u = (wa ? used[a] : 0) | (wb ? used[b] : 0);
        int c = 0;
        while (c < 63 && ((u >> c) & 1)) c++;
        col[j] = c;
        if (wa) used[a] |= 1ull << c;
        if (wb) used[b] |= 1ull << c;
        cnt[c + 1]++;
        if (c + 1 > ncol) ncol = c + 1;
    }
    for (int c = 0; c < 64; c++) cnt[c + 1] += cnt[c];
    for (int j = 0; j < m; j++) w->perm[cnt[col[j]]++] = j;
    w->last_colors = ncol;
    free(col); free(used);
    return w->perm;
}

/* frStepWorld, src/world.c:204-262 (no collision handler; the world is fixed-size here) */
void orc_step(orc_world *w, float dt) {
    if (w == NULL || dt <= 0.0f) return;
    pre_step(w);                                                          /* :207 */
    for (int i = 0; i < w->n; i++) {                                      /* :216-220 */
        if (w->mass[i] > 0.0f)                                            /* rigid_body.c:309-316 */
            w->force[i] = vadd(w->force[i], vmul(w->gravity, w->gravity_scale[i] * w->mass[i]));
        if (w->inv_mass[i] > 0.0f) {                                      /* rigid_body.c:418-427 */
            w->vel[i] = vadd(w->vel[i], vmul(w->force[i], w->inv_mass[i] * dt));
            w->omega[i] += (w->torque[i] * w->inv_inertia[i]) * dt;
        }
    }
    {                                                                     /* :222-235, quirks kept */
        int entry_count = w->n_entries;
        for (int j = 0; j < entry_count; j++) {
            int a = w->entries[j].a, b = w->entries[j].b;
            const orc_collision *value = &w->entries[j].value;
            for (int k = 0; k < value->count; k++)
                if (w->timestamp - value->contacts[k].timestamp > dt) { cache_del(w, a, b); break; }
        }
    }
    const int *perm = solve_permutation(w);
    for (int j = 0; j < w->n_entries; j++) warm_start(w, &w->entries[perm ? perm[j] : j]); /* :237-240 */
    float inv_dt = 1.0f / dt;
    for (int it = 0; it < ORC_ITERATIONS; it++)                           /* :242-249 */
        for (int j = 0; j < w->n_entries; j++) resolve(w, &w->entries[perm ? perm[j] : j], inv_dt);
    for (int i = 0; i < w->n; i++) integrate_position(w, i, dt);          /* :251-252 */
    for (int i = 0; i < w->n; i++) { w->force[i].x = w->force[i].y = w->torque[i] = 0.0f; } /* :481-482 */
    w->M = w->n_entries; w->Q = 0;
    for (int j = 0; j < w->n_entries; j++) w->Q += w->entries[j].value.count;
}

void orc_get_state(const orc_world *w, float *pos, float *angle, float *sincos, float *vel,
                   float *omega, float *aabb) {
    for (int i = 0; i < w->n; i++) {
        pos[2 * i] = w->tx[i].pos.x; pos[2 * i + 1] = w->tx[i].pos.y;
        angle[i] = w->tx[i].angle; sincos[2 * i] = w->tx[i].sn; sincos[2 * i + 1] = w->tx[i].cs;
        vel[2 * i] = w->vel[i].x; vel[2 * i + 1] = w->vel[i].y; omega[i] = w->omega[i];
        memcpy(&aabb[4 * i], &w->aabb[4 * i], 4 * sizeof(float));
    }
}

int orc_candidate_count(const orc_world *w) { return w->n_cand; }
void orc_get_candidates(const orc_world *w, int *out3) { memcpy(out3, w->cand, sizeof(orc_candidate) * (size_t) w->n_cand); }
int orc_cache_count(const orc_world *w) { return w->n_entries; }
int orc_entry_size(void) { return (int) sizeof(orc_entry); }
void orc_get_cache(const orc_world *w, void *out) { memcpy(out, w->entries, sizeof(orc_entry) * (size_t) w->n_entries); }
void orc_get_counters(const orc_world *w, long long *out4) { out4[0] = w->K; out4[1] = w->Cn; out4[2] = w->M; out4[3] = w->Q; }

/* unit-level hooks mirroring the public frApplyAccumulatedImpulses / frResolveCollision */
void orc_warm_start_pair(orc_world *w, int a, int b, void *collision) {
    orc_entry e; e.a = a; e.b = b; memcpy(&e.value, collision, sizeof e.value);
    warm_start(w, &e);
    memcpy(collision, &e.value, sizeof e.value);
}
void orc_resolve_pair(orc_world *w, int a, int b, void *collision, float inv_dt) {
    orc_entry e; e.a = a; e.b = b; memcpy(&e.value, collision, sizeof e.value);
    resolve(w, &e, inv_dt);
    memcpy(collision, &e.value, sizeof e.value);
}

/* standalone narrow phase on body indices of the world: the unit-level hook (frComputeCollision) */
int orc_collide(const orc_world *w, int a, int b, void *out_collision) {
    orc_collision col;
    memset(&col, 0, sizeof col);
    int hit = compute_collision(w, a, b, &col);
    memcpy(out_collision, &col, sizeof col);
    return hit;
}
