// fx_bodies.cu -- per-body kernels (K10), step bookkeeping and the single-item entry points that back
// the public per-body / per-pair functions of ferox.h.
//
// k_integrate_position replaces reference src/world.c:251-252 + :481-482: frIntegrateForBodyPosition
// (src/rigid_body.c:433-443) -> frSetBodyAngle (:225-238: normalise into [pi, 3*pi), sinf, cosf) ->
// frGetShapeAABB (src/geometry.c:183-216), fused with the force clear of frPostStepWorld.
// 116 B read + 56 B written per body (+ polygon vertices).
#include "fx_device.h"
#include "fx_kernels.h"
extern long long g_fx_launches;
#include "fx_narrow.cuh"
#include "fx_solver.cuh"
#include "fx_raycast.cuh"
#include "fx_scan.cuh"

#define BD_BLOCK 256

__device__ __forceinline__ float4 shape_aabb(const ShapeDev *s, Xf t) {
    if (s == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (s->type == FX_SHAPE_CIRCLE) {
        float r = s->radius;
        return make_float4(t.p.x - r, t.p.y - r, 2.0f * r, 2.0f * r);
    }
    if (s->type == FX_SHAPE_POLYGON) {
        float lox = FLT_MAX, loy = FLT_MAX, hix = -FLT_MAX, hiy = -FLT_MAX;
        const int nv = s->nv;
        for (int k = 0; k < nv; ++k) {
            V2 v = xf_apply(s->verts[k], t);
            if (lox > v.x) lox = v.x;
            if (loy > v.y) loy = v.y;
            if (hix < v.x) hix = v.x;
            if (hiy < v.y) hiy = v.y;
        }
        return make_float4(lox, loy, hix - lox, hiy - loy);
    }
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void integrate_position_one(const DevWorld &w, int i, float dt) {
    const int type = __float_as_int(w.mprop[i].w) & 0xff;
    if (type != FX_BODY_STATIC && dt > 0.0f) {
        float4 pr = w.posrot[i];
        const float4 v = w.vel[i];
        pr.x += v.x * dt;
        pr.y += v.y * dt;
        if (v.z != 0.0f) {
            const float old = w.angle[i];
            const float target = old + (v.z * dt);
            if (old != target) {                       // frSetBodyAngle early-out, rigid_body.c:226
                const float a = fx_normalize_angle(target);
                SinCos sc = fx_sincosf(a);
                w.angle[i] = a;
                pr.z = sc.s;
                pr.w = sc.c;
            }
        }
        w.posrot[i] = pr;
        const int si = w.shape[i].x;
        Xf t;
        t.p = v2(pr.x, pr.y); t.sn = pr.z; t.cs = pr.w;
        w.aabb[i] = shape_aabb(si >= 0 ? &w.shapes[si] : nullptr, t);
    }
}

__global__ void __launch_bounds__(BD_BLOCK) k_integrate_position(DevWorld w) {
    const float dt = w.scal->dt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.n_owned; i += gridDim.x * blockDim.x) {
        integrate_position_one(w, i, dt);
        w.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);    // frClearBodyForces, world.c:481-482
    }
}

__global__ void k_begin_step(DevWorld w) {
    StepCountersDev *c = w.ctr;
    c->n_entries = c->n_cand = c->n_hits = c->n_manifolds = c->n_contacts = c->n_colors = 0u;
    c->overflow &= (32u | 128u);      // a strip's border pack runs BEFORE this kernel: keep what it reported (cleared by k_finish_step)
    c->sort_passes = 0u;
    c->cell_min_x = c->cell_min_y = 0x7fffffff;
    c->cell_max_x = c->cell_max_y = (int) 0x80000000;
    c->grid_w = c->grid_h = 1u;
    c->key_bits = 0u;
    c->n_large = 0u;
    c->uncolored = 0u;
}

__global__ void k_finish_step(DevWorld w, StepCountersDev *snapshot) {
    *w.prev_count = w.ctr->n_manifolds;
    w.prev_count[1] = w.ctr->n_colors;              // the colouring squeezes its top colours next step (fx_solve.cu)
    *w.cur_flip ^= 1;
    *snapshot = *w.ctr;
    w.ctr->overflow = 0u;
}

static inline int grid_for(long long n, int block, int max_blocks) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int) g;
}

__global__ void k_set_scalars(StepScalars *dst, StepScalars v) { *dst = v; }
void fx_launch_set_scalars(StepScalars *dst, const StepScalars &v, cudaStream_t s) {
    k_set_scalars<<<1, 1, 0, s>>>(dst, v); ++g_fx_launches;
}

void fx_launch_begin_step(const DevWorld &w, const FxLaunchCtx &cx) {
    const size_t planes = (size_t) (w.sort_launch > 0 && w.sort_launch < FX_SORT_MAX_PASSES ? w.sort_launch : FX_SORT_MAX_PASSES);
    size_t bytes = cx.zero_fixed_bytes + planes * cx.digit_status_stride * sizeof(unsigned int);
    if (bytes > cx.zero_bytes) bytes = cx.zero_bytes;
    cudaMemsetAsync(cx.zero_base, 0, bytes, cx.stream);
    k_begin_step<<<1, 1, 0, cx.stream>>>(w); ++g_fx_launches;
}

void fx_launch_integrate_position(const DevWorld &w, const FxLaunchCtx &cx) {
    if (w.n_owned == 0) return;
    k_integrate_position<<<grid_for(w.n_owned, BD_BLOCK, cx.max_blocks), BD_BLOCK, 0, cx.stream>>>(w); ++g_fx_launches;
}

void fx_launch_finish_step(const DevWorld &w, const FxLaunchCtx &cx, StepCountersDev *snapshot) {
    k_finish_step<<<1, 1, 0, cx.stream>>>(w, snapshot); ++g_fx_launches;
}

// ---------------------------------------------------------------------------------------------
// single-item entry points
// ---------------------------------------------------------------------------------------------

// frIntegrateForBodyVelocity / frIntegrateForBodyPosition on one body of a (scratch) world
__global__ void k_single_body(DevWorld w, int i, int op, float dt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (op == 0) {
        float4 v = w.vel[i];
        const float4 mp = w.mprop[i], f = w.force[i];
        if (v.w > 0.0f && dt > 0.0f) {                  // src/rigid_body.c:418-427 (no gravity here)
            float s = v.w * dt;
            v.x = v.x + f.x * s;
            v.y = v.y + f.y * s;
            v.z += (f.z * mp.y) * dt;
            w.vel[i] = v;
        }
    } else {
        integrate_position_one(w, i, dt);
    }
}

// frApplyAccumulatedImpulses (op 0) / frResolveCollision (op 1) on bodies 0 and 1 of a scratch world
__global__ void k_single_pair_solve(DevWorld w, CollisionPod *col, int op, float inv_dt) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float4 va4 = w.vel[0], vb4 = w.vel[1];
    const float inv_ma = va4.w, inv_mb = vb4.w, inv_ia = w.mprop[0].y, inv_ib = w.mprop[1].y;
    if (inv_ma + inv_mb <= 0.0f) {
        if (op == 0) {
            if ((__float_as_int(w.mprop[0].w) & 0xff) == FX_BODY_STATIC) w.vel[0] = make_float4(0.f, 0.f, 0.f, va4.w);
            if ((__float_as_int(w.mprop[1].w) & 0xff) == FX_BODY_STATIC) w.vel[1] = make_float4(0.f, 0.f, 0.f, vb4.w);
        }
        return;
    }
    if (op == 1 && inv_dt <= 0.0f) return;
    const float4 pa4 = w.posrot[0], pb4 = w.posrot[1];
    const V2 n = col->direction, t = vperp_right(n);
    Vel3 va = { va4.x, va4.y, va4.z }, vb = { vb4.x, vb4.y, vb4.z };
    for (int c = 0; c < col->count && c < 2; ++c) {
        ContactPod &ct = col->contacts[c];
        ContactK k;
        contact_arms(k, ct.point, v2(pa4.x, pa4.y), v2(pb4.x, pb4.y));
        if (op == 0) {
            contact_masses(k, n, t, inv_ma, inv_mb, inv_ia, inv_ib);
            ct.normalMass = k.nmass;
            ct.tangentMass = k.tmass;
            contact_warm_start(k, n, t, ct.normalScalar, ct.tangentScalar, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
        } else {
            k.bias = contact_bias(ct.depth, inv_dt);
            k.nmass = ct.normalMass;
            k.tmass = ct.tangentMass;
            contact_resolve(k, n, t, col->friction, col->restitution, ct.normalScalar, ct.tangentScalar, inv_ma, inv_mb,
                            inv_ia, inv_ib, va, vb);
        }
    }
    w.vel[0] = make_float4(va.x, va.y, va.w, inv_ma);
    w.vel[1] = make_float4(vb.x, vb.y, vb.w, inv_mb);
}

__global__ void k_sincos(const float *a, int n, float *s, float *c) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        SinCos r = fx_sincosf(a[i]);
        s[i] = r.s;
        c[i] = r.c;
    }
}

// ---------------------------------------------------------------------------------------------
// world ray casts on the device state (reference src/world.c:291-320 + src/broad_phase.c:130-179)
//
// The reference re-inserts every body into the spatial hash, queries the ray's bounding box and casts the ray
// against every body found, in ascending index order.  A body is found iff its inclusive cell rectangle
// intersects the box's (App. A.2), so the same visiting set and order come out of ONE ordered sweep over the
// bodies: a block per ray, 256 bodies per round, the ray test on the bodies whose rectangle matches, hits
// compacted in index order.  Thousands of rays (the sensor fans of batched RL worlds) go in one launch.
// ---------------------------------------------------------------------------------------------
#define RC_BLOCK 256
struct RayDev {
    V2 origin, direction;
    float max_distance;
};
static_assert(sizeof(RayDev) == 20, "frRay layout");

__global__ void __launch_bounds__(RC_BLOCK) k_raycast_batch(DevWorld w, const RayDev *rays, int nrays, int max_hits, RayHitPod *hits, int *counts) {
    __shared__ unsigned int s_run;
    for (int r = blockIdx.x; r < nrays; r += gridDim.x) {
        const RayDev ray = rays[r];
        const V2 o = ray.origin, dir = vunit(ray.direction);
        const V2 e = vadd(o, vscale(dir, ray.max_distance));
        // the query box of src/world.c:311-314 and its cell rectangle (src/broad_phase.c:138-142)
        const float bx = pin_fminf(o.x, e.x), by = pin_fminf(o.y, e.y), bw = fabsf(e.x - o.x), bh = fabsf(e.y - o.y);
        const int qx0 = cell_coord(bx, w.inv_cell), qy0 = cell_coord(by, w.inv_cell);
        const int qx1 = cell_coord(bx + bw, w.inv_cell), qy1 = cell_coord(by + bh, w.inv_cell);
        if (threadIdx.x == 0) s_run = 0u;
        __syncthreads();
        for (int base = 0; base < w.n_owned; base += RC_BLOCK) {
            const int i = base + (int) threadIdx.x;
            bool found = false;
            RayHitPod hit;
            memset(&hit, 0, sizeof hit);
            if (i < w.n_owned) {
                const float4 a = w.aabb[i];
                const int x0 = cell_coord(a.x, w.inv_cell), y0 = cell_coord(a.y, w.inv_cell);
                const int x1 = cell_coord(a.x + a.z, w.inv_cell), y1 = cell_coord(a.y + a.w, w.inv_cell);
                if (x0 <= qx1 && qx0 <= x1 && y0 <= qy1 && qy0 <= y1) {
                    const int2 sh = w.shape[i];
                    if (sh.x >= 0) {
                        const ShapeDev *s = &w.shapes[sh.x];
                        const float4 p = w.posrot[i];
                        Xf t;
                        t.p = v2(p.x, p.y); t.sn = p.z; t.cs = p.w;
                        found = ray_vs_shape(s->type, s->radius, s->nv, s->verts, t, o, dir, ray.max_distance, hit);
                        hit.body = i;
                    }
                }
            }
            unsigned int total;
            const unsigned int excl = block_exclusive_scan<RC_BLOCK>(found ? 1u : 0u, &total);
            const unsigned int pos = s_run + excl;
            if (found && (int) pos < max_hits) hits[(size_t) r * (size_t) max_hits + pos] = hit;
            __syncthreads();
            if (threadIdx.x == 0) s_run += total;
            __syncthreads();
        }
        if (threadIdx.x == 0) counts[r] = (int) s_run;
        __syncthreads();
    }
}

// point queries: which bodies contain each point (frBodyContainsPoint per body, ascending index), a block per point
__global__ void __launch_bounds__(RC_BLOCK) k_point_query_batch(DevWorld w, const V2 *points, int npoints, int max_hits, int *bodies, int *counts) {
    __shared__ unsigned int s_run;
    for (int r = blockIdx.x; r < npoints; r += gridDim.x) {
        const V2 pt = points[r];
        if (threadIdx.x == 0) s_run = 0u;
        __syncthreads();
        for (int base = 0; base < w.n_owned; base += RC_BLOCK) {
            const int i = base + (int) threadIdx.x;
            bool found = false;
            if (i < w.n_owned) {
                const int2 sh = w.shape[i];
                if (sh.x >= 0) {
                    const ShapeDev *s = &w.shapes[sh.x];
                    const float4 p = w.posrot[i];
                    Xf t;
                    t.p = v2(p.x, p.y); t.sn = p.z; t.cs = p.w;
                    found = shape_contains_point(s->type, s->radius, s->nv, s->verts, t, pt);
                }
            }
            unsigned int total;
            const unsigned int excl = block_exclusive_scan<RC_BLOCK>(found ? 1u : 0u, &total);
            const unsigned int pos = s_run + excl;
            if (found && (int) pos < max_hits) bodies[(size_t) r * (size_t) max_hits + pos] = i;
            __syncthreads();
            if (threadIdx.x == 0) s_run += total;
            __syncthreads();
        }
        if (threadIdx.x == 0) counts[r] = (int) s_run;
        __syncthreads();
    }
}
void fx_launch_point_query_batch(const DevWorld &w, const FxLaunchCtx &cx, const void *points, int npoints, int max_hits, int *bodies, int *counts) {
    if (npoints <= 0) return;
    k_point_query_batch<<<npoints < cx.max_blocks ? npoints : cx.max_blocks, RC_BLOCK, 0, cx.stream>>>(w, (const V2 *) points, npoints, max_hits, bodies, counts);
    ++g_fx_launches;
}

void fx_launch_raycast_batch(const DevWorld &w, const FxLaunchCtx &cx, const void *rays, int nrays, int max_hits, void *hits, int *counts) {
    if (nrays <= 0) return;
    k_raycast_batch<<<nrays < cx.max_blocks ? nrays : cx.max_blocks, RC_BLOCK, 0, cx.stream>>>(w, (const RayDev *) rays, nrays, max_hits, (RayHitPod *) hits, counts);
    ++g_fx_launches;
}

void fx_launch_single_body(const DevWorld &w, cudaStream_t s, int i, int op, float dt) { k_single_body<<<1, 32, 0, s>>>(w, i, op, dt); ++g_fx_launches; }
void fx_launch_single_pair_solve(const DevWorld &w, cudaStream_t s, CollisionPod *col, int op, float inv_dt) {
    k_single_pair_solve<<<1, 32, 0, s>>>(w, col, op, inv_dt); ++g_fx_launches;
}
void fx_launch_sincos(const float *a, int n, float *s, float *c, cudaStream_t st) {
    if (n <= 0) return;
    k_sincos<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(a, n, s, c); ++g_fx_launches;
}
