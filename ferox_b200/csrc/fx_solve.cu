// fx_solve.cu -- contact colouring, warm start and the Projected Gauss-Seidel sweeps (K7-K9).
//
// Replaces reference src/world.c:237-249 (warm start over the cache array, then
// FR_WORLD_ITERATION_COUNT sweeps of frResolveCollision in array order).
//
// COLOURED MODE (production).  The whole solve is ONE cooperative kernel; phases are separated by
// grid-wide barriers instead of kernel launches (10 iterations x ~7 colours + warm start is ~80
// dependent phases; a launch per phase would cost more than the phases themselves):
//   1. greedy colouring of the manifold graph (two manifolds conflict iff they share a body that
//      the solver WRITES: invMass > 0 or invInertia > 0; static/kinematic bodies are read-only,
//      SURVEY.md App. A.13): Luby-style rounds, a manifold takes the lowest colour free on both
//      bodies when it holds the minimum (hashed) ticket on both; tickets go through atomicMin, so
//      the colouring is deterministic.
//   2. counting sort of manifold indices by colour -> `order`.
//   3. per colour: build the colour-major solver records (lever arms, unit perpendiculars, bias,
//      effective masses -- everything that depends on positions only) and apply the warm start.
//   4. iterations x colours: one thread per manifold streams its record, gathers the two body
//      velocities {vx, vy, omega, invMass} (16 B each), updates both contacts, scatters.
//   5. accumulated impulses + effective masses are written back to the cache.
// Per Gauss-Seidel sweep the algorithmic traffic is 44 B + 2*32 B per manifold and 56 B + 8 B per
// contact (SURVEY.md section 8d counts 80 M + 36 Q; we keep r1/r2/u1/u2 precomputed: +24 B/contact).
//
// SERIAL MODE (parity).  One thread visits the manifolds in the reference's cache-array order
// (maintained by k_serial_order, which replays the stb_ds insertion / swap-delete discipline,
// reference src/external/stb_ds.h:1436-1448, 1506-1520) with the same per-contact arithmetic
// (fx_solver.cuh), which makes a step bit-comparable with the CPU reference.
#include <cooperative_groups.h>
#include <cstdlib>
#include <utility>

#include "fx_device.h"
#include "fx_kernels.h"
extern long long g_fx_launches;
#include "fx_scan.cuh"
#include "fx_solver.cuh"

namespace cg = cooperative_groups;

#ifndef SV_BLOCK
#define SV_BLOCK 512   // [measured] 65k bodies: 128 -> 0.359, 256 -> 0.339, 512 -> 0.330 ms per solve; no difference at 1 M
#endif
// per-contact solver records: slot-major [slot][2] for small worlds (the rare second contact of a slot shares the
// sector of the first, so its load is an L1 hit inside a latency-bound phase) or CONTACT-MAJOR [2][slot] for
// large ones (93 % of the manifolds of a pile carry one contact, Q/M = 1.07: side by side, every sweep dragged
// the unused second 16 bytes of each sector through HBM).  [measured] contact-major: solve 2.48 -> 2.25 ms at
// 1 M bodies, 0.334 -> 0.352 ms at 65 k; the host picks per world size (fill_dev_params).
#define SOL_AT(w, s, c) ((size_t) (c) * (size_t) (w).sol_stride + (size_t) (s) * (size_t) (w).sol_mult)

__device__ __forceinline__ unsigned int ticket_of(unsigned int m, unsigned int round) {
    // bijection on 26 bits (odd multipliers and xor-shifts are invertible), so tickets are unique
    unsigned int x = m & 0x3ffffffu;
    x = (x * 0x2545F49u) & 0x3ffffffu;
    x ^= x >> 13;
    x = (x * 0x1E3779Bu) & 0x3ffffffu;
    x ^= x >> 11;
    return ((63u - round) << 26) | x;
}

__device__ __forceinline__ bool body_is_written(const DevWorld &w, int i) {
    return w.vel[i].w > 0.0f || w.mprop[i].y > 0.0f;
}

// Where the sweeps find a body's velocity {vx, vy, omega, invMass}: the world array (or a block's shared-memory copy
// of an index range of it), or -- k_sweep_cluster -- the distributed shared memory of a thread-block cluster.  The
// arithmetic below is written once against this accessor, so every solver kernel produces the same bits.
struct VelArray {
    float4 *p;
    int lo;
    __device__ __forceinline__ float4 load(int i) const { return p[i - lo]; }
    __device__ __forceinline__ void store(int i, float4 v) const { p[i - lo] = v; }
};

// solver record word `count_src`: [1:0] contacts to solve (0: both bodies immovable), [2] / [3] body A / B is STATIC and is
// brought to rest by the warm start of such a manifold (src/rigid_body.c:339-347), [31:4] source manifold index
#define CS_COUNT(cs) ((cs) & 3)
#define CS_ZERO_A 4
#define CS_ZERO_B 8
#define CS_SRC(cs) ((unsigned int) (cs) >> 4)

// what the warm start of one manifold needs (all of it is in its solver record)
struct WarmRec {
    int2 b;
    int cs;
    float4 nt, mat;      // {n, t}, {friction, restitution, invInertiaA, invInertiaB}
    float4 r[2];         // lever arms of the (up to) two contacts
    float2 lam[2];       // cached impulses
};

// build the solver record of manifold m at slot s: everything that depends on positions only (src/rigid_body.c:350-385)
template <bool LEAN>
__device__ __forceinline__ WarmRec build_record(const DevWorld &w, const ManifoldSet &man, unsigned int m, unsigned int s, float inv_ma,
                                                float inv_mb, float inv_dt) {
    WarmRec q;
    const int2 b = man.body[m];
    const float4 hdr = man.hdr[m];
    const int count = man.meta[m].x;
    const V2 n = v2(hdr.x, hdr.y);
    const V2 t = vperp_right(n);
    const float4 mpa = w.mprop[b.x], mpb = w.mprop[b.y];
    const float inv_ia = mpa.y, inv_ib = mpb.y;
    const float4 pa4 = w.posrot[b.x], pb4 = w.posrot[b.y];
    const bool active = (inv_ma + inv_mb > 0.0f);           // src/rigid_body.c:338
    int cs = (active ? count : 0) | (int) (m << 4);
    if (!active) {
        if ((__float_as_int(mpa.w) & 0xff) == FX_BODY_STATIC) cs |= CS_ZERO_A;
        if ((__float_as_int(mpb.w) & 0xff) == FX_BODY_STATIC) cs |= CS_ZERO_B;
    }
    q.b = b;
    q.cs = cs;
    q.nt = make_float4(n.x, n.y, t.x, t.y);
    q.mat = make_float4(hdr.z, hdr.w, inv_ia, inv_ib);
    q.r[0] = q.r[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    q.lam[0] = q.lam[1] = make_float2(0.f, 0.f);
    w.sol.body[s] = b;
    if (LEAN) w.sol.n[s] = make_float2(n.x, n.y);
    else w.sol.nt[s] = q.nt;
    w.sol.mat[s] = q.mat;
    w.sol.count_src[s] = cs;
    if (!active) return q;
#pragma unroll
    for (int c = 0; c < 2; ++c) {                      // (unrolled: q.r / q.lam stay in registers)
        if (c >= count) break;
        const float4 g = man.geo[2 * m + c];
        const float4 im = man.imp[2 * m + c];
        ContactK k;
        contact_arms(k, v2(g.x, g.y), v2(pa4.x, pa4.y), v2(pb4.x, pb4.y));
        k.bias = contact_bias(g.z, inv_dt);
        contact_masses(k, n, t, inv_ma, inv_mb, inv_ia, inv_ib);
        q.r[c] = make_float4(k.r1.x, k.r1.y, k.r2.x, k.r2.y);
        q.lam[c] = make_float2(im.y, im.w);
        w.sol.r[SOL_AT(w, s, c)] = q.r[c];
        if (!LEAN) w.sol.u[SOL_AT(w, s, c)] = make_float4(k.u1.x, k.u1.y, k.u2.x, k.u2.y);
        w.sol.k[SOL_AT(w, s, c)] = make_float4(k.bias, k.nmass, k.tmass, 0.0f);
        w.sol.lam[SOL_AT(w, s, c)] = q.lam[c];
    }
    return q;
}

// LEAN records (w.sol_lean: large worlds swept by k_solve_colored, whose sweeps stream the records from DRAM): a
// record keeps only what cannot be recomputed bit for bit from the rest of it -- the normal (the tangent is
// frVector2RightNormal of it, as in build_record) and the lever arms (the unit perpendiculars `u` of the relative
// velocity's angular term are frVector2LeftNormal of them, contact_arms) -- and a sweep reads 76 bytes per manifold
// instead of 100 (40 instead of 56 for a second contact).  [measured] 1 M circles: solve 1.83 -> 1.72 ms.  Small and
// batched worlds keep the full records: their phases are latency-bound and the square roots sit on the critical path
// ([measured] with lean records: config 2 solve 0.294 -> 0.305 ms, config 4 0.96 -> 1.13 ms).
template <bool LEAN>
__device__ __forceinline__ float4 load_nt(const DevWorld &w, unsigned int s) {
    if (!LEAN) return w.sol.nt[s];
    const float2 n = w.sol.n[s];
    const V2 t = vperp_right(v2(n.x, n.y));
    return make_float4(n.x, n.y, t.x, t.y);
}

template <bool LEAN>
__device__ __forceinline__ WarmRec load_warm_record(const DevWorld &w, unsigned int s, int2 b, int cs) {
    WarmRec q;
    q.b = b;
    q.cs = cs;
    q.nt = load_nt<LEAN>(w, s);
    q.mat = w.sol.mat[s];
    const int count = CS_COUNT(cs);
    q.r[0] = q.r[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    q.lam[0] = q.lam[1] = make_float2(0.f, 0.f);
    if (count > 0) { q.r[0] = w.sol.r[SOL_AT(w, s, 0)]; q.lam[0] = w.sol.lam[SOL_AT(w, s, 0)]; }
    if (count > 1) { q.r[1] = w.sol.r[SOL_AT(w, s, 1)]; q.lam[1] = w.sol.lam[SOL_AT(w, s, 1)]; }
    return q;
}

// the warm start of one manifold (src/rigid_body.c:333-411) from its record
template <class Vel>
__device__ __forceinline__ void warm_start_apply(const WarmRec &q, const Vel &vel) {
    const int count = CS_COUNT(q.cs);
    if (count == 0) {
        // static bodies are brought to rest (src/rigid_body.c:339-347); nothing else happens
        if (q.cs & CS_ZERO_A) { const float4 v = vel.load(q.b.x); vel.store(q.b.x, make_float4(0.f, 0.f, 0.f, v.w)); }
        if (q.cs & CS_ZERO_B) { const float4 v = vel.load(q.b.y); vel.store(q.b.y, make_float4(0.f, 0.f, 0.f, v.w)); }
        return;
    }
    const float4 va4 = vel.load(q.b.x), vb4 = vel.load(q.b.y);
    const float inv_ma = va4.w, inv_mb = vb4.w, inv_ia = q.mat.z, inv_ib = q.mat.w;
    const V2 n = v2(q.nt.x, q.nt.y), t = v2(q.nt.z, q.nt.w);
    Vel3 va = { va4.x, va4.y, va4.z }, vb = { vb4.x, vb4.y, vb4.z };
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c >= count) break;
        ContactK k;
        k.r1 = v2(q.r[c].x, q.r[c].y);
        k.r2 = v2(q.r[c].z, q.r[c].w);
        contact_warm_start(k, n, t, q.lam[c].x, q.lam[c].y, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
    }
    // a body that cannot move is left alone: its value would not change, and thousands of manifolds of one
    // colour share a floor ([measured] 1 M circles on 8 floors: solve 6.5 ms with these stores)
    if (inv_ma > 0.0f || inv_ia > 0.0f) vel.store(q.b.x, make_float4(va.x, va.y, va.w, inv_ma));
    if (inv_mb > 0.0f || inv_ib > 0.0f) vel.store(q.b.y, make_float4(vb.x, vb.y, vb.w, inv_mb));
}

template <bool LEAN, class Vel>
__device__ __forceinline__ void prepare_and_warm_start(const DevWorld &w, const ManifoldSet &man, unsigned int m, unsigned int s,
                                                       const Vel &vel, float inv_dt) {
    const int2 b = man.body[m];
    warm_start_apply(build_record<LEAN>(w, man, m, s, vel.load(b.x).w, vel.load(b.y).w, inv_dt), vel);
}

// what a sweep reads of a slot through the slot index alone ("level 1"); it does not depend on any velocity
struct SweepRec {
    int cs;
    int2 b;
    float4 nt, mat, r0, u0, k0;      // (u0: full records only)
    float2 l0;
};
template <bool LEAN>
__device__ __forceinline__ SweepRec sweep_load_rest(const DevWorld &w, unsigned int s, int2 b, int cs) {
    SweepRec q;
    q.cs = cs;
    q.b = b;
    q.nt = load_nt<LEAN>(w, s);
    q.mat = w.sol.mat[s];
    q.r0 = w.sol.r[SOL_AT(w, s, 0)]; q.u0 = LEAN ? q.r0 : w.sol.u[SOL_AT(w, s, 0)]; q.k0 = w.sol.k[SOL_AT(w, s, 0)];
    q.l0 = w.sol.lam[SOL_AT(w, s, 0)];
    return q;
}
template <bool LEAN>
__device__ __forceinline__ SweepRec sweep_load(const DevWorld &w, unsigned int s) {
    return sweep_load_rest<LEAN>(w, s, w.sol.body[s], w.sol.count_src[s]);
}
// one Gauss-Seidel update of the manifold at slot s; va4 / vb4: the velocities of its bodies as just loaded
template <bool LEAN, class Vel>
__device__ __forceinline__ void sweep_apply(const DevWorld &w, unsigned int s, const SweepRec &q, float4 va4, float4 vb4, const Vel &vel) {
    const int count = CS_COUNT(q.cs);
    // level 2: the velocities (addressed through b) and, for the few two-contact manifolds, the second record
    const int2 b = q.b;
    float4 r1 = q.r0, u1 = q.u0, k1 = q.k0;
    float2 l0 = q.l0, l1 = q.l0;
    if (count > 1) {
        r1 = w.sol.r[SOL_AT(w, s, 1)];
        if (!LEAN) u1 = w.sol.u[SOL_AT(w, s, 1)];
        k1 = w.sol.k[SOL_AT(w, s, 1)];
        l1 = w.sol.lam[SOL_AT(w, s, 1)];
    }
    Vel3 va = { va4.x, va4.y, va4.z }, vb = { vb4.x, vb4.y, vb4.z };
    const V2 n = v2(q.nt.x, q.nt.y), t = v2(q.nt.z, q.nt.w);
    ContactK k;
    k.r1 = v2(q.r0.x, q.r0.y); k.r2 = v2(q.r0.z, q.r0.w);
    if (LEAN) { k.u1 = vperp_left(k.r1); k.u2 = vperp_left(k.r2); }
    else { k.u1 = v2(q.u0.x, q.u0.y); k.u2 = v2(q.u0.z, q.u0.w); }
    k.bias = q.k0.x; k.nmass = q.k0.y; k.tmass = q.k0.z;
    contact_resolve(k, n, t, q.mat.x, q.mat.y, l0.x, l0.y, va4.w, vb4.w, q.mat.z, q.mat.w, va, vb);
    if (count > 1) {
        k.r1 = v2(r1.x, r1.y); k.r2 = v2(r1.z, r1.w);
        if (LEAN) { k.u1 = vperp_left(k.r1); k.u2 = vperp_left(k.r2); }
        else { k.u1 = v2(u1.x, u1.y); k.u2 = v2(u1.z, u1.w); }
        k.bias = k1.x; k.nmass = k1.y; k.tmass = k1.z;
        contact_resolve(k, n, t, q.mat.x, q.mat.y, l1.x, l1.y, va4.w, vb4.w, q.mat.z, q.mat.w, va, vb);
    }
    if (va4.w > 0.0f || q.mat.z > 0.0f) vel.store(b.x, make_float4(va.x, va.y, va.w, va4.w));     // movable bodies only
    if (vb4.w > 0.0f || q.mat.w > 0.0f) vel.store(b.y, make_float4(vb.x, vb.y, vb.w, vb4.w));
    w.sol.lam[SOL_AT(w, s, 0)] = l0;
    if (count > 1) w.sol.lam[SOL_AT(w, s, 1)] = l1;
}
template <bool LEAN, class Vel>
__device__ __forceinline__ void sweep_one(const DevWorld &w, unsigned int s, const Vel &vel) {
    const SweepRec q = sweep_load<LEAN>(w, s);
    if (CS_COUNT(q.cs) == 0) return;
    sweep_apply<LEAN>(w, s, q, vel.load(q.b.x), vel.load(q.b.y), vel);
}

// `phases`: bit 0 = colouring + ordering + solver records + warm start, bit 1 = `iterations` sweeps,
// bit 2 = write-back.  A plain step launches all three at once; a strip-decomposed world launches
// them separately so that ghost velocities can be refreshed between sweeps.
#define SV_SETUP 1
#define SV_SWEEP 2
#define SV_WRITEBACK 4
#define SV_RECORDS_ONLY 8     // with SV_SETUP: colouring, ordering and the solver records, but no warm start (k_sweep_cluster follows)
// Two builds of the same kernel: MINB = 1 may use 128 registers (96 used, no spills; one block per SM: the
// latency-bound small-world case), MINB = 2 is held to 64 registers so that two blocks fit an SM (large worlds
// want the threads).  [measured] solve at 65k bodies 0.323 (MINB 1) vs 0.330 ms; at 1 M bodies 2.46 vs 2.27 ms.
// LEAN: the records hold the normal and the lever arms only (load_nt above); the host picks the build from w.sol_lean.
template <int MINB, bool LEAN>
__global__ void __launch_bounds__(SV_BLOCK, MINB) k_solve_colored(DevWorld w, int iterations, int phases) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int s_hist[FX_MAX_COLORS];
    __shared__ unsigned int s_off[FX_MAX_COLORS + 1];
    __shared__ unsigned int s_ncolors;
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int M = w.ctr->n_manifolds;
    const float inv_dt = w.scal->inv_dt;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // per-colour loops: consecutive WARPS go to different blocks, so that a colour with fewer manifolds
    // than the grid has threads keeps a few warps busy on every SM instead of filling the first blocks
    // only; a warp still reads 32 consecutive records ([measured] solve 0.357 -> 0.338 ms at 65k bodies)
    const unsigned int vtid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u);
    unsigned int *color_cnt = w.color_scratch, *color_fill = w.color_scratch + FX_MAX_COLORS,
                 *color_rem = w.color_scratch + 2 * FX_MAX_COLORS;
    // opt-in stage trace (FEROX_SOLVE_TRACE=1, frxGetSolveTrace): one thread stamps the global timer behind the barriers
#define SV_TRACE(slot) do { if (w.solve_trace != nullptr && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); w.solve_trace[slot] = t_; } } while (0)
    SV_TRACE(0);

    if (phases & SV_SETUP) {
    // ---- phase 0: per-body colouring state.  Manifolds that persist from the previous step arrive
    // with the colour they had (carried through the contact cache); those colours are re-validated
    // here (a body that became dynamic may now see one colour twice) and only new manifolds go
    // through the greedy rounds below.
    unsigned int *const dup_flag = w.color_scratch + 3 * FX_MAX_COLORS;
    for (unsigned int i = tid; i < (unsigned int) w.n; i += nth) {
        w.body_mask[i] = 0ull;
        w.body_dup[i] = 0ull;
        w.winner[i] = 0xffffffffu;
        w.winner[(size_t) w.n + i] = 0xffffffffu;
    }
    for (unsigned int i = tid; i < 4 * FX_MAX_COLORS; i += nth) w.color_scratch[i] = 0u;
    grid.sync();
    SV_TRACE(1);
    // Colours are inherited from step to step, and contacts come and go: without care the palette creeps up (14
    // colours for a 1 M-circle pile whose bodies touch at most 7 others).  Every step the manifolds holding the
    // top two colours of the previous solve go back into the rounds and take the lowest colour that is free
    // now; once the palette is tight that is a small fraction of the manifolds.
    const unsigned int prev_colors = w.prev_count[1];
    const int squeeze_from = (w.color_squeeze > 0 && prev_colors > 4u) ? (int) prev_colors - w.color_squeeze : 0x7fffffff;
    for (unsigned int m = tid; m < M; m += nth) {
        int2 b = man.body[m];
        int2 cb = make_int2(body_is_written(w, b.x) ? b.x : -1, body_is_written(w, b.y) ? b.y : -1);
        w.cbody[m] = cb;
        int col = man.meta[m].y;
        if (col >= squeeze_from) {
            col = -1;
            int2 meta = man.meta[m];
            meta.y = -1;
            man.meta[m] = meta;
        }
        if (col >= 0) {
            const unsigned long long bit = 1ull << col;
            bool clash = false;
            if (cb.x >= 0 && (atomicOr(&w.body_mask[cb.x], bit) & bit)) { atomicOr(&w.body_dup[cb.x], bit); clash = true; }
            if (cb.y >= 0 && (atomicOr(&w.body_mask[cb.y], bit) & bit)) { atomicOr(&w.body_dup[cb.y], bit); clash = true; }
            if (clash) *dup_flag = 1u;
        }
    }
    grid.sync();
    SV_TRACE(2);
    if (*reinterpret_cast<volatile unsigned int *>(dup_flag)) {     // rare (a body changed type or flags): grid-uniform
        for (unsigned int m = tid; m < M; m += nth) {
            int2 meta = man.meta[m];
            if (meta.y < 0) continue;
            const unsigned long long bit = 1ull << meta.y;
            const int2 cb = w.cbody[m];
            if ((cb.x >= 0 && (w.body_dup[cb.x] & bit)) || (cb.y >= 0 && (w.body_dup[cb.y] & bit))) {
                meta.y = -1;                              // every manifold of a clash is recoloured
                man.meta[m] = meta;
            }
        }
        grid.sync();
    }

    // ---- phase 1: colouring rounds (Luby style: a manifold takes the lowest free colour when its ticket is the
    // smallest on both of its bodies).  One grid barrier per round: the claims of round r+1 go to the other half of
    // `winner` while stragglers may still read the claims of round r; tickets shrink from round to round, so a
    // half needs no reset when it is reused two rounds later.
    for (unsigned int m = tid; m < M; m += nth) {
        if (man.meta[m].y >= 0) continue;
        const int2 cb = w.cbody[m];
        const unsigned int tk = ticket_of(m, 0u);
        if (cb.x >= 0) atomicMin(&w.winner[cb.x], tk);
        if (cb.y >= 0) atomicMin(&w.winner[cb.y], tk);
    }
    for (unsigned int round = 0; round <= 64u; ++round) {
        grid.sync();
        if (round > 0u && (color_rem[round - 1u] == 0u || round == 64u)) { if (tid == 0) w.ctr->uncolored = round; SV_TRACE(3); break; }
        const unsigned int *win = w.winner + (size_t) (round & 1u) * (size_t) w.n;
        unsigned int *next = w.winner + (size_t) ((round + 1u) & 1u) * (size_t) w.n;
        unsigned int left = 0u;
        for (unsigned int m = tid; m < M; m += nth) {
            int2 meta = man.meta[m];
            if (meta.y >= 0) continue;
            const int2 cb = w.cbody[m];
            const unsigned int tk = ticket_of(m, round);
            const bool won = (cb.x < 0 || win[cb.x] == tk) && (cb.y < 0 || win[cb.y] == tk);
            if (!won) {
                ++left;
                const unsigned int tn = ticket_of(m, round + 1u);        // claim for the next round
                if (cb.x >= 0) atomicMin(&next[cb.x], tn);
                if (cb.y >= 0) atomicMin(&next[cb.y], tn);
                continue;
            }
            unsigned long long used = (cb.x >= 0 ? w.body_mask[cb.x] : 0ull) | (cb.y >= 0 ? w.body_mask[cb.y] : 0ull);
            int col = __ffsll((long long) ~used) - 1;
            if (col < 0) { col = FX_MAX_COLORS - 1; atomicOr(&w.ctr->overflow, 8u); }
            meta.y = col;
            man.meta[m] = meta;
            if (cb.x >= 0) w.body_mask[cb.x] |= (1ull << col);
            if (cb.y >= 0) w.body_mask[cb.y] |= (1ull << col);
        }
        if (left) atomicAdd(&color_rem[round], left);
    }

    // ---- phase 2: counting sort of manifold indices by colour
    if (threadIdx.x < FX_MAX_COLORS) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    for (unsigned int m = tid; m < M; m += nth) {
        int col = man.meta[m].y;
        if (col < 0) { col = FX_MAX_COLORS - 1; atomicOr(&w.ctr->overflow, 8u); }   // > 64 rounds
        atomicAdd(&s_hist[col], 1u);
    }
    __syncthreads();
    if (threadIdx.x < FX_MAX_COLORS && s_hist[threadIdx.x]) atomicAdd(&color_cnt[threadIdx.x], s_hist[threadIdx.x]);
    grid.sync();
    SV_TRACE(4);
    if (threadIdx.x == 0) {
        unsigned int run = 0u, nc = 0u;
        for (int c = 0; c < FX_MAX_COLORS; ++c) {
            s_off[c] = run;
            unsigned int k = color_cnt[c];
            run += k;
            if (k) nc = (unsigned int) c + 1u;
        }
        s_off[FX_MAX_COLORS] = run;
        s_ncolors = nc;
        if (blockIdx.x == 0) {
            for (int c = 0; c <= FX_MAX_COLORS; ++c) w.color_off[c] = s_off[c];
            w.ctr->n_colors = nc;
        }
    }
    __syncthreads();
    for (unsigned int base = blockIdx.x * blockDim.x; base < M; base += nth) {
        unsigned int m = base + threadIdx.x;
        int col = -1;
        if (m < M) { col = man.meta[m].y; if (col < 0) col = FX_MAX_COLORS - 1; }
        // warp-aggregated cursor bump: one atomic per distinct colour per warp
        unsigned int peers = __match_any_sync(0xffffffffu, col);
        unsigned int before = __popc(peers & ((1u << lane_id()) - 1u));
        unsigned int pos = 0u;
        if (col >= 0 && before == 0u) pos = atomicAdd(&color_fill[col], (unsigned int) __popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, __ffs(peers) - 1);
        // The solver record of manifold m is built HERE, in manifold order, and written to its slot: everything in it
        // depends on positions only.  The manifold arrays are read densely (whole sectors, every byte used); built per
        // colour through order[] instead, a warp read every tenth manifold of a stretch and half of every sector it
        // touched was another colour's ([measured] 1 M circles: ordering 49 us + ten record/warm-start phases 294 us).
        if (col >= 0) {
            const unsigned int slot = s_off[col] + pos + before;
            w.order[slot] = m;
            const int2 b = man.body[m];
            build_record<LEAN>(w, man, m, slot, w.vel[b.x].w, w.vel[b.y].w, inv_dt);
        }
    }
    grid.sync();
    SV_TRACE(5);

    const VelArray gvel = { w.vel, 0 };
    if (phases & SV_RECORDS_ONLY) return;          // the cluster kernel warm-starts from the records
    // ---- phase 3: warm start from the records, one colour at a time
    for (unsigned int c = 0; c < s_ncolors; ++c) {
        if (s_off[c] == s_off[c + 1]) continue;      // an emptied colour: nothing to do, no barrier (uniform)
        for (unsigned int s = s_off[c] + vtid; s < s_off[c + 1]; s += nth)
            warm_start_apply(load_warm_record<LEAN>(w, s, w.sol.body[s], w.sol.count_src[s]), gvel);
        grid.sync();
    }
    SV_TRACE(6);
    } else {
        // a later launch of the same step: the colour layout is in global memory
        if (threadIdx.x <= FX_MAX_COLORS) s_off[threadIdx.x] = w.color_off[threadIdx.x];
        if (threadIdx.x == 0) s_ncolors = w.ctr->n_colors;
        __syncthreads();
    }
    const unsigned int ncolors = s_ncolors;
    const VelArray gvel2 = { w.vel, 0 };
    // ---- phase 4: Gauss-Seidel sweeps
    // ([measured] fetching the next colour's records BEFORE the barrier -- software pipelining across grid.sync() --
    // was bit-identical and slower: solve 0.313 -> 0.432 ms at 65k bodies; the barrier's fence waits for the loads)
    // ([measured] sweeping the sparse trailing colours -- config 2: 25371, 23942, 22512, 20101, 16451, 10719, 1263, 33 manifolds
    // per colour -- by block 0 alone behind ONE grid barrier was bit-identical and slower, solve 0.296 -> 0.343 ms: three
    // block-serial rounds of dependent L2 round trips cost more than the barrier they save)
    // ([measured] asking L2 for the NEXT phase's records -- prefetch.global.L2 on the very slots a thread sweeps next, issued at
    // the start of a phase so that DRAM keeps streaming under the gather, the arithmetic and the barrier -- at 1 M bodies,
    // where a colour's records are ~40 MB and come from DRAM in every sweep: bit-identical, solve 1.837 -> 2.003 ms.  Nothing
    // is saturated there (DRAM 31 %, L2 34 %); two colours' records do not stay in L2 next to the velocities)
    // ([measured] sweeping TWO manifolds per thread side by side in the lean builds -- cp.async lands both records in the
    // thread's own shared-memory slots, both velocity pairs are gathered at once, then the two updates run; bit-identical,
    // tests/test_gpu_large.py -- at 1 M circles, where a thread has one to three manifolds per phase: solve 1.757 -> 1.768 ms.
    // A launch list of the strip step (profiles/r02_launches_config5_n1.csv), which launches the stages separately, shows
    // where a 1 M-body solve goes: set-up + colouring + records + warm start 0.49 ms, ten sweeps 0.107 ms each, write-back 0.065)
    // ([measured] fetching a thread's NEXT record between the halves of a split grid barrier -- grid.barrier_arrive(), loads,
    // grid.barrier_wait() -- was bit-identical and no faster either, solve 0.297 -> 0.309 ms at 65k bodies: the phase is
    // bound by the barrier and the velocity gather that can only start behind it, not by the record fetch)
    if (phases & SV_SWEEP)
    for (int it = 0; it < iterations; ++it)
        for (unsigned int c = 0; c < ncolors; ++c) {
            if (s_off[c] == s_off[c + 1]) continue;
            for (unsigned int s = s_off[c] + vtid; s < s_off[c + 1]; s += nth) sweep_one<LEAN>(w, s, gvel2);
            grid.sync();
            if (c + 1 == ncolors && it < 23) SV_TRACE(7 + it);
        }
    // ---- phase 5: write the impulses and effective masses back to the cache
    if (phases & SV_WRITEBACK)
    for (unsigned int s = tid; s < M; s += nth) {
        const int cs = w.sol.count_src[s];
        const unsigned int m = CS_SRC(cs);
        const int count = CS_COUNT(cs);
        for (int c = 0; c < count; ++c) {
            float4 kk = w.sol.k[SOL_AT(w, s, c)];
            float2 lam = w.sol.lam[SOL_AT(w, s, c)];
            man.imp[2 * m + c] = make_float4(kk.y, lam.x, kk.z, lam.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Block-local solve: one CTA per independent group of manifolds (a sub-world of a batch, or a whole
// small world).  The group's body velocities are staged in shared memory and the ~100 dependent
// phases of a solve are separated by __syncthreads() instead of grid-wide barriers: a phase costs
// an L2 round trip for its constraint record instead of ~3 us.  Same colouring rules, same
// per-contact arithmetic, same phase order as k_solve_colored.  Used when every group's bodies sit
// in one index range of at most FX_LOCAL_MAX_BODIES (the host knows the ranges).
// ---------------------------------------------------------------------------------------------
__global__ void k_group_count(DevWorld w) {
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int M = w.ctr->n_manifolds;
    for (unsigned int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
        atomicAdd(&w.gcnt[w.group[man.body[m].x]], 1u);
}
__global__ void __launch_bounds__(1024) k_group_scan(DevWorld w) {
    __shared__ unsigned int s_carry;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (int base = 0; base < w.n_groups; base += 1024) {
        const int g = base + (int) threadIdx.x;
        unsigned int v = g < w.n_groups ? w.gcnt[g] : 0u, tot;
        unsigned int ex = block_exclusive_scan<1024>(v, &tot);
        if (g < w.n_groups) { w.goff[g] = s_carry + ex; w.gfill[g] = 0u; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}
__global__ void k_group_scatter(DevWorld w) {
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int M = w.ctr->n_manifolds;
    for (unsigned int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const int g = w.group[man.body[m].x];
        w.gorder[w.goff[g] + atomicAdd(&w.gfill[g], 1u)] = m;
    }
}

__global__ void __launch_bounds__(1024, 1) k_solve_local(DevWorld w, int iterations, int max_bodies) {
    extern __shared__ __align__(16) unsigned char lsm[];
    float4 *svel = reinterpret_cast<float4 *>(lsm);
    unsigned long long *smask = reinterpret_cast<unsigned long long *>(lsm + (size_t) max_bodies * 16);
    unsigned long long *sdup = smask + max_bodies;
    unsigned int *swin = reinterpret_cast<unsigned int *>(sdup + max_bodies);
    __shared__ unsigned int s_hist[FX_MAX_COLORS], s_off[FX_MAX_COLORS + 1], s_fill[FX_MAX_COLORS];
    __shared__ unsigned int s_left, s_ncolors;
    const ManifoldSet man = w.man[*w.cur_flip];
    const float inv_dt = w.scal->inv_dt;
    const unsigned int tid = threadIdx.x, nth = blockDim.x;
    const bool batch = w.group != nullptr;
    const int ngroups = batch ? w.n_groups : 1;
    for (int g = (int) blockIdx.x; g < ngroups; g += (int) gridDim.x) {
        const int lo = batch ? w.grange[g].x : 0, nb = batch ? w.grange[g].y : w.n;
        const unsigned int mbase = batch ? w.goff[g] : 0u, mc = batch ? w.gcnt[g] : w.ctr->n_manifolds;
        const unsigned int *list = batch ? (w.gorder + mbase) : nullptr;
        __syncthreads();
        for (unsigned int i = tid; i < (unsigned int) nb; i += nth) {
            svel[i] = w.vel[lo + i];
            smask[i] = 0ull;
            sdup[i] = 0ull;
            swin[i] = 0xffffffffu;
        }
        if (tid < FX_MAX_COLORS) { s_hist[tid] = 0u; s_fill[tid] = 0u; }
        __syncthreads();
        // ---- colours inherited through the cache: mark, detect clashes
        for (unsigned int k = tid; k < mc; k += nth) {
            const unsigned int m = list ? list[k] : k;
            const int2 b = man.body[m];
            int2 cb = make_int2((svel[b.x - lo].w > 0.0f || w.mprop[b.x].y > 0.0f) ? b.x - lo : -1,
                                (svel[b.y - lo].w > 0.0f || w.mprop[b.y].y > 0.0f) ? b.y - lo : -1);
            w.cbody[m] = cb;
            const int col = man.meta[m].y;
            if (col >= 0) {
                const unsigned long long bit = 1ull << col;
                if (cb.x >= 0 && (atomicOr(&smask[cb.x], bit) & bit)) atomicOr(&sdup[cb.x], bit);
                if (cb.y >= 0 && (atomicOr(&smask[cb.y], bit) & bit)) atomicOr(&sdup[cb.y], bit);
            }
        }
        __syncthreads();
        for (unsigned int k = tid; k < mc; k += nth) {
            const unsigned int m = list ? list[k] : k;
            int2 meta = man.meta[m];
            if (meta.y < 0) continue;
            const unsigned long long bit = 1ull << meta.y;
            const int2 cb = w.cbody[m];
            if ((cb.x >= 0 && (sdup[cb.x] & bit)) || (cb.y >= 0 && (sdup[cb.y] & bit))) { meta.y = -1; man.meta[m] = meta; }
        }
        __syncthreads();
        // ---- greedy rounds for the new manifolds
        unsigned int rounds = 0u;
        for (unsigned int round = 0; round < 64u; ++round) {
            if (tid == 0) s_left = 0u;
            for (unsigned int k = tid; k < mc; k += nth) {
                const unsigned int m = list ? list[k] : k;
                if (man.meta[m].y >= 0) continue;
                const int2 cb = w.cbody[m];
                const unsigned int tk = ticket_of(m, round);
                if (cb.x >= 0) atomicMin(&swin[cb.x], tk);
                if (cb.y >= 0) atomicMin(&swin[cb.y], tk);
            }
            __syncthreads();
            unsigned int left = 0u;
            for (unsigned int k = tid; k < mc; k += nth) {
                const unsigned int m = list ? list[k] : k;
                int2 meta = man.meta[m];
                if (meta.y >= 0) continue;
                const int2 cb = w.cbody[m];
                const unsigned int tk = ticket_of(m, round);
                const bool won = (cb.x < 0 || swin[cb.x] == tk) && (cb.y < 0 || swin[cb.y] == tk);
                if (!won) { ++left; continue; }
                const unsigned long long used = (cb.x >= 0 ? smask[cb.x] : 0ull) | (cb.y >= 0 ? smask[cb.y] : 0ull);
                int col = __ffsll((long long) ~used) - 1;
                if (col < 0) { col = FX_MAX_COLORS - 1; atomicOr(&w.ctr->overflow, 8u); }
                meta.y = col;
                man.meta[m] = meta;
                if (cb.x >= 0) smask[cb.x] |= (1ull << col);
                if (cb.y >= 0) smask[cb.y] |= (1ull << col);
            }
            if (left) atomicAdd(&s_left, left);
            __syncthreads();
            rounds = round + 1u;
            const bool done = (s_left == 0u);
            __syncthreads();
            if (done) break;
        }
        // ---- counting sort of the group's manifolds by colour -> its segment of `order`
        for (unsigned int k = tid; k < mc; k += nth) {
            const unsigned int m = list ? list[k] : k;
            int col = man.meta[m].y;
            if (col < 0) { col = FX_MAX_COLORS - 1; atomicOr(&w.ctr->overflow, 8u); }
            atomicAdd(&s_hist[col], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int run = 0u, nc = 0u;
            for (int c = 0; c < FX_MAX_COLORS; ++c) {
                s_off[c] = run;
                run += s_hist[c];
                if (s_hist[c]) nc = (unsigned int) c + 1u;
            }
            s_off[FX_MAX_COLORS] = run;
            s_ncolors = nc;
            atomicMax(&w.ctr->n_colors, nc);
            atomicMax(&w.ctr->uncolored, rounds);
        }
        __syncthreads();
        for (unsigned int k = tid; k < mc; k += nth) {
            const unsigned int m = list ? list[k] : k;
            int col = man.meta[m].y;
            if (col < 0) col = FX_MAX_COLORS - 1;
            w.order[mbase + s_off[col] + atomicAdd(&s_fill[col], 1u)] = m;
        }
        __syncthreads();
        // ---- solver records + warm start, then the sweeps: block barriers only
        const unsigned int ncolors = s_ncolors;
        const VelArray lvel = { svel, lo };
        for (unsigned int c = 0; c < ncolors; ++c) {
            if (s_off[c] == s_off[c + 1]) continue;
            for (unsigned int s = mbase + s_off[c] + tid; s < mbase + s_off[c + 1]; s += nth) prepare_and_warm_start<false>(w, man, w.order[s], s, lvel, inv_dt);
            __syncthreads();
        }
        for (int it = 0; it < iterations; ++it)
            for (unsigned int c = 0; c < ncolors; ++c) {
                if (s_off[c] == s_off[c + 1]) continue;
                for (unsigned int s = mbase + s_off[c] + tid; s < mbase + s_off[c + 1]; s += nth) sweep_one<false>(w, s, lvel);
                __syncthreads();
            }
        for (unsigned int s = mbase + tid; s < mbase + mc; s += nth) {
            const int cs = w.sol.count_src[s];
            const unsigned int m = CS_SRC(cs);
            const int count = CS_COUNT(cs);
            for (int c = 0; c < count; ++c) {
                const float4 kk = w.sol.k[SOL_AT(w, s, c)];
                const float2 lam = w.sol.lam[SOL_AT(w, s, c)];
                man.imp[2 * m + c] = make_float4(kk.y, lam.x, kk.z, lam.y);
            }
        }
        for (unsigned int i = tid; i < (unsigned int) nb; i += nth) w.vel[lo + i] = svel[i];
    }
}

// ---------------------------------------------------------------------------------------------
// serial (reference-order) mode
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int so_slot(unsigned long long key, unsigned int mask) {
    unsigned long long z = (key ^ (key >> 29)) * 0xBF58476D1CE4E5B9ull;
    return (unsigned int) (z >> 32) & mask;
}
__device__ int so_find(const SerialOrder &o, unsigned long long key) {
    unsigned int mask = o.t_cap - 1u, s = so_slot(key, mask);
    for (;;) {
        unsigned long long k = o.t_key[s];
        if (k == key) return (int) s;
        if (k == 0xffffffffffffffffull) return -1;
        s = (s + 1u) & mask;
    }
}
__device__ void so_put(const SerialOrder &o, unsigned long long key) {   // hmputs: in place or append
    if (so_find(o, key) >= 0) return;
    unsigned int pos = (*o.r_len)++;
    o.r_key[pos] = key;
    unsigned int mask = o.t_cap - 1u, s = so_slot(key, mask);
    while (o.t_key[s] != 0xffffffffffffffffull) s = (s + 1u) & mask;
    o.t_key[s] = key;
    o.t_val[s] = pos;
}
__device__ void so_erase_slot(const SerialOrder &o, int s) {   // backward-shift deletion of an index slot
    unsigned int mask = o.t_cap - 1u, hole = (unsigned int) s, j = (unsigned int) s;
    for (;;) {
        j = (j + 1u) & mask;
        unsigned long long kj = o.t_key[j];
        if (kj == 0xffffffffffffffffull) break;
        unsigned int home = so_slot(kj, mask);
        if (((j - home) & mask) >= ((j - hole) & mask)) {
            o.t_key[hole] = kj;
            o.t_val[hole] = o.t_val[j];
            hole = j;
        }
    }
    o.t_key[hole] = 0xffffffffffffffffull;
}
__device__ __forceinline__ int key_group(const DevWorld &w, unsigned long long key) {
    if (w.group == nullptr) return 0;
    int i = w.idx_of_uid[(unsigned int) (key >> 32)];
    return i >= 0 ? w.group[i] : -1;
}
// hmdel: swap-delete.  In a batch every sub-world must keep the array discipline it would have on
// its own, so the hole is filled with the last entry OF THE SAME SUB-WORLD and the slot that entry
// leaves is closed by a stable shift (other sub-worlds keep their relative order).
__device__ void so_del(const DevWorld &w, const SerialOrder &o, unsigned long long key) {
    int s = so_find(o, key);
    if (s < 0) return;
    const unsigned int old_index = o.t_val[s], len = *o.r_len;
    so_erase_slot(o, s);
    unsigned int final_index = len - 1u;
    if (w.group != nullptr) {
        const int g = key_group(w, key);
        while (final_index > old_index && key_group(w, o.r_key[final_index]) != g) --final_index;
    }
    if (old_index != final_index) {
        unsigned long long moved = o.r_key[final_index];
        o.r_key[old_index] = moved;
        o.t_val[so_find(o, moved)] = old_index;
    }
    for (unsigned int p = final_index; p + 1u < len; ++p) {      // no-op in a plain world
        unsigned long long k = o.r_key[p + 1u];
        o.r_key[p] = k;
        o.t_val[so_find(o, k)] = p;
    }
    *o.r_len = len - 1u;
}

__device__ __forceinline__ int cache_lookup(const DevWorld &w, unsigned long long key) {
    const unsigned int mask = w.cap_hash - 1u;
    unsigned long long z = (key ^ (key >> 31)) * 0x9E3779B97F4A7C15ull;
    unsigned int s = (unsigned int) (z >> 32) & mask;
    const uint2 *keys = w.man[*w.cur_flip].key;      // the table was just rebuilt over this step's array
    for (;;) {
        const unsigned int v = w.h_slot[s];
        if (v == 0xffffffffu) return -1;
        const uint2 k = keys[v];
        if ((((unsigned long long) k.x << 32) | (unsigned long long) k.y) == key) return (int) v;
        s = (s + 1u) & mask;
    }
}

// replay this step's candidate outcomes (lexicographic order) on the reference-order key array,
// then the eviction loop of src/world.c:222-235 and the removal of entries that did not survive
__global__ void k_serial_order(DevWorld w, SerialOrder o) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const unsigned int ncand = w.ctr->n_cand;
    for (unsigned int c = 0; c < ncand; ++c) {
        int2 pr = w.cand[c];
        unsigned long long key = ((unsigned long long) w.uid[pr.x] << 32) | (unsigned long long) w.uid[pr.y];
        if (w.cand_hit[c]) so_put(o, key);
        else so_del(w, o, key);
    }
    // entries that the new dense array no longer holds (timestamp rule, departed bodies): the
    // reference walks j = 0 .. entryCount-1 of the pre-loop length while hmdel shrinks the array, so
    // the element swapped into slot j is skipped (App. A.7); replayed here on the key array
    unsigned int entry_count = *o.r_len;
    for (unsigned int j = 0; j < entry_count; ++j) {
        if (j >= *o.r_len) break;   // (the reference reads stale tail bytes here; not replayed)
        unsigned long long key = o.r_key[j];
        if (cache_lookup(w, key) < 0) so_del(w, o, key);
    }
    // anything still missing (skipped by the quirk above but gone from the dense array) must go,
    // otherwise the solver would dereference a missing entry
    for (unsigned int j = 0; j < *o.r_len;) {
        if (cache_lookup(w, o.r_key[j]) < 0) so_del(w, o, o.r_key[j]);
        else ++j;
    }
}

__global__ void k_serial_reindex(SerialOrder o) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const unsigned int len = *o.r_len, mask = o.t_cap - 1u;
    for (unsigned int p = 0; p < len; ++p) {
        unsigned long long key = o.r_key[p];
        unsigned int s = so_slot(key, mask);
        while (o.t_key[s] != 0xffffffffffffffffull) s = (s + 1u) & mask;
        o.t_key[s] = key;
        o.t_val[s] = p;
    }
}

__global__ void k_solve_serial(DevWorld w, SerialOrder o, int iterations) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int len = *o.r_len;
    const float inv_dt = w.scal->inv_dt;
    for (int it = -1; it < iterations; ++it) {
        for (unsigned int r = 0; r < len; ++r) {
            int m = cache_lookup(w, o.r_key[r]);
            if (m < 0) continue;
            const int2 b = man.body[m];
            const float4 hdr = man.hdr[m];
            const int count = man.meta[m].x;
            const V2 n = v2(hdr.x, hdr.y), t = vperp_right(n);
            float4 va4 = w.vel[b.x], vb4 = w.vel[b.y];
            const float inv_ma = va4.w, inv_mb = vb4.w, inv_ia = w.mprop[b.x].y, inv_ib = w.mprop[b.y].y;
            if (!(inv_ma + inv_mb > 0.0f)) {
                if (it < 0) {
                    const int ta = __float_as_int(w.mprop[b.x].w) & 0xff, tb = __float_as_int(w.mprop[b.y].w) & 0xff;
                    if (ta == FX_BODY_STATIC) w.vel[b.x] = make_float4(0.f, 0.f, 0.f, va4.w);
                    if (tb == FX_BODY_STATIC) w.vel[b.y] = make_float4(0.f, 0.f, 0.f, vb4.w);
                }
                continue;
            }
            const float4 pa4 = w.posrot[b.x], pb4 = w.posrot[b.y];
            Vel3 va = { va4.x, va4.y, va4.z }, vb = { vb4.x, vb4.y, vb4.z };
            for (int c = 0; c < count; ++c) {
                const float4 g = man.geo[2 * m + c];
                float4 im = man.imp[2 * m + c];
                ContactK k;
                contact_arms(k, v2(g.x, g.y), v2(pa4.x, pa4.y), v2(pb4.x, pb4.y));
                if (it < 0) {
                    contact_masses(k, n, t, inv_ma, inv_mb, inv_ia, inv_ib);
                    im.x = k.nmass;
                    im.z = k.tmass;
                    contact_warm_start(k, n, t, im.y, im.w, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
                } else {
                    k.bias = contact_bias(g.z, inv_dt);
                    k.nmass = im.x;
                    k.tmass = im.z;
                    contact_resolve(k, n, t, hdr.z, hdr.w, im.y, im.w, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
                }
                man.imp[2 * m + c] = im;
            }
            w.vel[b.x] = make_float4(va.x, va.y, va.w, inv_ma);
            w.vel[b.y] = make_float4(vb.x, vb.y, vb.w, inv_mb);
        }
    }
    w.ctr->n_colors = 0;
}

// ---------------------------------------------------------------------------------------------
// collision-handler bridge (src/world.c:209-214, 254-259): the live manifolds as an array of the PUBLIC records
// -- {first index, second index, frCollision} = 100 bytes, include/ferox_b200.h frxManifold -- in the order the
// reference visits its cache array, packed on the device so that the host gets ONE contiguous download it can hand
// to the callbacks in place, and takes ONE upload back (whatever the callbacks edited is in it).
// ---------------------------------------------------------------------------------------------
#define MB_BLOCK 128
#define MB_WORDS 25     // sizeof(ManifoldRec) / 4
static_assert(sizeof(ManifoldRec) == 4 * MB_WORDS, "frxManifold layout");

// reference-order mode: dense index of every entry of the reference-order key array
__global__ void __launch_bounds__(MB_BLOCK) k_visit_index(DevWorld w, SerialOrder o, unsigned int *idx) {
    const unsigned int len = *o.r_len;
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < len; r += gridDim.x * blockDim.x) {
        const int m = cache_lookup(w, o.r_key[r]);
        idx[r] = m < 0 ? 0xffffffffu : (unsigned int) m;
    }
}

// `head`: {first, second, count, 0} per record -- all the host's callback loop itself needs to read (16 B instead of two
// cache lines of the 100-byte record, which only a callback that looks at its manifold touches)
__global__ void __launch_bounds__(MB_BLOCK) k_pack_collisions(DevWorld w, const unsigned int *idx, unsigned int n, ManifoldRec *out, int4 *head) {
    __shared__ unsigned int s_rec[MB_BLOCK * MB_WORDS];       // odd record stride: conflict-free per-thread fills
    const ManifoldSet man = w.man[*w.cur_flip];
    for (unsigned int base = blockIdx.x * MB_BLOCK; base < n; base += gridDim.x * MB_BLOCK) {
        const unsigned int r = base + threadIdx.x;
        ManifoldRec rec;
        memset(&rec, 0, sizeof rec);
        rec.first = rec.second = -1;
        const unsigned int m = r < n ? (idx ? idx[r] : r) : 0xffffffffu;
        if (m != 0xffffffffu) {
            const int2 b = man.body[m];
            const float4 hdr = man.hdr[m];
            rec.first = b.x; rec.second = b.y;
            rec.value.count = man.meta[m].x;
            rec.value.direction = v2(hdr.x, hdr.y);
            rec.value.friction = hdr.z; rec.value.restitution = hdr.w;
            for (int k = 0; k < 2; ++k) {
                const float4 g = man.geo[2 * m + k], im = man.imp[2 * m + k];
                ContactPod &ct = rec.value.contacts[k];
                ct.id = man.cid[2 * m + k];
                ct.point = v2(g.x, g.y); ct.depth = g.z; ct.timestamp = g.w;
                ct.normalMass = im.x; ct.normalScalar = im.y; ct.tangentMass = im.z; ct.tangentScalar = im.w;
            }
        }
        if (r < n) head[r] = make_int4(rec.first, rec.second, rec.value.count, 0);
        const unsigned int *src = reinterpret_cast<const unsigned int *>(&rec);
#pragma unroll
        for (int k = 0; k < MB_WORDS; ++k) s_rec[threadIdx.x * MB_WORDS + k] = src[k];
        __syncthreads();
        const unsigned int words = min(MB_BLOCK, n - base) * MB_WORDS;
        unsigned int *dst = reinterpret_cast<unsigned int *>(out + base);
        for (unsigned int t = threadIdx.x; t < words; t += MB_BLOCK) dst[t] = s_rec[t];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MB_BLOCK) k_unpack_collisions(DevWorld w, const unsigned int *idx, unsigned int n, const ManifoldRec *in) {
    __shared__ unsigned int s_rec[MB_BLOCK * MB_WORDS];
    const ManifoldSet man = w.man[*w.cur_flip];
    for (unsigned int base = blockIdx.x * MB_BLOCK; base < n; base += gridDim.x * MB_BLOCK) {
        const unsigned int words = min(MB_BLOCK, n - base) * MB_WORDS;
        const unsigned int *src = reinterpret_cast<const unsigned int *>(in + base);
        for (unsigned int t = threadIdx.x; t < words; t += MB_BLOCK) s_rec[t] = src[t];
        __syncthreads();
        const unsigned int r = base + threadIdx.x;
        const unsigned int m = r < n ? (idx ? idx[r] : r) : 0xffffffffu;
        if (m != 0xffffffffu) {
            ManifoldRec rec;
            unsigned int *dst = reinterpret_cast<unsigned int *>(&rec);
#pragma unroll
            for (int k = 0; k < MB_WORDS; ++k) dst[k] = s_rec[threadIdx.x * MB_WORDS + k];
            int2 meta = man.meta[m];
            meta.x = rec.value.count < 0 ? 0 : (rec.value.count > 2 ? 2 : rec.value.count);
            man.meta[m] = meta;
            man.hdr[m] = make_float4(rec.value.direction.x, rec.value.direction.y, rec.value.friction, rec.value.restitution);
            for (int k = 0; k < 2; ++k) {
                const ContactPod &ct = rec.value.contacts[k];
                man.cid[2 * m + k] = ct.id;
                man.geo[2 * m + k] = make_float4(ct.point.x, ct.point.y, ct.depth, ct.timestamp);
                man.imp[2 * m + k] = make_float4(ct.normalMass, ct.normalScalar, ct.tangentMass, ct.tangentScalar);
            }
        }
        __syncthreads();
    }
}

void fx_launch_visit_index(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o, unsigned int *idx, unsigned int cap) {
    long long g = ((long long) cap + MB_BLOCK - 1) / MB_BLOCK;
    k_visit_index<<<(int) (g < 1 ? 1 : (g > cx.max_blocks ? cx.max_blocks : g)), MB_BLOCK, 0, cx.stream>>>(w, o, idx); ++g_fx_launches;
}
void fx_launch_pack_collisions(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *idx, unsigned int n, ManifoldRec *out, int4 *head) {
    if (n == 0) return;
    long long g = ((long long) n + MB_BLOCK - 1) / MB_BLOCK;
    k_pack_collisions<<<(int) (g > cx.max_blocks ? cx.max_blocks : g), MB_BLOCK, 0, cx.stream>>>(w, idx, n, out, head); ++g_fx_launches;
}
void fx_launch_unpack_collisions(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *idx, unsigned int n, const ManifoldRec *in) {
    if (n == 0) return;
    long long g = ((long long) n + MB_BLOCK - 1) / MB_BLOCK;
    k_unpack_collisions<<<(int) (g > cx.max_blocks ? cx.max_blocks : g), MB_BLOCK, 0, cx.stream>>>(w, idx, n, in); ++g_fx_launches;
}

// ---------------------------------------------------------------------------------------------
// Cluster solve (sm_100a thread-block clusters + distributed shared memory + bulk async copies).
//
// A mid-size world (config 2: 65 k bodies, 124 k manifolds) is bound by its ~100 DEPENDENT colour phases, not by
// bytes: in k_solve_colored every phase pays a grid-wide barrier over 148 SMs (~1.5 us), a record fetch and a
// scattered velocity gather through L2, and a store drain -- 2.9 us per phase, 9 % of the HBM roofline.  Here the
// warm start and the sweeps of the WHOLE world run inside ONE thread-block cluster (16 CTAs = 16 SMs on one GPC):
//   * every body's velocity lives in the cluster's distributed shared memory for the whole solve (body i in CTA
//     i mod 16, slot i / 16): the gathers and scatters of a phase are DSMEM accesses (~215 cycles, no L2 queueing);
//   * phases are separated by the hardware cluster barrier (barrier.cluster.arrive.release / wait.acquire, ~380
//     cycles) instead of a grid barrier;
//   * the body indices of the NEXT phase's manifolds -- the head of the dependent chain -- are prefetched into a
//     3-stage shared-memory ring by bulk async copies (cp.async.bulk + mbarrier complete_tx), which no barrier
//     fence waits for; the rest of a record (position-only constants, 80 B) is fetched with coalesced loads that
//     overlap the DSMEM gather.
// Same colours, same phase order, same per-contact arithmetic (the templates above) as k_solve_colored: the two
// kernels produce identical bits (tests/test_gpu_cluster.py).  The colouring, the ordering and the records are
// built by k_solve_colored(SV_SETUP | SV_RECORDS_ONLY) on the full grid just before.
// ---------------------------------------------------------------------------------------------
#define CL_BLOCK 1024
#define CL_STAGES 3
#define CL_PAD 8
#define CL_STAGE_ELEMS (CL_BLOCK + CL_PAD)
#define CL_FIXED_SMEM (128 + CL_STAGES * CL_STAGE_ELEMS * (sizeof(int2) + sizeof(int)))

__device__ __forceinline__ unsigned int cl_smem_addr(const void *p) { return (unsigned int) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cl_mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cl_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cl_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cl_mbar_try_wait(unsigned long long *bar, unsigned int parity) {
    unsigned int ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(cl_smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// bulk async copy global -> this CTA's shared memory, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void cl_bulk_load(void *dst_smem, const void *src_gmem, unsigned int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cl_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(cl_smem_addr(bar)) : "memory");
}

// velocities in the cluster's distributed shared memory: body i -> CTA (i & mask), slot (i >> shift)
struct ClusterVel {
    float4 *base;                 // this CTA's slots
    unsigned int mask, shift;
    __device__ __forceinline__ float4 *at(int i) const {
        return cg::this_cluster().map_shared_rank(base, (unsigned int) i & mask) + ((unsigned int) i >> shift);
    }
    __device__ __forceinline__ float4 load(int i) const { return *at(i); }
    __device__ __forceinline__ void store(int i, float4 v) const { *at(i) = v; }
};

// CTA `rank` of `cs` takes slots [a, b) of colour c
__device__ __forceinline__ void cl_subrange(const unsigned int *s_off, unsigned int c, unsigned int rank, unsigned int cs, unsigned int &a,
                                            unsigned int &b) {
    const unsigned int first = s_off[c], len = s_off[c + 1] - first, per = (len + cs - 1u) / cs;
    a = first + min(len, rank * per);
    b = first + min(len, (rank + 1u) * per);
}

__global__ void __launch_bounds__(CL_BLOCK, 1) k_sweep_cluster(DevWorld w, int iterations) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int CS = cluster.num_blocks(), rank = cluster.block_rank(), tid = threadIdx.x;
    extern __shared__ __align__(128) unsigned char cl_smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(cl_smem);                  // [CL_STAGES]
    int2 *sbody = reinterpret_cast<int2 *>(cl_smem + 128);                                        // [CL_STAGES][CL_STAGE_ELEMS]
    int *scs = reinterpret_cast<int *>(sbody + CL_STAGES * CL_STAGE_ELEMS);                       // [CL_STAGES][CL_STAGE_ELEMS]
    float4 *svel = reinterpret_cast<float4 *>(scs + CL_STAGES * CL_STAGE_ELEMS);                  // [ceil(n / CS)]
    __shared__ unsigned int s_off[FX_MAX_COLORS + 1];
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int M = w.ctr->n_manifolds, ncolors = w.ctr->n_colors;
    const unsigned int n = (unsigned int) w.n, per = (n + CS - 1u) / CS;
    const ClusterVel vel = { svel, CS - 1u, 31u - (unsigned int) __clz((int) CS) };

    if (tid <= FX_MAX_COLORS) s_off[tid] = w.color_off[tid];
    if (tid == 0)
        for (int k = 0; k < CL_STAGES; ++k) cl_mbar_init(&bars[k], 1u);
    for (unsigned int l = tid; l < per; l += CL_BLOCK) {
        const unsigned int i = l * CS + rank;
        svel[l] = i < n ? w.vel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // ---- producer (thread 0): the body-index / count words of this CTA's next chunks, CL_STAGES - 1 chunks ahead
    int p_it = -1;
    unsigned int p_c = 0u, p_k = 0u, p_issued = 0u;
    auto issue_next = [&]() {
        while (p_it < iterations) {
            if (p_c >= ncolors) { p_c = 0u; ++p_it; continue; }
            unsigned int a, b;
            cl_subrange(s_off, p_c, rank, CS, a, b);
            const unsigned int ak = a + p_k * CL_BLOCK;
            if (ak >= b) { ++p_c; p_k = 0u; continue; }
            const unsigned int bk = min(b, ak + CL_BLOCK), a4 = ak & ~3u, cnt4 = (bk - a4 + 3u) & ~3u;
            const unsigned int stage = p_issued % CL_STAGES;
            unsigned long long *bar = &bars[stage];
            cl_mbar_expect_tx(bar, cnt4 * (unsigned int) (sizeof(int2) + sizeof(int)));
            cl_bulk_load(sbody + stage * CL_STAGE_ELEMS, w.sol.body + a4, cnt4 * (unsigned int) sizeof(int2), bar);
            cl_bulk_load(scs + stage * CL_STAGE_ELEMS, w.sol.count_src + a4, cnt4 * (unsigned int) sizeof(int), bar);
            ++p_issued;
            ++p_k;
            return;
        }
    };
    if (tid == 0)
        for (int k = 0; k < CL_STAGES; ++k) issue_next();
    cluster.sync();                       // every CTA's velocities are staged before anyone gathers

    // ---- phases: warm start (it = -1), then `iterations` sweeps; one cluster barrier per non-empty colour
    unsigned int consumed = 0u;
    for (int it = -1; it < iterations; ++it)
        for (unsigned int c = 0; c < ncolors; ++c) {
            if (s_off[c] == s_off[c + 1]) continue;               // an emptied colour: no work, no barrier (uniform)
            unsigned int a, b;
            cl_subrange(s_off, c, rank, CS, a, b);
            for (unsigned int ak = a; ak < b; ak += CL_BLOCK) {
                const unsigned int bk = min(b, ak + CL_BLOCK), a4 = ak & ~3u;
                const unsigned int stage = consumed % CL_STAGES, parity = (consumed / CL_STAGES) & 1u;
                while (!cl_mbar_try_wait(&bars[stage], parity)) { }
                const unsigned int s = ak + tid;
                const bool valid = s < bk;
                int2 body = make_int2(0, 0);
                int cs = 0;
                if (valid) {
                    body = sbody[stage * CL_STAGE_ELEMS + (s - a4)];
                    cs = scs[stage * CL_STAGE_ELEMS + (s - a4)];
                }
                __syncthreads();                                  // every thread holds its words: the stage may be refilled
                if (tid == 0) issue_next();
                ++consumed;
                if (!valid) continue;
                if (it < 0) {
                    warm_start_apply(load_warm_record<false>(w, s, body, cs), vel);
                } else if (CS_COUNT(cs) != 0) {
                    float4 *pa = vel.at(body.x), *pb = vel.at(body.y);
                    const float4 va4 = *pa, vb4 = *pb;            // DSMEM gathers in flight while the record streams in from L2
                    const SweepRec q = sweep_load_rest<false>(w, s, body, cs);
                    sweep_apply<false>(w, s, q, va4, vb4, vel);
                }
            }
            cluster.sync();
        }

    // ---- velocities back to the world array, impulses and effective masses back to the cache
    for (unsigned int l = tid; l < per; l += CL_BLOCK) {
        const unsigned int i = l * CS + rank;
        if (i < n) w.vel[i] = svel[l];
    }
    for (unsigned int s = rank * CL_BLOCK + tid; s < M; s += CS * CL_BLOCK) {
        const int cs = w.sol.count_src[s];
        const unsigned int m = CS_SRC(cs);
        const int count = CS_COUNT(cs);
        for (int c = 0; c < count; ++c) {
            const float4 kk = w.sol.k[SOL_AT(w, s, c)];
            const float2 lam = w.sol.lam[SOL_AT(w, s, c)];
            man.imp[2 * m + c] = make_float4(kk.y, lam.x, kk.z, lam.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch surface
// ---------------------------------------------------------------------------------------------
// the shared-memory opt-in is a per-DEVICE function attribute: set when a world is created on a device
void fx_init_solver_attributes() {
    cudaFuncSetAttribute(k_solve_local, cudaFuncAttributeMaxDynamicSharedMemorySize, FX_LOCAL_MAX_BODIES * 36);
    cudaFuncSetAttribute(k_sweep_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_sweep_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, FX_CLUSTER_MAX_SMEM);
    (void) cudaGetLastError();
}

// Largest cluster (16, else 8 CTAs) that can hold `bodies` velocities in its distributed shared memory AND is
// launchable on the current device; 0: the world does not fit (k_solve_colored takes it).
int fx_cluster_size_for(int bodies) {
    // opt-in (FEROX_CLUSTER=8|16): [measured, B200] bit-identical to the grid solver but SLOWER at every size it fits --
    // config 2: solve 0.685 ms in a 16-CTA cluster, 1.07 ms in 8 CTAs, 0.296 ms on the grid: sixteen SMs issue the
    // ~430 warp instructions per manifold update 9x slower than 148 do, and DSMEM gathers (~20 B/cycle per SM) are a
    // throughput path, not a latency path (profiles/r02_cluster_v1_*)
    static const int pin = [] { const char *e = std::getenv("FEROX_CLUSTER"); return e ? std::atoi(e) : 0; }();   // 0: off, 8 / 16: pin the size
    if (pin <= 0 || bodies <= 0) return 0;
    for (int cs = 16; cs >= 8; cs >>= 1) {
        if (pin > 0 && cs != pin) continue;
        const size_t smem = CL_FIXED_SMEM + (size_t) ((bodies + cs - 1) / cs) * sizeof(float4);
        if (smem > FX_CLUSTER_MAX_SMEM) continue;
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned int) cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned int) cs); cfg.blockDim = dim3(CL_BLOCK); cfg.dynamicSmemBytes = smem;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, k_sweep_cluster, &cfg) == cudaSuccess && clusters >= 1) return cs;
        (void) cudaGetLastError();
    }
    return 0;
}

void fx_launch_solve_cluster(const DevWorld &w, const FxLaunchCtx &cx, int iterations, long long work_hint, int cluster_size) {
    // colouring, ordering and solver records on the full grid ...
    fx_launch_solve_colored(w, cx, 0, work_hint, SV_SETUP | SV_RECORDS_ONLY);
    // ... warm start, sweeps and write-back inside one cluster
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned int) cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned int) cluster_size);
    cfg.blockDim = dim3(CL_BLOCK);
    cfg.dynamicSmemBytes = CL_FIXED_SMEM + (size_t) ((w.n + cluster_size - 1) / cluster_size) * sizeof(float4);
    cfg.stream = cx.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_sweep_cluster, w, iterations);
    ++g_fx_launches;
}

int fx_query_coop_blocks(int device) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_colored<2, false>, SV_BLOCK, 0);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

static bool solve_wide(long long work_hint) {
    // FEROX_SOLVE_WIDE=0|1 pins the variant (tests compare the two: they must agree bit for bit)
    static const int pin = [] { const char *e = std::getenv("FEROX_SOLVE_WIDE"); return e ? std::atoi(e) : -1; }();
    return pin >= 0 ? pin != 0 : work_hint >= 400000;
}

int fx_solve_colored_blocks(const FxLaunchCtx &cx, long long work_hint) {
    // A phase holds ~1/8 of the manifolds (one colour).  Fewer resident blocks make the ~100 grid
    // barriers of a solve cheaper, so take just enough blocks to give every manifold of a phase its
    // own thread, in whole multiples of the SM count, at least one block per SM.
    long long want = (work_hint / 6 + SV_BLOCK - 1) / SV_BLOCK;
    int per_sm = solve_wide(work_hint) ? cx.coop_blocks / (cx.num_sms > 0 ? cx.num_sms : 1) : 1;
    long long k = (want + cx.num_sms - 1) / cx.num_sms;
    if (k < 1) k = 1;
    if (k > per_sm) k = per_sm;
    int blocks = (int) (k * cx.num_sms);
    if (cx.solve_blocks_cap > 0 && cx.solve_blocks_cap <= blocks) blocks = cx.solve_blocks_cap;    // dev / test knob
    if (work_hint <= 2048) blocks = (int) ((work_hint + SV_BLOCK - 1) / SV_BLOCK) < 1 ? 1 : (int) ((work_hint + SV_BLOCK - 1) / SV_BLOCK);
    return blocks;
}

void fx_launch_solve_colored(const DevWorld &w, const FxLaunchCtx &cx, int iterations, long long work_hint, int phases) {
    const int blocks = fx_solve_colored_blocks(cx, work_hint);
    DevWorld wc = w;
    void *args[] = { (void *) &wc, (void *) &iterations, (void *) &phases };
    const void *kernel = solve_wide(work_hint) ? (w.sol_lean ? (const void *) k_solve_colored<2, true> : (const void *) k_solve_colored<2, false>)
                                               : (w.sol_lean ? (const void *) k_solve_colored<1, true> : (const void *) k_solve_colored<1, false>);
    cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(SV_BLOCK), args, 0, cx.stream);
    ++g_fx_launches;
}

void fx_launch_solve_local(const DevWorld &w, const FxLaunchCtx &cx, int iterations, int max_bodies) {
    cudaStream_t st = cx.stream;
    const bool batch = w.group != nullptr;
    if (batch) {
        cudaMemsetAsync(w.gcnt, 0, (size_t) w.n_groups * sizeof(unsigned int), st);
        const int g = cx.num_sms * 4;
        k_group_count<<<g, 256, 0, st>>>(w); ++g_fx_launches;
        k_group_scan<<<1, 1024, 0, st>>>(w); ++g_fx_launches;
        k_group_scatter<<<g, 256, 0, st>>>(w); ++g_fx_launches;
    }
    const size_t smem = (size_t) max_bodies * 36u;      // opt-in size: fx_init_solver_attributes (per device)
    const int threads = max_bodies <= 512 ? 256 : 1024;
    int blocks = batch ? w.n_groups : 1;
    const int cap = cx.num_sms * 16;
    if (blocks > cap) blocks = cap;
    k_solve_local<<<blocks, threads, smem, st>>>(w, iterations, max_bodies); ++g_fx_launches;
}

void fx_launch_serial_reindex(const FxLaunchCtx &cx, const SerialOrder &o) { k_serial_reindex<<<1, 32, 0, cx.stream>>>(o); ++g_fx_launches; }
void fx_launch_serial_order(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o) {
    k_serial_order<<<1, 32, 0, cx.stream>>>(w, o); ++g_fx_launches;
}
void fx_launch_solve_serial_with(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o, int iterations) {
    k_solve_serial<<<1, 32, 0, cx.stream>>>(w, o, iterations); ++g_fx_launches;
}
