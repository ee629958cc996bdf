// fx_contacts.cu -- narrow phase over the candidate list + the GPU contact cache (kernels K5, K6).
//
// Replaces the body of the reference's pre-step callback, src/world.c:328-410: SAT per candidate
// pair, contact-cache delete on a miss, warm-start transfer by contact id on a hit, material mixing
// for new entries -- and the stale-entry eviction of src/world.c:222-235.
//
// Cache design.  The reference keeps one persistent stb_ds map (key = body pair) whose dense entry
// array is what the solver iterates.  Here the dense array is REBUILT every step in a deterministic
// order -- first this step's hits in candidate order, then the entries of the previous step that no
// candidate touched (such stale entries keep being solved, they are not dropped)
// -- and an open-addressing hash map (key = uidA<<32|uidB -> index into the previous dense array)
// serves the warm-start lookups.  An entry of the previous array dies iff (a) a candidate with its
// key missed (world.c:347-356), or (b) the timestamp rule fires (world.c:229-233), or (c) one of its
// bodies left the world.
//
// HBM bytes (algorithmic): narrow phase C*(8 + 2*16 + 2*8 + shape bytes) read, 112 B per hit written,
// one 12-byte hash probe per candidate; carry-over 112 B read+written per stale entry; hash rebuild
// 12 B per slot cleared + 12 B per entry.
#include <cstdlib>

#include "fx_device.h"
#include "fx_kernels.h"
extern long long g_fx_launches;
#include "fx_narrow.cuh"
#include "fx_scan.cuh"

#define NP_BLOCK 128
#define HASH_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ unsigned int hash_slot(unsigned long long key, unsigned int mask) {
    unsigned long long z = (key ^ (key >> 31)) * 0x9E3779B97F4A7C15ull;
    return (unsigned int) (z >> 32) & mask;
}

// `keys`: the key column of the dense array the table was built over (the previous set while the narrow phase runs)
__device__ __forceinline__ int hash_find(const DevWorld &w, unsigned long long key, const uint2 *keys) {
    const unsigned int mask = w.cap_hash - 1u;
    unsigned int s = hash_slot(key, mask);
    for (unsigned int probes = 0; probes < w.cap_hash; ++probes) {
        const unsigned int v = w.h_slot[s];
        if (v == 0xffffffffu) return -1;
        const uint2 k = keys[v];
        if ((((unsigned long long) k.x << 32) | (unsigned long long) k.y) == key) return (int) v;
        s = (s + 1u) & mask;
    }
    return -1;
}
__device__ __forceinline__ void hash_insert(const DevWorld &w, const uint2 *keys, unsigned int i) {
    const unsigned int mask = w.cap_hash - 1u;
    const uint2 k = keys[i];
    const unsigned long long key = ((unsigned long long) k.x << 32) | (unsigned long long) k.y;
    unsigned int s = hash_slot(key, mask);
    for (unsigned int probes = 0; probes < w.cap_hash; ++probes) {
        const unsigned int v = atomicCAS(&w.h_slot[s], 0xffffffffu, i);
        if (v == 0xffffffffu) return;
        const uint2 o = keys[v];
        if (o.x == k.x && o.y == k.y) return;            // the pair is already indexed (cannot happen: keys are unique)
        s = (s + 1u) & mask;
    }
}

__device__ __forceinline__ Xf load_xf(const DevWorld &w, int i) {
    float4 p = w.posrot[i];
    Xf t;
    t.p = v2(p.x, p.y);
    t.sn = p.z;
    t.cs = p.w;
    return t;
}

__device__ __forceinline__ Xf shfl_xf(Xf t, int src) {
    Xf r;
    r.p.x = __shfl_sync(0xffffffffu, t.p.x, src);
    r.p.y = __shfl_sync(0xffffffffu, t.p.y, src);
    r.sn = __shfl_sync(0xffffffffu, t.sn, src);
    r.cs = __shfl_sync(0xffffffffu, t.cs, src);
    return r;
}

// Shape records of the pairs a warp is about to evaluate cooperatively, staged in shared memory.
// Every lane copies the two records of ITS pair with independent 16-byte loads (so the global/L2
// latency of 32 pairs overlaps); the warp-cooperative evaluation then reads them at smem latency
// instead of paying a dependent L2 round trip per pair inside the serial ballot loop.
struct __align__(16) ShapeStage {
    ShapeDev s[2];
};

__device__ __forceinline__ void stage_shape(ShapeDev *dst, const ShapeDev *src) {
    const float4 *g = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(dst);
    float4 h0 = g[0], h1 = g[1];                       // type, nv, radius, friction | restitution, pad
    d[0] = h0;
    d[1] = h1;
    const int nv = __float_as_int(h0.y);
    if (__float_as_int(h0.x) == FX_SHAPE_POLYGON) {
        const int q = (nv + 1) >> 1;                   // float4 = two vertices
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < q) { d[2 + k] = g[2 + k]; d[6 + k] = g[6 + k]; }
    }
}

// One warp evaluates the 32 candidates its lanes hold: circle/circle per lane; every other pair by
// a group of 8 lanes (one lane per polygon axis), four pairs at a time, polygon/polygon pairs and
// circle/polygon pairs in separate sweeps so the four groups of a sweep run the same code.  A
// group's result is handed to the lane that owns the candidate through that lane's (now consumed)
// shape-stage slot.  Returns the lane's own manifold.
__device__ __forceinline__ void warp_narrowphase(const DevWorld &w, ShapeStage *stage, bool valid, int bi, int bj, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    int2 si = make_int2(-1, 0), sj = make_int2(-1, 0);
    Xf ti = { { 0.f, 0.f }, 0.f, 1.f }, tj = ti;
    if (valid) {
        si = w.shape[bi];
        sj = w.shape[bj];
        ti = load_xf(w, bi);
        tj = load_xf(w, bj);
    }
    const bool has_shapes = valid && si.x >= 0 && sj.x >= 0;
    const bool cc = has_shapes && si.y == FX_SHAPE_CIRCLE && sj.y == FX_SHAPE_CIRCLE;
    const bool pp = has_shapes && si.y == FX_SHAPE_POLYGON && sj.y == FX_SHAPE_POLYGON;
    const unsigned int lane = lane_id();
    const int grp = (int) (lane >> 3), k = (int) (lane & 7u);
    const unsigned int gmask = 0xffu << (lane & 24u);
    if (has_shapes && !cc) {
        stage_shape(&stage[lane].s[0], &w.shapes[si.x]);
        stage_shape(&stage[lane].s[1], &w.shapes[sj.x]);
    }
    if (cc) collide_circles(w.shapes[si.x].radius, ti, w.shapes[sj.x].radius, tj, m);
    __syncwarp();
    const unsigned int all_pp = __ballot_sync(0xffffffffu, pp);
    const unsigned int all_other = __ballot_sync(0xffffffffu, has_shapes && !cc && !pp);
#pragma unroll 1
    for (int sweep = 0; sweep < 2; ++sweep) {
        unsigned int todo = sweep == 0 ? all_pp : all_other;
        while (todo) {
            // the next (up to) four candidates, one per group
            int src = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                int f = todo ? __ffs(todo) - 1 : -1;
                if (todo) todo &= todo - 1u;
                if (g == grp) src = f;
            }
            const int from = src >= 0 ? src : 0;
            Xf ta = shfl_xf(ti, from), tb = shfl_xf(tj, from);
            if (src >= 0) {
                Manifold mm;
                grp_collide(gmask, k, &stage[src].s[0], ta, &stage[src].s[1], tb, mm);
                __syncwarp(gmask);
                if (k == 0) *reinterpret_cast<Manifold *>(&stage[src]) = mm;
            }
        }
    }
    __syncwarp();
    if (has_shapes && !cc) m = *reinterpret_cast<const Manifold *>(&stage[lane]);
    __syncwarp();
}

// Lane-per-pair variant (default): every lane evaluates its own candidate with the reference's
// loops; polygon data staged per lane in shared memory (LaneShapes, 8 KB per warp).
__device__ __forceinline__ void lane_narrowphase(const DevWorld &w, LaneShapes *sh, bool valid, int bi, int bj, Manifold &m) {
    m.count = 0;
    m.has_ids = 0;
    if (!valid) return;
    const int2 si = w.shape[bi], sj = w.shape[bj];
    if (si.x < 0 || sj.x < 0) return;
    const Xf ti = load_xf(w, bi), tj = load_xf(w, bj);
    const ShapeDev *ga = &w.shapes[si.x], *gb = &w.shapes[sj.x];
    const int lane = (int) lane_id();
    const float4 ha = *reinterpret_cast<const float4 *>(ga), hb = *reinterpret_cast<const float4 *>(gb);
    const int ta = __float_as_int(ha.x), tb = __float_as_int(hb.x);     // live shape types
    LanePoly pa = { sh, 0, lane, __float_as_int(ha.y) }, pb = { sh, 1, lane, __float_as_int(hb.y) };
    if (ta == FX_SHAPE_POLYGON) lane_stage_poly(sh, 0, lane, ga, pa.nv);
    if (tb == FX_SHAPE_POLYGON) lane_stage_poly(sh, 1, lane, gb, pb.nv);
    if (ta == FX_SHAPE_CIRCLE && tb == FX_SHAPE_CIRCLE) collide_circles(ha.z, ti, hb.z, tj, m);
    else if (ta == FX_SHAPE_CIRCLE && tb == FX_SHAPE_POLYGON) lane_collide_circle_poly(true, ha.z, pb, ti, tj, m);
    else if (ta == FX_SHAPE_POLYGON && tb == FX_SHAPE_CIRCLE) lane_collide_circle_poly(false, hb.z, pa, ti, tj, m);
    else if (ta == FX_SHAPE_POLYGON && tb == FX_SHAPE_POLYGON) lane_collide_polys(pa, ti, pb, tj, m);
}

union NarrowStage {
    ShapeStage group[NP_BLOCK];                 // 8-lanes-per-pair variant
    LaneShapes lane[NP_BLOCK / 32];             // lane-per-pair variant
};

__global__ void __launch_bounds__(NP_BLOCK) k_narrowphase(DevWorld w, unsigned long long *status, unsigned int *ticket) {
    __shared__ NarrowStage s_stage;
    const unsigned int ncand = w.ctr->n_cand;
    const float stamp = w.scal->timestamp;
    const unsigned int ntiles = (ncand + NP_BLOCK - 1) / NP_BLOCK;
    const int cs = *w.cur_flip;
    const ManifoldSet cur = w.man[cs], prev = w.man[cs ^ 1];
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.ctr->n_hits = 0;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if (tile >= ntiles) return;
        const unsigned int c = tile * NP_BLOCK + threadIdx.x;
        const bool valid = c < ncand;
        int2 pr = valid ? w.cand[c] : make_int2(0, 0);
        Manifold m;
        if (w.narrow_mode != 1) lane_narrowphase(w, &s_stage.lane[threadIdx.x >> 5], valid, pr.x, pr.y, m);
        else warp_narrowphase(w, &s_stage.group[(threadIdx.x >> 5) * 32], valid, pr.x, pr.y, m);
        const bool hit = valid && m.count > 0;

        // contact-cache probe (hit or miss): src/world.c:347-394
        unsigned long long key = 0ull;
        int old = -1;
        float friction = 0.0f, restitution = 0.0f;
        float lam_n[2] = { 0.0f, 0.0f }, lam_t[2] = { 0.0f, 0.0f };
        int color = -1;
        if (valid) {
            key = ((unsigned long long) w.uid[pr.x] << 32) | (unsigned long long) w.uid[pr.y];
            old = hash_find(w, key, prev.key);
            if (old >= 0) w.touched[old] = 1;
            w.cand_hit[c] = hit ? 1 : 0;
        }
        if (hit) {
            if (old >= 0) {
                float4 oh = prev.hdr[old];
                friction = oh.z;
                restitution = oh.w;
                const int2 ometa = prev.meta[old];
                const int ocount = ometa.x;
                color = ometa.y;                      // keep last step's colour (validated by the solver)
                int oid0 = prev.cid[2 * old], oid1 = prev.cid[2 * old + 1];
                for (int i = 0; i < m.count; ++i) {
                    int oi = -1;
                    if (ocount > 0 && m.id[i] == oid0) oi = 0;
                    else if (ocount > 1 && m.id[i] == oid1) oi = 1;
                    if (oi < 0) continue;
                    float4 oimp = prev.imp[2 * old + oi];
                    lam_n[i] = oimp.y;
                    lam_t[i] = oimp.w;
                }
            } else {                                   // new entry: src/world.c:395-404
                const ShapeDev *s1 = &w.shapes[w.shape[pr.x].x], *s2 = &w.shapes[w.shape[pr.y].x];
                friction = 0.5f * (s1->friction + s2->friction);
                restitution = pin_fminf(s1->restitution, s2->restitution);
                if (friction < 0.0f) friction = 0.0f;
                if (restitution < 0.0f) restitution = 0.0f;
            }
        }
        unsigned int total;
        unsigned int excl = block_exclusive_scan<NP_BLOCK>(hit ? 1u : 0u, &total);
        unsigned int base = tile_exclusive_prefix(status, tile, total);
        if (hit) {
            unsigned int pos = base + excl;
            if (pos < w.cap_manifolds) {
                cur.key[pos] = make_uint2(w.uid[pr.x], w.uid[pr.y]);
                cur.body[pos] = pr;
                cur.hdr[pos] = make_float4(m.dir.x, m.dir.y, friction, restitution);
                cur.meta[pos] = make_int2(m.count, color);
                // contacts[1] of a one-point manifold is a copy taken BEFORE stamping / warm-start
                // transfer (src/collision.c:292,374,...; src/world.c:362-394): id/point/depth only
                cur.geo[2 * pos] = make_float4(m.p[0].x, m.p[0].y, m.depth[0], stamp);
                cur.geo[2 * pos + 1] = make_float4(m.p[1].x, m.p[1].y, m.depth[1], (m.count > 1) ? stamp : 0.0f);
                cur.imp[2 * pos] = make_float4(0.0f, lam_n[0], 0.0f, lam_t[0]);
                cur.imp[2 * pos + 1] = make_float4(0.0f, (m.count > 1) ? lam_n[1] : 0.0f, 0.0f, (m.count > 1) ? lam_t[1] : 0.0f);
                cur.cid[2 * pos] = m.id[0];
                cur.cid[2 * pos + 1] = m.id[1];
            }
        }
        // contact points of this tile -> Q
        unsigned int q = hit ? (unsigned int) m.count : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
        if (lane_id() == 0 && q) atomicAdd(&w.ctr->n_contacts, q);
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            unsigned int h = base + total;
            if (h > w.cap_manifolds) { atomicOr(&w.ctr->overflow, 4u); h = w.cap_manifolds; }
            w.ctr->n_hits = h;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Narrow phase, production layout: 1,024 candidates per tile, 256 threads.
//   A  every thread looks at 4 CONSECUTIVE candidates: circle/circle pairs are solved on the spot,
//      the others are appended to a per-block circle/polygon list or polygon/polygon list (smem)
//   B  the lists are processed with dense lanes (one lane per pair, all 32 lanes of a warp on the
//      same code path): type divergence is gone ([measured] 5.8 active threads per warp before)
//   C  results sit in per-candidate smem slots, so the output stays in CANDIDATE order: one block
//      scan + one block-wide look-back per 1,024 candidates, then probe / warm-start transfer / write
// ---------------------------------------------------------------------------------------------
#define NP2_BLOCK 256
// candidates per thread: 4 (tiles of 1,024) for large worlds, 2 for small ones, where more, smaller tiles keep more
// blocks in flight ([measured] narrow phase at 65k bodies 0.091 -> 0.077 ms with 2; at 1 M bodies 0.525 -> 0.557 ms)

struct NpSlot {
    float dx, dy, p0x, p0y, p1x, p1y, d0, d1;
    int id0, id1, count;
};

__device__ __forceinline__ void slot_store(NpSlot &s, const Manifold &m) {
    s.dx = m.dir.x; s.dy = m.dir.y;
    s.p0x = m.p[0].x; s.p0y = m.p[0].y; s.p1x = m.p[1].x; s.p1y = m.p[1].y;
    s.d0 = m.depth[0]; s.d1 = m.depth[1];
    s.id0 = m.id[0]; s.id1 = m.id[1];
    s.count = m.count;
}

#define NP2_SMEM(ITEMS) ((size_t) NP2_BLOCK * (ITEMS) * (sizeof(NpSlot) + 2 * sizeof(unsigned short)))

template <int NP2_ITEMS>
__global__ void __launch_bounds__(NP2_BLOCK, NP2_ITEMS >= 4 ? 3 : 4) k_narrowphase_tiled(DevWorld w, unsigned long long *status, unsigned int *ticket) {
    constexpr unsigned int NP2_TILE = NP2_BLOCK * NP2_ITEMS;
    extern __shared__ __align__(16) unsigned char np_smem[];
    NpSlot *slot = reinterpret_cast<NpSlot *>(np_smem);
    unsigned short *list_cp = reinterpret_cast<unsigned short *>(np_smem + NP2_TILE * sizeof(NpSlot));
    unsigned short *list_pp = list_cp + NP2_TILE;
    __shared__ unsigned int s_ncp, s_npp;
    const unsigned int ncand = w.ctr->n_cand;
    const float stamp = w.scal->timestamp;
    const unsigned int ntiles = (ncand + NP2_TILE - 1) / NP2_TILE;
    const int cs = *w.cur_flip;
    const ManifoldSet cur = w.man[cs], prev = w.man[cs ^ 1];
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.ctr->n_hits = 0;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if (tile >= ntiles) return;
        if (threadIdx.x == 0) { s_ncp = 0u; s_npp = 0u; }
        __syncthreads();
        const unsigned int first = tile * NP2_TILE + threadIdx.x * NP2_ITEMS;
        int2 pr[NP2_ITEMS];
        // ---- A: classify; circle/circle on the spot
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            const unsigned int c = first + j, local = threadIdx.x * NP2_ITEMS + j;
            pr[j] = (c < ncand) ? w.cand[c] : make_int2(-1, -1);
            slot[local].count = 0;
            if (pr[j].x < 0) continue;
            const int2 si = w.shape[pr[j].x], sj = w.shape[pr[j].y];
            if (si.x < 0 || sj.x < 0) continue;
            if (si.y == FX_SHAPE_CIRCLE && sj.y == FX_SHAPE_CIRCLE) {
                Manifold m;
                collide_circles(w.shapes[si.x].radius, load_xf(w, pr[j].x), w.shapes[sj.x].radius, load_xf(w, pr[j].y), m);
                if (m.count) slot_store(slot[local], m);
            } else {
                // Two shapes whose bounding circles are apart have a separating axis: the reference's SAT / circle
                // tests return "no collision" for them (src/collision.c:340-348,465-469), so this candidate is a miss
                // without running them.  The reference has no such test (it goes from shared cells straight to
                // SAT, App. A.2) and at cell = 2 body diameters most polygon candidates are this far apart.
                const float4 pi4 = w.posrot[pr[j].x], pj4 = w.posrot[pr[j].y];
                const float dx = pj4.x - pi4.x, dy = pj4.y - pi4.y;
                const float reach = w.shapes[si.x].bound + w.shapes[sj.x].bound;
                if (dx * dx + dy * dy > reach * reach * 1.0001f) continue;      // NaN compares false: falls through
                if (si.y == FX_SHAPE_POLYGON && sj.y == FX_SHAPE_POLYGON) {
                    list_pp[atomicAdd(&s_npp, 1u)] = (unsigned short) local;
                } else if ((si.y == FX_SHAPE_CIRCLE && sj.y == FX_SHAPE_POLYGON) || (si.y == FX_SHAPE_POLYGON && sj.y == FX_SHAPE_CIRCLE)) {
                    list_cp[atomicAdd(&s_ncp, 1u)] = (unsigned short) local;
                }
            }
        }
        __syncthreads();
        // ---- B: dense sweeps over the two lists, AT THE SAME TIME: after the early miss a tile keeps a few dozen pairs
        // of each kind, far fewer than the block has lanes, so the lower half of the block takes the circle/polygon
        // list while the upper half takes the polygon/polygon list (warp-uniform split)
        // ([measured] 65k bodies, 512-candidate tiles: narrow phase 0.079 -> 0.073 ms; with lists longer than half a block
        // the split costs extra passes -- 4,096 batched worlds, 1,024-candidate tiles: 0.545 -> 0.735 ms -- so those tiles
        // give every list the whole block, and the 1,024-candidate build never splits)
        const unsigned int ncp = s_ncp, npp = s_npp;
        const bool split = NP2_ITEMS <= 2 && ncp <= NP2_BLOCK / 2 && npp <= NP2_BLOCK / 2;    // small-world build only
        const unsigned int half = split ? NP2_BLOCK / 2 : NP2_BLOCK, t0 = split ? (threadIdx.x & (NP2_BLOCK / 2 - 1u)) : threadIdx.x;
        if (!split || threadIdx.x < NP2_BLOCK / 2) {
            for (unsigned int t = t0; t < ncp; t += half) {
                const unsigned int local = list_cp[t];
                const int2 p = w.cand[tile * NP2_TILE + local];
                const int2 si = w.shape[p.x], sj = w.shape[p.y];
                const bool circle_first = (si.y == FX_SHAPE_CIRCLE);
                const ShapeDev *sc = &w.shapes[circle_first ? si.x : sj.x], *sp = &w.shapes[circle_first ? sj.x : si.x];
                GlobalPoly poly = { sp, sp->nv };
                Manifold m;
                lane_collide_circle_poly(circle_first, sc->radius, poly, load_xf(w, p.x), load_xf(w, p.y), m);
                if (m.count) slot_store(slot[local], m);
            }
        }
        if (!split || threadIdx.x >= NP2_BLOCK / 2) {
            for (unsigned int t = t0; t < npp; t += half) {
                const unsigned int local = list_pp[t];
                const int2 p = w.cand[tile * NP2_TILE + local];
                const ShapeDev *sa = &w.shapes[w.shape[p.x].x], *sb = &w.shapes[w.shape[p.y].x];
                GlobalPoly pa = { sa, sa->nv }, pb = { sb, sb->nv };
                Manifold m;
                lane_collide_polys(pa, load_xf(w, p.x), pb, load_xf(w, p.y), m);
                if (m.count) slot_store(slot[local], m);
            }
        }
        __syncthreads();
        // ---- C: ordered compaction (candidate order), cache probe, write.
        // Every candidate -- hit or miss -- probes the previous step's cache (src/world.c:347-394), and a hit then reads
        // its old entry: a chain of five dependent scattered loads (uid -> table slot -> key -> header / ids -> impulses).
        // The chains of a thread's candidates are walked SIDE BY SIDE, one level at a time, with no store in between (a
        // store would pin the later loads behind it), and the probe goes out before the block's look-back, which it does
        // not depend on; [measured] one candidate after the other: 57 % of the kernel's stall samples sat here, on
        // the long scoreboard at 13 active lanes per warp.
        unsigned int mine = 0u, q = 0u;
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            const int cnt = slot[threadIdx.x * NP2_ITEMS + j].count;
            mine += cnt > 0 ? 1u : 0u;
            q += (unsigned int) cnt;
        }
        unsigned int total;
        const unsigned int excl = block_exclusive_scan<NP2_BLOCK>(mine, &total);
        tile_publish_aggregate(status, tile, total);                // the tiles behind this one do not wait for its probes
        // level 1: the uids of both bodies
        unsigned int ua[NP2_ITEMS], ub[NP2_ITEMS];
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            ua[j] = ub[j] = 0u;
            if (pr[j].x >= 0) { ua[j] = w.uid[pr[j].x]; ub[j] = w.uid[pr[j].y]; }
        }
        // level 2: the first table slot of every probe; level 3: the key it points at
        const unsigned int hmask = w.cap_hash - 1u;
        unsigned int hv[NP2_ITEMS];
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            hv[j] = 0xffffffffu;
            if (pr[j].x >= 0) hv[j] = w.h_slot[hash_slot(((unsigned long long) ua[j] << 32) | (unsigned long long) ub[j], hmask)];
        }
        uint2 hk[NP2_ITEMS];
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            hk[j] = make_uint2(0u, 0u);
            if (hv[j] != 0xffffffffu) hk[j] = prev.key[hv[j]];
        }
        int old[NP2_ITEMS];
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            old[j] = -1;
            if (hv[j] == 0xffffffffu) continue;                     // empty slot: the pair has no entry
            if (hk[j].x == ua[j] && hk[j].y == ub[j]) { old[j] = (int) hv[j]; continue; }
            // the slot belongs to another pair: walk on from it (rare; hash_find restarts at the home slot)
            old[j] = hash_find(w, ((unsigned long long) ua[j] << 32) | (unsigned long long) ub[j], prev.key);
        }
#pragma unroll
        for (int j = 0; j < NP2_ITEMS; ++j) {
            if (pr[j].x < 0) continue;
            if (old[j] >= 0) w.touched[old[j]] = 1;
            w.cand_hit[first + j] = slot[threadIdx.x * NP2_ITEMS + j].count > 0 ? 1 : 0;
        }
        const unsigned int base = tile_exclusive_prefix(status, tile, total);
        unsigned int pos = base + excl;
        // level 4: header, contact count / colour and feature ids of the old entries of this thread's hits -- two candidates at
        // a time (four would not fit the 64 registers that keep four blocks on an SM)
        constexpr int NP2_PAIR = NP2_ITEMS < 2 ? 1 : 2;
#pragma unroll
        for (int j0 = 0; j0 < NP2_ITEMS; j0 += NP2_PAIR) {
            float4 oh[NP2_PAIR];
            int2 ometa[NP2_PAIR];
            int oid0[NP2_PAIR], oid1[NP2_PAIR];
#pragma unroll
            for (int i = 0; i < NP2_PAIR; ++i) {
                const int j = j0 + i;
                oh[i] = make_float4(0.f, 0.f, 0.f, 0.f); ometa[i] = make_int2(0, -1); oid0[i] = oid1[i] = 0;
                if (old[j] >= 0 && slot[threadIdx.x * NP2_ITEMS + j].count > 0) {
                    oh[i] = prev.hdr[old[j]];
                    ometa[i] = prev.meta[old[j]];
                    oid0[i] = prev.cid[2 * old[j]];
                    oid1[i] = prev.cid[2 * old[j] + 1];
                }
            }
#pragma unroll
            for (int i = 0; i < NP2_PAIR; ++i) {
                const int j = j0 + i;
                if (pr[j].x < 0) continue;
                const NpSlot &sl = slot[threadIdx.x * NP2_ITEMS + j];
                if (sl.count <= 0) continue;
                float friction, restitution;
                float lam_n0 = 0.0f, lam_t0 = 0.0f, lam_n1 = 0.0f, lam_t1 = 0.0f;
                int color = -1;
                if (old[j] >= 0) {
                    friction = oh[i].z;
                    restitution = oh[i].w;
                    color = ometa[i].y;                                 // keep last step's colour
                    int o0 = -1, o1 = -1;
                    if (ometa[i].x > 0 && sl.id0 == oid0[i]) o0 = 0; else if (ometa[i].x > 1 && sl.id0 == oid1[i]) o0 = 1;
                    if (sl.count > 1) { if (ometa[i].x > 0 && sl.id1 == oid0[i]) o1 = 0; else if (ometa[i].x > 1 && sl.id1 == oid1[i]) o1 = 1; }
                    // level 5: the impulses (the two loads of an entry are independent of each other)
                    float4 oi0 = make_float4(0.f, 0.f, 0.f, 0.f), oi1 = oi0;
                    if (o0 >= 0) oi0 = prev.imp[2 * old[j] + o0];
                    if (o1 >= 0) oi1 = prev.imp[2 * old[j] + o1];
                    if (o0 >= 0) { lam_n0 = oi0.y; lam_t0 = oi0.w; }
                    if (o1 >= 0) { lam_n1 = oi1.y; lam_t1 = oi1.w; }
                } else {                                                // new entry: src/world.c:395-404
                    const ShapeDev *s1 = &w.shapes[w.shape[pr[j].x].x], *s2 = &w.shapes[w.shape[pr[j].y].x];
                    friction = 0.5f * (s1->friction + s2->friction);
                    restitution = pin_fminf(s1->restitution, s2->restitution);
                    if (friction < 0.0f) friction = 0.0f;
                    if (restitution < 0.0f) restitution = 0.0f;
                }
                if (pos < w.cap_manifolds) {
                    const bool two = sl.count > 1;
                    cur.key[pos] = make_uint2(ua[j], ub[j]);
                    cur.body[pos] = pr[j];
                    cur.hdr[pos] = make_float4(sl.dx, sl.dy, friction, restitution);
                    cur.meta[pos] = make_int2(sl.count, color);
                    cur.geo[2 * pos] = make_float4(sl.p0x, sl.p0y, sl.d0, stamp);
                    cur.geo[2 * pos + 1] = make_float4(sl.p1x, sl.p1y, sl.d1, two ? stamp : 0.0f);
                    cur.imp[2 * pos] = make_float4(0.0f, lam_n0, 0.0f, lam_t0);
                    cur.imp[2 * pos + 1] = make_float4(0.0f, two ? lam_n1 : 0.0f, 0.0f, two ? lam_t1 : 0.0f);
                    cur.cid[2 * pos] = sl.id0;
                    cur.cid[2 * pos + 1] = sl.id1;
                }
                ++pos;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
        if (lane_id() == 0 && q) atomicAdd(&w.ctr->n_contacts, q);
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            unsigned int h = base + total;
            if (h > w.cap_manifolds) { atomicOr(&w.ctr->overflow, 4u); h = w.cap_manifolds; }
            w.ctr->n_hits = h;
        }
    }
}

// stale entries of the previous array that no candidate probed survive (App. A.7) unless the
// timestamp rule (src/world.c:229-233) or a departed body removes them
// Tiles of NP_BLOCK x CO_ITEMS entries: almost every entry of a settled pile was probed this step, so the work per
// tile is tiny and the kernel is bound by the length of the look-back chain ([measured] 1 M bodies, 128-entry
// tiles: 140 us for 3.7 MB of traffic).
template <int CO_ITEMS>
__global__ void __launch_bounds__(NP_BLOCK) k_carry_over(DevWorld w, unsigned long long *status, unsigned int *ticket) {
    const unsigned int nprev = *w.prev_count;
    const float stamp = w.scal->timestamp, dt = w.scal->dt;
    const unsigned int tile_size = NP_BLOCK * CO_ITEMS;
    const unsigned int ntiles = (nprev + tile_size - 1) / tile_size;
    const int cs = *w.cur_flip;
    const ManifoldSet cur = w.man[cs], prev = w.man[cs ^ 1];
    const unsigned int nhits = w.ctr->n_hits;
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.ctr->n_manifolds = nhits;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if (tile >= ntiles) return;
        const unsigned int first = tile * tile_size + threadIdx.x * CO_ITEMS;
        unsigned int keep_mask = 0u, kept = 0u, q = 0u;
        int2 body[CO_ITEMS];
#pragma unroll
        for (int j = 0; j < CO_ITEMS; ++j) {
            const unsigned int p = first + j;
            body[j] = make_int2(-1, -1);
            if (p < nprev && !w.touched[p]) {
                uint2 k = prev.key[p];
                body[j].x = w.idx_of_uid[k.x];
                body[j].y = w.idx_of_uid[k.y];
                const int count = prev.meta[p].x;
                bool keep = body[j].x >= 0 && body[j].y >= 0;
                for (int c = 0; c < count && keep; ++c)
                    if (stamp - prev.geo[2 * p + c].w > dt) keep = false;
                if (keep) { keep_mask |= 1u << j; ++kept; q += (unsigned int) count; }
            }
        }
        unsigned int total;
        unsigned int excl = block_exclusive_scan<NP_BLOCK>(kept, &total);
        unsigned int base = tile_exclusive_prefix(status, tile, total);
        unsigned int pos = nhits + base + excl;
#pragma unroll
        for (int j = 0; j < CO_ITEMS; ++j) {
            if (!(keep_mask & (1u << j))) continue;
            const unsigned int p = first + j;
            if (pos < w.cap_manifolds) {
                cur.key[pos] = prev.key[p];
                cur.body[pos] = body[j];
                cur.hdr[pos] = prev.hdr[p];
                cur.meta[pos] = prev.meta[p];
                cur.geo[2 * pos] = prev.geo[2 * p];
                cur.geo[2 * pos + 1] = prev.geo[2 * p + 1];
                cur.imp[2 * pos] = prev.imp[2 * p];
                cur.imp[2 * pos + 1] = prev.imp[2 * p + 1];
                cur.cid[2 * pos] = prev.cid[2 * p];
                cur.cid[2 * pos + 1] = prev.cid[2 * p + 1];
            }
            ++pos;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
        if (lane_id() == 0 && q) atomicAdd(&w.ctr->n_contacts, q);
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            unsigned int mcount = nhits + base + total;
            if (mcount > w.cap_manifolds) { atomicOr(&w.ctr->overflow, 4u); mcount = w.cap_manifolds; }
            w.ctr->n_manifolds = mcount;
        }
    }
}

// Four insertions per thread side by side: the four keys are loaded, then the four claims (atomicCAS on the home slot) go
// out back to back -- an insertion is one L2 round trip that nothing else in the thread depends on -- and only a claim
// that found its slot taken walks on ([measured] one at a time, 2.4 M manifolds: 67 us at 199 long-scoreboard cycles per
// issued instruction).  Small worlds only ([measured] cache stage at 65k bodies 28.6 -> 26.0 us; at 1 M bodies, where the
// table no longer sits in L2 and the atomics are throughput-bound, 167 -> 191 us: the large-world set keeps one at a time).
__global__ void __launch_bounds__(NP_BLOCK) k_hash_build(DevWorld w) {
    const unsigned int m = w.ctr->n_manifolds;
    const ManifoldSet cur = w.man[*w.cur_flip];
    const unsigned int nth = gridDim.x * blockDim.x, mask = w.cap_hash - 1u;
    if (w.sol_mult == 1u) {                                   // the large-world kernel set (fx_world.cpp fill_dev_params)
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += nth) hash_insert(w, cur.key, i);
        return;
    }
    for (unsigned int base = blockIdx.x * blockDim.x + threadIdx.x; base < m; base += 4u * nth) {
        uint2 k[4];
        unsigned int home[4], seen[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int i = base + (unsigned int) j * nth;
            k[j] = make_uint2(0u, 0u);
            if (i < m) k[j] = cur.key[i];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int i = base + (unsigned int) j * nth;
            home[j] = hash_slot(((unsigned long long) k[j].x << 32) | (unsigned long long) k[j].y, mask);
            seen[j] = 0xffffffffu;
            if (i < m) seen[j] = atomicCAS(&w.h_slot[home[j]], 0xffffffffu, i);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int i = base + (unsigned int) j * nth;
            if (i >= m || seen[j] == 0xffffffffu) continue;             // claimed its home slot
            // taken: linear probing from the next slot on (hash_insert, minus the probe that was just made)
            unsigned int sl = (home[j] + 1u) & mask;
            for (unsigned int probes = 1; probes < w.cap_hash; ++probes) {
                const unsigned int v = atomicCAS(&w.h_slot[sl], 0xffffffffu, i);
                if (v == 0xffffffffu) break;
                sl = (sl + 1u) & mask;
            }
        }
    }
}

// after the table was re-allocated: index the PREVIOUS dense array again (warm-start lookups)
__global__ void __launch_bounds__(NP_BLOCK) k_hash_build_prev(DevWorld w) {
    const unsigned int m = *w.prev_count;
    const ManifoldSet prev = w.man[*w.cur_flip ^ 1];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) hash_insert(w, prev.key, i);
}

static inline int grid_for(long long n, int block, int max_blocks) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int) g;
}

void fx_launch_rehash_prev(const DevWorld &w, const FxLaunchCtx &cx) {
    k_hash_build_prev<<<grid_for((long long) w.cap_manifolds, NP_BLOCK, cx.max_blocks), NP_BLOCK, 0, cx.stream>>>(w); ++g_fx_launches;
}

void fx_init_solver_attributes();
// per-DEVICE opt-in shared memory sizes (a process may hold worlds on several GPUs)
void fx_init_kernel_attributes() {
    cudaFuncSetAttribute(k_narrowphase_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) NP2_SMEM(4));
    fx_init_solver_attributes();
}

void fx_launch_narrowphase(const DevWorld &w, const FxLaunchCtx &cx) {
    cudaStream_t s = cx.stream;
    cudaMemsetAsync(w.touched, 0, w.cap_manifolds, s);
    if (w.narrow_mode == 0) {
        static const int force_items = [] { const char *e = std::getenv("FEROX_NP_ITEMS"); return e ? std::atoi(e) : 0; }();   // dev knob
        const int items = force_items ? force_items : (w.sol_mult == 1u ? 4 : 2);
        if (items == 4) k_narrowphase_tiled<4><<<grid_for((long long) w.cap_cand, NP2_BLOCK * 4, cx.max_blocks), NP2_BLOCK, NP2_SMEM(4), s>>>(w, cx.status[2], cx.tickets + 6);
        else if (items == 2) k_narrowphase_tiled<2><<<grid_for((long long) w.cap_cand, NP2_BLOCK * 2, cx.max_blocks), NP2_BLOCK, NP2_SMEM(2), s>>>(w, cx.status[2], cx.tickets + 6);
        else k_narrowphase_tiled<1><<<grid_for((long long) w.cap_cand, NP2_BLOCK, cx.max_blocks), NP2_BLOCK, NP2_SMEM(1), s>>>(w, cx.status[2], cx.tickets + 6);
        ++g_fx_launches;
        return;
    }
    k_narrowphase<<<grid_for((long long) w.cap_cand, NP_BLOCK, cx.max_blocks), NP_BLOCK, 0, s>>>(w, cx.status[2], cx.tickets + 6); ++g_fx_launches;
}

void fx_launch_cache_update(const DevWorld &w, const FxLaunchCtx &cx) {
    cudaStream_t s = cx.stream;
    // large worlds (the host's contact-major switch, fill_dev_params): 1,024-entry tiles shorten the look-back chain;
    // small ones keep one entry per thread ([measured] 65k bodies: 8 entries per thread cost +6 us)
    if (w.sol_mult == 1u) k_carry_over<8><<<grid_for((long long) w.cap_manifolds, NP_BLOCK * 8, cx.max_blocks), NP_BLOCK, 0, s>>>(w, cx.status[3], cx.tickets + 7);
    else k_carry_over<1><<<grid_for((long long) w.cap_manifolds, NP_BLOCK, cx.max_blocks), NP_BLOCK, 0, s>>>(w, cx.status[3], cx.tickets + 7);
    ++g_fx_launches;
    cudaMemsetAsync(w.h_slot, 0xff, (size_t) w.cap_hash * sizeof(unsigned int), s);
    k_hash_build<<<grid_for((long long) w.cap_manifolds, NP_BLOCK, cx.max_blocks), NP_BLOCK, 0, s>>>(w); ++g_fx_launches;
}

// ---------------------------------------------------------------------------------------------
// batched stand-alone narrow phase (frComputeCollision / frxComputeCollisions)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NP_BLOCK) k_collide_pairs(DevWorld w, const int2 *pairs, int n, Manifold *out) {
    __shared__ NarrowStage s_stage;
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (n + stride - 1) / stride;
    for (int r = 0; r < rounds; ++r) {
        int c = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool valid = c < n;
        int2 pr = valid ? pairs[c] : make_int2(0, 0);
        Manifold m;
        if (w.narrow_mode == 0) lane_narrowphase(w, &s_stage.lane[threadIdx.x >> 5], valid, pr.x, pr.y, m);
        else warp_narrowphase(w, &s_stage.group[(threadIdx.x >> 5) * 32], valid, pr.x, pr.y, m);
        if (valid) out[c] = m;
    }
}

void fx_launch_collide_pairs(const DevWorld &w, const FxLaunchCtx &cx, const int2 *pairs, int n, Manifold *out) {
    if (n <= 0) return;
    k_collide_pairs<<<grid_for(n, NP_BLOCK, cx.max_blocks), NP_BLOCK, 0, cx.stream>>>(w, pairs, n, out); ++g_fx_launches;
}
