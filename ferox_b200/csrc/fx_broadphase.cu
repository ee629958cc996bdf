// fx_broadphase.cu -- sort-based spatial hash (kernels K1-K4 of SURVEY.md section 2.2).
//
// Replaces reference src/world.c:432-444 + src/broad_phase.c:99-179 (insert every body into a hash of
// cells, then query every body and de-duplicate) and, fused into the first pass, the velocity
// integration of src/world.c:216-220.  The candidate rule is the reference's, exactly: bodies i < j
// are candidates iff their inclusive, truncated-toward-zero cell rectangles intersect and
// invMass(i) + invMass(j) > 0 -- there is no AABB test.
//
// Stages (all sizes live on the device; no host round trip inside a step):
//   k_body_prepass   per body: [gravity + velocity integration], cell rectangle, cells covered,
//                    world cell bounds (atomics), exclusive offsets via decoupled look-back scan
//   k_emit_entries   dense key geometry (grid_w/h, key bits, radix passes: plan_keys), then per body:
//                    (cellKey, packedBody) for every covered cell; bodies covering more
//                    than FX_LARGE_CELLS cells are expanded by a whole block each (e.g. a ground
//                    plane spanning thousands of cells)
//   k_sort_pass (x passes)  onesweep-style LSD radix sort, 8-bit digits (histograms counted by k_emit_entries), one
//                    kernel per digit: tiles ranked per warp, reordered in shared memory, chained-scan look-back
//                    per digit, streamed out in digit order; stable, deterministic
//   k_emit_pairs     per sorted entry: walk the rest of its cell run; a pair is emitted only in its
//                    canonical cell (the min-corner of the rectangle intersection), which removes
//                    duplicates without a hash set; ordered append via look-back
// HBM bytes per step (algorithmic): prepass 68 N, emit 16 N + 8 K, sort (8+8) K per pass + 4 K
// histogram, pair emission 8 K + 8 C.
#include "fx_device.h"
#include "fx_kernels.h"
extern long long g_fx_launches;
#include "fx_scan.cuh"

#define BP_BLOCK 256
#define FX_LARGE_CELLS 32u
#define FX_MAX_CELLS_PER_BODY (1u << 24)

#ifndef SORT_ITEMS
#define SORT_ITEMS 8
#endif
#define SORT_TILE (BP_BLOCK * SORT_ITEMS)

// packed entry value: [28:0] body index, 29: entry is in the body's min-x column, 30: min-y row,
// 31: body has invMass > 0
#define ENT_BODY(v) ((v) & 0x1fffffffu)
#define ENT_X0 (1u << 29)
#define ENT_Y0 (1u << 30)
#define ENT_MASS (1u << 31)

struct CellRect {
    int x0, y0, x1, y1;
};

__device__ __forceinline__ CellRect body_rect(float4 a, float inv_cell) {
    CellRect r;
    r.x0 = cell_coord(a.x, inv_cell);
    r.y0 = cell_coord(a.y, inv_cell);
    r.x1 = cell_coord(a.x + a.z, inv_cell);
    r.y1 = cell_coord(a.y + a.w, inv_cell);
    return r;
}

__device__ __forceinline__ unsigned int rect_cells(CellRect r, bool *bad) {
    *bad = false;
    if (r.x1 < r.x0 || r.y1 < r.y0) return 0u;
    unsigned long long w = (unsigned long long) ((long long) r.x1 - r.x0 + 1);
    unsigned long long h = (unsigned long long) ((long long) r.y1 - r.y0 + 1);
    unsigned long long c = w * h;
    if (c > FX_MAX_CELLS_PER_BODY) { *bad = true; return 0u; }
    return (unsigned int) c;
}

// gravity + velocity integration, reference src/rigid_body.c:309-316 and 418-427 (called from
// src/world.c:216-220).  The force accumulator is consumed here and cleared by the position pass.
__device__ __forceinline__ float4 integrate_velocity(float4 v, float4 mp, float4 f, float gx, float gy, float dt) {
    float mass = mp.x, inv_inertia = mp.y, gscale = mp.z, inv_mass = v.w;
    if (mass > 0.0f) {
        float k = gscale * mass;
        f.x = f.x + gx * k;
        f.y = f.y + gy * k;
    }
    if (inv_mass > 0.0f && dt > 0.0f) {
        float s = inv_mass * dt;
        v.x = v.x + f.x * s;
        v.y = v.y + f.y * s;
        v.z += (f.z * inv_inertia) * dt;
    }
    return v;
}

// store_force: the accumulators keep the gravity term, as the reference's do until frPostStepWorld clears them
// (src/rigid_body.c:309-316); only a world with a collision handler can observe that (frGetBodyForce in postStep)
__global__ void __launch_bounds__(BP_BLOCK) k_integrate_velocity(DevWorld w, int store_force) {
    const StepScalars sc = *w.scal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.n; i += gridDim.x * blockDim.x) {
        const float4 mp = w.mprop[i];
        float4 f = w.force[i];
        w.vel[i] = integrate_velocity(w.vel[i], mp, f, sc.gx, sc.gy, sc.dt);
        if (store_force && mp.x > 0.0f) {
            const float k = mp.z * mp.x;
            f.x = f.x + sc.gx * k;
            f.y = f.y + sc.gy * k;
            w.force[i] = f;
        }
    }
}

__global__ void __launch_bounds__(BP_BLOCK) k_body_prepass(DevWorld w, int fuse_velocity, unsigned long long *status,
                                                           unsigned int *ticket) {
    const int ntiles = (w.n + BP_BLOCK - 1) / BP_BLOCK;
    const StepScalars sc = *w.scal;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if ((int) tile >= ntiles) return;
        int i = (int) tile * BP_BLOCK + (int) threadIdx.x;
        unsigned int cnt = 0u;
        CellRect r = { 0x7fffffff, 0x7fffffff, (int) 0x80000000, (int) 0x80000000 };
        bool valid = i < w.n;
        if (valid) {
            if (fuse_velocity)
                w.vel[i] = integrate_velocity(w.vel[i], w.mprop[i], w.force[i], sc.gx, sc.gy, sc.dt);
            CellRect rr = body_rect(w.aabb[i], w.inv_cell);
            bool bad;
            cnt = rect_cells(rr, &bad);
            if (bad) atomicOr(&w.ctr->overflow, 16u);
            if (cnt) r = rr;
            if (cnt > FX_LARGE_CELLS) {
                unsigned int slot = atomicAdd(&w.ctr->n_large, 1u);
                if (slot < w.cap_large) w.large_list[slot] = (unsigned int) i;
            }
        }
        // world cell bounds: warp shuffles, then one atomic per warp
        int mnx = r.x0, mny = r.y0, mxx = r.x1, mxy = r.y1;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, d));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
        }
        unsigned int total;
        unsigned int excl = block_exclusive_scan<BP_BLOCK>(cnt, &total);
        unsigned int base = tile_exclusive_prefix(status, tile, total);
        if (valid) {
            w.cell_cnt[i] = cnt;
            w.cell_off[i] = base + excl;
        }
        if ((int) tile == ntiles - 1 && threadIdx.x == 0) w.ctr->n_entries = base + total;
        // (after the look-back: the tile's prefix is published before this warp waits for a look at the bounds)
        if (lane_id() == 0 && mnx <= mxx) {
            // the bounds only ever widen, so a plain look first is safe: after the first tiles almost no warp
            // needs its atomics ([measured] 1 M bodies: 131 k atomics on four addresses otherwise)
            // (a stale L1 copy only makes a warp take an atomic it did not need)
            const StepCountersDev *vc = w.ctr;
            if (mnx < vc->cell_min_x) atomicMin(&w.ctr->cell_min_x, mnx);
            if (mny < vc->cell_min_y) atomicMin(&w.ctr->cell_min_y, mny);
            if (mxx > vc->cell_max_x) atomicMax(&w.ctr->cell_max_x, mxx);
            if (mxy > vc->cell_max_y) atomicMax(&w.ctr->cell_max_y, mxy);
        }
    }
}

// Dense key geometry (grid_w/h, key bits, radix passes) from the world cell bounds the prepass left behind.  Every
// block of k_emit_entries evaluates it (a few dozen integer operations) instead of waiting for a one-thread kernel;
// `publish` (one thread of the grid) also stores it for the kernels that follow.  Returns false when nothing may be
// emitted this step.
__device__ __forceinline__ bool plan_keys(const DevWorld &w, bool publish, unsigned long long *out_gw, unsigned long long *out_cells,
                                          unsigned int *out_passes, unsigned int *out_bits) {
    StepCountersDev *c = w.ctr;
    unsigned int flags = 0u;
    unsigned int n_entries = c->n_entries, n_large = c->n_large;
    if (n_entries > w.cap_entries) { flags |= 1u; n_entries = w.cap_entries; }
    if (n_large > w.cap_large) { flags |= 1u; n_large = w.cap_large; }
    unsigned long long gw = 1, gh = 1;
    const int mnx = c->cell_min_x, mny = c->cell_min_y, mxx = c->cell_max_x, mxy = c->cell_max_y;
    if (mnx <= mxx) {
        gw = (unsigned long long) ((long long) mxx - mnx + 1);
        gh = (unsigned long long) ((long long) mxy - mny + 1);
    }
    const unsigned long long groups = (w.group != nullptr && w.n_groups > 1) ? (unsigned long long) w.n_groups : 1ull;
    // The key is DENSE -- group * (W*H) + (y - ymin) * W + (x - xmin), so bodies of different sub-worlds never share a
    // run -- and 64 bits wide: the reference's hash is keyed by (x, y) and has no extent limit
    // (src/broad_phase.c:99-127), so a body thrown far away must not be able to refuse the step.  Worlds whose key
    // fits 32 bits (every pile) never touch the high words; the radix sort runs ceil(bits / 8) digit passes.
    bool ok = true;
    unsigned long long cells = 1;
    if (gh > (1ull << 62) / gw || gw * gh > (1ull << 62) / groups) {
        ok = false;                         // > 2^62 cells: both extents beyond 2^31 cells, i.e. non-finite coordinates
        gw = gh = 1;
    } else cells = gw * gh;
    unsigned int bits = 1;
    while (bits < 63 && (1ull << bits) < cells * groups) ++bits;
    unsigned int passes = (bits + 7) / 8;
    if ((int) passes > (w.sort_launch > 0 ? w.sort_launch : 8)) {
        // the host launched fewer digit passes than these keys need (it sizes them from the previous step's key width
        // plus one digit, and launches all eight after any setter moved a body): the world's extent grew more than
        // 256-fold within one step of free motion.  Refuse loudly, emit nothing; the next step re-measures.
        ok = false;
        passes = 0;
    }
    if (!ok) { flags |= 16u; n_entries = 0; }
    if (publish) {
        if (flags) atomicOr(&c->overflow, flags);
        c->n_entries = n_entries;
        c->n_large = n_large;
        c->grid_w = (unsigned int) (gw > 0xffffffffull ? 0xffffffffull : gw);
        c->grid_h = (unsigned int) (gh > 0xffffffffull ? 0xffffffffull : gh);
        c->key_bits = bits;
        c->sort_passes = passes;
    }
    *out_gw = gw;
    *out_cells = cells;
    *out_passes = passes;
    *out_bits = bits;
    return ok;
}

struct KeyPlan {
    unsigned long long gw, cells;
    unsigned int passes, wide;     // wide: the key needs more than 32 bits (the high words are live)
    int minx, miny;
};

__device__ __forceinline__ void emit_entry(const DevWorld &w, unsigned int pos, int x, int y, const CellRect &r,
                                           unsigned int body, unsigned int massbit, const KeyPlan &kp, unsigned int (*hist)[256]) {
    if (pos >= w.cap_entries) return;
    unsigned long long key = (unsigned long long) ((unsigned int) y - (unsigned int) kp.miny) * kp.gw +
                             (unsigned long long) ((unsigned int) x - (unsigned int) kp.minx);
    if (w.group != nullptr) key += (unsigned long long) w.group[body] * kp.cells;
    unsigned int v = body | massbit | (x == r.x0 ? ENT_X0 : 0u) | (y == r.y0 ? ENT_Y0 : 0u);
    w.keys[0][pos] = (unsigned int) key;
    if (kp.wide) w.keys_hi[0][pos] = (unsigned int) (key >> 32);
    w.vals[0][pos] = v;
    // the digit histograms of the radix sort are counted here, while the key is at hand (no separate pass over the keys)
    for (unsigned int p = 0; p < kp.passes; ++p) atomicAdd(&hist[p][(unsigned int) (key >> (8u * p)) & 255u], 1u);
}

// the last `large_blocks` blocks of the grid expand the bodies that cover more than FX_LARGE_CELLS cells (ground
// planes spanning thousands of cells), one body per block at a time; the others take one ordinary body per thread
__global__ void __launch_bounds__(BP_BLOCK) k_emit_entries(DevWorld w, int large_blocks, unsigned int *ghist) {
    __shared__ KeyPlan s_kp;
    __shared__ unsigned int s_ok;
    __shared__ unsigned int s_h[FX_SORT_MAX_PASSES][256];
    for (int t = threadIdx.x; t < FX_SORT_MAX_PASSES * 256; t += blockDim.x) (&s_h[0][0])[t] = 0u;
    if (threadIdx.x == 0) {
        unsigned int bits;
        s_ok = plan_keys(w, blockIdx.x == 0, &s_kp.gw, &s_kp.cells, &s_kp.passes, &bits) ? 1u : 0u;
        s_kp.wide = bits > 32u ? 1u : 0u;
        s_kp.minx = w.ctr->cell_min_x;
        s_kp.miny = w.ctr->cell_min_y;
    }
    __syncthreads();
    if (!s_ok) return;
    const KeyPlan kp = s_kp;
    const int body_blocks = (int) gridDim.x - large_blocks;
    if ((int) blockIdx.x >= body_blocks) {
        const unsigned int nl = min(w.ctr->n_large, w.cap_large);
        for (unsigned int l = blockIdx.x - body_blocks; l < nl; l += (unsigned int) large_blocks) {
            unsigned int i = w.large_list[l];
            CellRect r = body_rect(w.aabb[i], w.inv_cell);
            unsigned int rw = (unsigned int) (r.x1 - r.x0 + 1);
            unsigned int cnt = w.cell_cnt[i], pos = w.cell_off[i];
            unsigned int massbit = (w.vel[i].w > 0.0f && (int) i < w.n_owned) ? ENT_MASS : 0u;
            for (unsigned int t = threadIdx.x; t < cnt; t += blockDim.x)
                emit_entry(w, pos + t, r.x0 + (int) (t % rw), r.y0 + (int) (t / rw), r, i, massbit, kp, s_h);
        }
    } else
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.n; i += body_blocks * blockDim.x) {
        unsigned int cnt = w.cell_cnt[i];
        if (cnt == 0u || cnt > FX_LARGE_CELLS) continue;
        CellRect r = body_rect(w.aabb[i], w.inv_cell);
        unsigned int pos = w.cell_off[i];
        // ghost rows never count as 'has mass': ghost/ghost and ghost/static pairs belong to the owner
        unsigned int massbit = (w.vel[i].w > 0.0f && i < w.n_owned) ? ENT_MASS : 0u;
        for (int y = r.y0; y <= r.y1; ++y)
            for (int x = r.x0; x <= r.x1; ++x) emit_entry(w, pos++, x, y, r, (unsigned int) i, massbit, kp, s_h);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (int) kp.passes * 256; t += blockDim.x) {
        const unsigned int v = (&s_h[0][0])[t];
        if (v) atomicAdd(&ghist[t], v);
    }
}

// ---------------------------------------------------------------------------------------------
// radix sort of (key, value) entries
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BP_BLOCK) k_sort_hist(SortJob job) {
    __shared__ unsigned int s_h[4][256];     // 32-bit keys only (the candidate ordering of the parity mode)
    const unsigned int n = *job.n;
    const unsigned int passes = min(job.passes_dev ? *job.passes_dev : job.passes, 4u);
    for (int t = threadIdx.x; t < 4 * 256; t += blockDim.x) (&s_h[0][0])[t] = 0u;
    __syncthreads();
    const unsigned int *keys = job.keys[0];
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        unsigned int k = keys[e];
        for (unsigned int p = 0; p < passes; ++p) atomicAdd(&s_h[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 4 * 256; t += blockDim.x) {
        unsigned int v = (&s_h[0][0])[t];
        if (v) atomicAdd(&job.ghist[t], v);
    }
}

// per-digit tile status: [31:30] flag, [29:0] count
#define DS_AGG (1u << 30)
#define DS_PFX (2u << 30)
#define DS_VAL ((1u << 30) - 1u)

// WIDE: keys of more than 32 bits (the host launches this build only when it launches more than four digit passes;
// the narrow build keeps the high words out of its registers and its shared memory)
//
// One digit pass, "onesweep" style.  A tile of 2,048 consecutive entries: (1) every warp ranks its 256 entries by
// digit (match_any, stable in entry order); (2) thread d owns digit d: per-warp bases, the tile's count of d, its
// aggregate published at once for the tiles behind; (3) the entries are written to shared memory IN DIGIT ORDER
// (tile-local offsets from a block scan over the 256 counts) while (4) thread d resolves its digit's global base
// by decoupled look-back; (5) the block streams the reordered tile out: consecutive threads hold consecutive entries
// of a digit run, so the stores are whole sectors of consecutive addresses instead of one scattered 4-byte store per
// entry and array ([measured] at 2.3 M entries: 55 us per pass before, scattering straight from registers).
template <bool WIDE>
__global__ void __launch_bounds__(BP_BLOCK) k_sort_pass(SortJob job, int pass) {
    __shared__ unsigned int s_wc[BP_BLOCK / 32][256];   // per-warp digit counters -> exclusive bases
    __shared__ unsigned int s_toff[256];                // first tile-local slot of each digit
    __shared__ unsigned int s_adj[256];                 // global position of a digit's entry = s_adj[digit] + tile-local slot
    __shared__ unsigned int s_key[SORT_TILE], s_val[SORT_TILE];
    __shared__ unsigned int s_hi[WIDE ? SORT_TILE : 1];
    const unsigned int n = *job.n;
    const unsigned int passes = job.passes_dev ? *job.passes_dev : job.passes;
    if ((unsigned int) pass >= passes) return;
    // wide keys (> 32 bits): the high words travel with the entries; passes 4..7 take their digit from them
    const bool wide = WIDE && job.bits_dev != nullptr && *job.bits_dev > 32u;
    // (selected, not indexed: indexing the parameter struct by `pass` would have it copied to local memory)
    const bool odd = (pass & 1) != 0;
    const unsigned int *kin = odd ? job.keys[1] : job.keys[0], *vin = odd ? job.vals[1] : job.vals[0];
    unsigned int *kout = odd ? job.keys[0] : job.keys[1], *vout = odd ? job.vals[0] : job.vals[1];
    const unsigned int *hin = wide ? (odd ? job.keys_hi[1] : job.keys_hi[0]) : nullptr;
    unsigned int *hout = wide ? (odd ? job.keys_hi[0] : job.keys_hi[1]) : nullptr;
    const bool hi_digit = WIDE && pass >= 4;
    unsigned int *dstatus = job.dstatus + (size_t) pass * job.dstatus_stride;
    unsigned int *ticket = job.tickets + pass;
    const unsigned int shift = 8u * (unsigned int) (pass & 3);
    const unsigned int ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    const unsigned int warp = threadIdx.x >> 5, lane = lane_id();

    // exclusive scan of the global digit histogram of this pass (256 bins = 256 threads)
    unsigned int tot;
    const unsigned int bin_base = block_exclusive_scan<BP_BLOCK>(job.ghist[pass * 256 + threadIdx.x], &tot);

    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if (tile >= ntiles) return;
        for (int t = threadIdx.x; t < (BP_BLOCK / 32) * 256; t += blockDim.x) (&s_wc[0][0])[t] = 0u;
        __syncthreads();
        unsigned int key[SORT_ITEMS], khi[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
        const unsigned int tile_base = tile * SORT_TILE;
        const unsigned int tile_n = min((unsigned int) SORT_TILE, n - tile_base);
        const unsigned int base = tile_base + warp * (32 * SORT_ITEMS);
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            unsigned int e = base + j * 32 + lane;
            bool ok = e < n;
            key[j] = ok ? kin[e] : 0xffffffffu;
            khi[j] = (WIDE && ok && wide) ? hin[e] : 0u;
            val[j] = ok ? vin[e] : 0u;
        }
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            bool ok = base + j * 32 + lane < n;
            unsigned int d = ((hi_digit ? khi[j] : key[j]) >> shift) & 255u;
            unsigned int m = ok ? d : (256u + lane);
            unsigned int peers = __match_any_sync(0xffffffffu, m);
            unsigned int before = __popc(peers & ((1u << lane) - 1u));
            unsigned int pre = ok ? s_wc[warp][d] : 0u;
            __syncwarp();
            if (ok && before == 0u) s_wc[warp][d] = pre + __popc(peers);
            __syncwarp();
            rank[j] = pre + before;
        }
        __syncthreads();
        // thread d owns digit d: turn the per-warp counts into exclusive warp bases, get the tile total
        const unsigned int d = threadIdx.x;
        unsigned int run = 0u;
#pragma unroll
        for (int wv = 0; wv < BP_BLOCK / 32; ++wv) {
            unsigned int c = s_wc[wv][d];
            s_wc[wv][d] = run;
            run += c;
        }
        // chained scan over tiles, one chain per digit: the aggregate goes out before anything else happens here
        unsigned int *st = dstatus + (size_t) tile * 256u + d;
        if (tile > 0) *reinterpret_cast<volatile unsigned int *>(st) = DS_AGG | run;
        unsigned int tile_tot;
        const unsigned int toff = block_exclusive_scan<BP_BLOCK>(run, &tile_tot);      // (syncs: the warp bases are visible after it)
        s_toff[d] = toff;
        __syncthreads();
        // the tile in digit order, in shared memory
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            if (base + j * 32 + lane < n) {
                unsigned int dd = ((hi_digit ? khi[j] : key[j]) >> shift) & 255u;
                unsigned int slot = s_toff[dd] + s_wc[warp][dd] + rank[j];
                s_key[slot] = key[j];
                s_val[slot] = val[j];
                if (WIDE) s_hi[slot] = khi[j];
            }
        }
        // look-back of digit d (the tiles ahead published their aggregates before reordering, as this one did)
        unsigned int excl = 0u;
        if (tile > 0) {
            int t = (int) tile - 1;
            for (;;) {
                unsigned int sv;
                const volatile unsigned int *ps = dstatus + (size_t) t * 256u + d;
                do { sv = *ps; } while ((sv >> 30) == 0u);
                excl += (sv & DS_VAL);
                if ((sv >> 30) == 2u) break;
                --t;
            }
        }
        *reinterpret_cast<volatile unsigned int *>(st) = DS_PFX | (excl + run);
        s_adj[d] = bin_base + excl - toff;
        __syncthreads();
        // stream the reordered tile out
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const unsigned int slot = j * BP_BLOCK + threadIdx.x;
            if (slot < tile_n) {
                const unsigned int k = s_key[slot];
                const unsigned int h = WIDE ? s_hi[slot] : 0u;
                const unsigned int dd = ((hi_digit ? h : k) >> shift) & 255u;
                const unsigned int pos = s_adj[dd] + slot;
                kout[pos] = k;
                if (WIDE && wide) hout[pos] = h;
                vout[pos] = s_val[slot];
            }
        }
        // (the next tile's first barrier -- after the counter reset -- separates these reads from its writes to s_key /
        // s_val: those happen two barriers later; s_toff / s_adj likewise)
    }
}

static void launch_sort(const SortJob &job, long long cap, int max_passes, int max_blocks, cudaStream_t s, bool with_hist = true) {
    long long gh = (cap + BP_BLOCK * 8 - 1) / (BP_BLOCK * 8), gp = (cap + SORT_TILE - 1) / SORT_TILE;
    int gridh = (int) (gh < 1 ? 1 : (gh > max_blocks ? max_blocks : gh));
    int gridp = (int) (gp < 1 ? 1 : (gp > max_blocks ? max_blocks : gp));
    if (with_hist) { k_sort_hist<<<gridh, BP_BLOCK, 0, s>>>(job); ++g_fx_launches; }
    for (int p = 0; p < max_passes; ++p) {
        if (max_passes > 4) k_sort_pass<true><<<gridp, BP_BLOCK, 0, s>>>(job, p);
        else k_sort_pass<false><<<gridp, BP_BLOCK, 0, s>>>(job, p);
        ++g_fx_launches;
    }
}

// ---- lexicographic (i, j) ordering of the candidate list: serial / parity mode only -------------
__global__ void k_lex_init(DevWorld w, unsigned int *keys, unsigned int *vals) {
    const unsigned int n = w.ctr->n_cand;
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        keys[c] = (unsigned int) w.cand[c].y;
        vals[c] = c;
    }
}
__global__ void k_lex_mid(DevWorld w, const unsigned int *perm, unsigned int *keys) {
    const unsigned int n = w.ctr->n_cand;
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
        keys[c] = (unsigned int) w.cand[perm[c]].x;
}
__global__ void k_lex_gather(DevWorld w, const unsigned int *perm, int2 *tmp) {
    const unsigned int n = w.ctr->n_cand;
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) tmp[c] = w.cand[perm[c]];
}
__global__ void k_lex_copy(DevWorld w, const int2 *tmp) {
    const unsigned int n = w.ctr->n_cand;
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) w.cand[c] = tmp[c];
}

// ---------------------------------------------------------------------------------------------
// pair emission
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_ok(unsigned int a, unsigned int b) {
    unsigned int u = a | b;   // canonical cell: min-x column of one AND min-y row of one; some mass
    return (u & ENT_MASS) && (u & ENT_X0) && (u & ENT_Y0);
}

// ([measured] round 2: staging the tile and the 128 entries behind it in shared memory, so that both walks read shared
// memory, left the small-world build where it was (20.3 vs 20.5 us at 65k bodies) and made the large one SLOWER -- 107 vs
// 87 us at 2.4 M entries, 56 M warp instructions instead of 36 M: the runs are two or three entries long and L1 already
// serves the second walk.  Removed.)
// ITEMS consecutive entries per thread: large worlds take 4 (tiles of 1,024 entries: the look-back chain is 4x
// shorter), small ones 1 (more blocks in flight matter more there).
template <int ITEMS>
__global__ void __launch_bounds__(BP_BLOCK) k_emit_pairs(DevWorld w, unsigned long long *status, unsigned int *ticket) {
    const unsigned int n = w.ctr->n_entries;
    const unsigned int sel = w.ctr->sort_passes & 1u;
    const unsigned int *keys = w.keys[sel], *vals = w.vals[sel];
    const unsigned int *khi = w.ctr->key_bits > 32u ? w.keys_hi[sel] : nullptr;      // wide keys: a run shares both words
    const unsigned int tile_size = BP_BLOCK * ITEMS;
    const unsigned int ntiles = (n + tile_size - 1) / tile_size;
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) w.ctr->n_cand = 0;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if (tile >= ntiles) return;
        const unsigned int first = tile * tile_size + threadIdx.x * ITEMS;
        unsigned int cnt[ITEMS], sum = 0u;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const unsigned int e = first + j;
            cnt[j] = 0u;
            if (e < n) {
                const unsigned int k = keys[e], v = vals[e], kh = khi ? khi[e] : 0u;
                for (unsigned int f = e + 1; f < n && keys[f] == k && (!khi || khi[f] == kh); ++f) cnt[j] += pair_ok(v, vals[f]) ? 1u : 0u;
            }
            sum += cnt[j];
        }
        unsigned int total;
        unsigned int excl = block_exclusive_scan<BP_BLOCK>(sum, &total);
        unsigned int base = tile_exclusive_prefix(status, tile, total);
        unsigned int pos = base + excl;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (!cnt[j]) continue;
            const unsigned int e = first + j;
            const unsigned int k = keys[e], v = vals[e], kh = khi ? khi[e] : 0u;
            for (unsigned int f = e + 1; f < n && keys[f] == k && (!khi || khi[f] == kh); ++f) {
                unsigned int vf = vals[f];
                if (pair_ok(v, vf)) {
                    // entries of one cell are in ascending body order (stable sort of an ascending
                    // emission), so ENT_BODY(v) < ENT_BODY(vf): the reference's i < j (world.c:333)
                    if (pos < w.cap_cand) w.cand[pos] = make_int2((int) ENT_BODY(v), (int) ENT_BODY(vf));
                    ++pos;
                }
            }
        }
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            unsigned int c = base + total;
            if (c > w.cap_cand) { atomicOr(&w.ctr->overflow, 2u); c = w.cap_cand; }
            w.ctr->n_cand = c;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch surface
// ---------------------------------------------------------------------------------------------
static inline int grid_for(long long n, int block, int max_blocks) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int) g;
}

void fx_launch_integrate_velocity(const DevWorld &w, const FxLaunchCtx &cx, bool store_force) {
    if (w.n == 0) return;
    k_integrate_velocity<<<grid_for(w.n, BP_BLOCK, cx.max_blocks), BP_BLOCK, 0, cx.stream>>>(w, store_force ? 1 : 0); ++g_fx_launches;
}

void fx_launch_broadphase(const DevWorld &w, const FxLaunchCtx &cx, bool fuse_velocity) {
    cudaStream_t s = cx.stream;
    const int G = cx.max_blocks;
    k_body_prepass<<<grid_for(w.n, BP_BLOCK, G), BP_BLOCK, 0, s>>>(w, fuse_velocity ? 1 : 0, cx.status[0], cx.tickets + 0); ++g_fx_launches;
    k_emit_entries<<<grid_for(w.n, BP_BLOCK, G) + 64, BP_BLOCK, 0, s>>>(w, 64, cx.sort_hist); ++g_fx_launches;
    SortJob job;
    job.keys[0] = w.keys[0]; job.keys[1] = w.keys[1]; job.vals[0] = w.vals[0]; job.vals[1] = w.vals[1];
    job.keys_hi[0] = w.keys_hi[0]; job.keys_hi[1] = w.keys_hi[1]; job.bits_dev = &w.ctr->key_bits;
    job.n = &w.ctr->n_entries; job.passes_dev = &w.ctr->sort_passes; job.passes = FX_SORT_MAX_PASSES;
    job.ghist = cx.sort_hist; job.dstatus = cx.digit_status; job.dstatus_stride = cx.digit_status_stride;
    job.tickets = cx.tickets + 8;
    launch_sort(job, (long long) w.cap_entries, w.sort_launch > 0 && w.sort_launch < FX_SORT_MAX_PASSES ? w.sort_launch : FX_SORT_MAX_PASSES, G, s,
                false /* histograms: k_emit_entries */);
    if (w.sol_mult == 1u) k_emit_pairs<4><<<grid_for((long long) w.cap_entries, BP_BLOCK * 4, G), BP_BLOCK, 0, s>>>(w, cx.status[1], cx.tickets + 1);   // large world
    else k_emit_pairs<1><<<grid_for((long long) w.cap_entries, BP_BLOCK, G), BP_BLOCK, 0, s>>>(w, cx.status[1], cx.tickets + 1);
    ++g_fx_launches;
}

// Serial mode: reorder the candidate list into the reference's visiting order (i ascending, then j
// ascending; src/broad_phase.c:167-178 + src/world.c:438-443).  Two stable LSD sorts of a
// permutation (by j, then by i).  lex.* are scratch arrays of cap_cand elements.
void fx_launch_lexsort_candidates(const DevWorld &w, const FxLaunchCtx &cx, const FxLexScratch &lex, int body_bits) {
    cudaStream_t s = cx.stream;
    const int G = cx.max_blocks, passes = (body_bits + 7) / 8;
    const int g = grid_for((long long) w.cap_cand, BP_BLOCK, G);
    for (int round = 0; round < 2; ++round) {
        if (round == 0) { k_lex_init<<<g, BP_BLOCK, 0, s>>>(w, lex.keys[0], lex.vals[0]); ++g_fx_launches; }
        SortJob job;
        // after `passes` passes the data sits in buffer (passes & 1); round 1 starts from there
        int src = (round == 0) ? 0 : (passes & 1);
        if (round == 1) { k_lex_mid<<<g, BP_BLOCK, 0, s>>>(w, lex.vals[src], lex.keys[src]); ++g_fx_launches; }
        job.keys[0] = lex.keys[src]; job.keys[1] = lex.keys[src ^ 1];
        job.vals[0] = lex.vals[src]; job.vals[1] = lex.vals[src ^ 1];
        job.keys_hi[0] = job.keys_hi[1] = nullptr; job.bits_dev = nullptr;
        job.n = &w.ctr->n_cand; job.passes_dev = nullptr; job.passes = (unsigned int) passes;
        job.ghist = lex.ghist + round * 1024; job.dstatus = lex.dstatus + (size_t) round * 4 * lex.dstatus_stride;
        job.dstatus_stride = lex.dstatus_stride; job.tickets = lex.tickets + round * 4;
        launch_sort(job, (long long) w.cap_cand, passes, G, s);
    }
    const int fin = 0;   // round 0 ends in buffer (passes & 1), round 1 starts there and ends in buffer 0
    k_lex_gather<<<g, BP_BLOCK, 0, s>>>(w, lex.vals[fin], lex.tmp); ++g_fx_launches;
    k_lex_copy<<<g, BP_BLOCK, 0, s>>>(w, lex.tmp); ++g_fx_launches;
}
