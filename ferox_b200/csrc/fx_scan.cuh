// fx_scan.cuh -- block scans and the single-pass "ordered tile" machinery (decoupled look-back).
//
// Every variable-length stage of the step (cell-entry emission, pair emission, manifold compaction,
// stale-entry carry-over) appends a data-dependent number of items per tile and needs the items in
// a DETERMINISTIC order (tile order) without a host round trip.  Tiles are claimed through an atomic
// ticket, so a tile only ever waits on tiles whose blocks are already running: the look-back below
// cannot deadlock regardless of how many blocks are resident.
#pragma once
#include <cuda_runtime.h>

#include "fx_math.cuh"

#define FX_WARP 32

FX_D unsigned int lane_id() { return threadIdx.x & 31u; }

FX_D unsigned int warp_inclusive_scan(unsigned int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, v, d);
        if ((int) lane_id() >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across a block of BLOCK threads (BLOCK <= 1024).
// Returns the exclusive prefix; *total receives the block sum (valid in every thread).
template <int BLOCK>
FX_D unsigned int block_exclusive_scan(unsigned int v, unsigned int *total) {
    __shared__ unsigned int s_warp[BLOCK / 32];
    __shared__ unsigned int s_total;
    unsigned int inc = warp_inclusive_scan(v);
    unsigned int w = threadIdx.x >> 5;
    if (lane_id() == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned int x = (lane_id() < BLOCK / 32) ? s_warp[lane_id()] : 0u;
        unsigned int xi = warp_inclusive_scan(x);
        if (lane_id() < BLOCK / 32) s_warp[lane_id()] = xi - x;
        if (lane_id() == 31) s_total = xi;
    }
    __syncthreads();
    unsigned int r = s_warp[w] + inc - v;
    *total = s_total;
    __syncthreads();
    return r;
}

// ---- tile status words: [63:62] flag (0 = empty, 1 = tile aggregate, 2 = inclusive prefix) ----
#define FX_TILE_AGG (1ull << 62)
#define FX_TILE_PFX (2ull << 62)
#define FX_TILE_VAL ((1ull << 62) - 1ull)

FX_D unsigned long long ld_status(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
FX_D void st_status(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// Claim the next tile of a persistent, ordered sweep.  Block-uniform result.
FX_D unsigned int claim_tile(unsigned int *ticket) {
    __shared__ unsigned int s_tile;
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    return s_tile;
}

// Exclusive prefix of `tile_total` over all tiles with a smaller index.  Called by the whole block
// with a block-uniform tile_total.  The WHOLE BLOCK looks back, blockDim.x predecessors per hop: it
// adds up tile aggregates until it meets a tile that already knows its inclusive prefix.  With
// hundreds of tiles in flight the prefix frontier advances one window per L2 round trip, so the
// window width sets the speed of every ordered sweep ([measured] narrow phase, 4,216 tiles of the
// 65k-body config: 1 thread 186 us, one warp 177 us; see DESIGN.md for the block-wide figure).
// A tile may publish its aggregate as soon as it knows it and do other work before it looks back (the tiles behind it
// then never wait for that work); tile_exclusive_prefix stores the same word again, which is harmless.
FX_D void tile_publish_aggregate(unsigned long long *status, unsigned int tile, unsigned int tile_total) {
    if (tile > 0 && threadIdx.x == 0) st_status(&status[tile], FX_TILE_AGG | (unsigned long long) tile_total);
}

FX_D unsigned int tile_exclusive_prefix(unsigned long long *status, unsigned int tile,
                                        unsigned int tile_total) {
    __shared__ unsigned long long s_part[32];
    __shared__ int s_near[32];
    __shared__ unsigned int s_prefix;
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = (blockDim.x + 31u) >> 5;
    unsigned long long excl = 0ull;
    if (tile > 0) {
        if (tid == 0) st_status(&status[tile], FX_TILE_AGG | (unsigned long long) tile_total);
        int base = (int) tile - 1;                      // nearest predecessor of this window
        for (;;) {
            const int idx = base - (int) tid;
            unsigned long long s = FX_TILE_PFX;         // before the first tile: prefix 0
            if (idx >= 0) {
                do { s = ld_status(&status[idx]); } while ((s >> 62) == 0ull);
            }
            // nearest predecessor (smallest tid) that already holds an inclusive prefix
            int near = ((s >> 62) == 2ull) ? (int) tid : 0x7fffffff;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) near = min(near, __shfl_xor_sync(0xffffffffu, near, d));
            if (lane == 0) s_near[warp] = near;
            __syncthreads();
            int stop = 0x7fffffff;
            for (unsigned int k = 0; k < nwarps; ++k) stop = min(stop, s_near[k]);
            unsigned long long v = ((int) tid <= stop) ? (s & FX_TILE_VAL) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) s_part[warp] = v;
            __syncthreads();
            for (unsigned int k = 0; k < nwarps; ++k) excl += s_part[k];
            __syncthreads();
            if (stop != 0x7fffffff) break;
            base -= (int) blockDim.x;
        }
    }
    if (tid == 0) {
        st_status(&status[tile], FX_TILE_PFX | (excl + (unsigned long long) tile_total));
        s_prefix = (unsigned int) excl;
    }
    __syncthreads();
    return s_prefix;
}
