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
// with a block-uniform tile_total; one thread walks back over the predecessors.
FX_D unsigned int tile_exclusive_prefix(unsigned long long *status, unsigned int tile,
                                        unsigned int tile_total) {
    __shared__ unsigned int s_prefix;
    if (threadIdx.x == 0) {
        unsigned long long excl = 0ull;
        if (tile > 0) {
            st_status(&status[tile], FX_TILE_AGG | (unsigned long long) tile_total);
            int t = (int) tile - 1;
            for (;;) {
                unsigned long long s;
                do { s = ld_status(&status[t]); } while ((s >> 62) == 0ull);
                excl += (s & FX_TILE_VAL);
                if ((s >> 62) == 2ull) break;
                --t;
            }
        }
        st_status(&status[tile], FX_TILE_PFX | (excl + (unsigned long long) tile_total));
        s_prefix = (unsigned int) excl;
    }
    __syncthreads();
    return s_prefix;
}
