// fx_strip.cu -- strip domain decomposition of ONE large world over several GPUs: device side.
//
// The reference has no counterpart (it is single-threaded); this is the
// "single large world" row of SURVEY.md section 8e / BASELINE config 5.  A rank owns the bodies added
// to its strip.  Before a step every strip packs the owned bodies that reach within `margin` of its
// LEFT cut (border bodies) into a fixed-size message; the left neighbour appends them as GHOST rows
// behind its own bodies, so its broadphase / narrow phase / contact cache / colouring run unchanged
// on owned + ghost rows.  Every pair across a cut is thus found and solved on exactly ONE rank (the
// strip on the left of the cut), with one set of accumulated impulses.  After each solver stage
// (warm start, then every Gauss-Seidel sweep) the two sides trade one message each:
//   left  -> right : the velocity change the stage gave each ghost      (the owner adds it)
//   right -> left  : the owner's own new velocity of each border body   (the ghost becomes that + the
//                                                                         change computed locally)
// so both copies end the stage with the same bits and momentum is exchanged exactly.  The coupling
// across a cut lags one stage (the owner solves its other contacts without this stage's cross
// impulse), i.e. block Gauss-Seidel within a strip, Jacobi across cuts.  Plain Jacobi coupling
// over-corrects a body that is pushed from both sides in the same stage ([measured] a 3,000-body pile
// cut at x = 30 diverged at the impact phase), so bodies that actually have a manifold across the cut
// are MASS-SPLIT: both solvers see them with twice the inverse mass / inertia and each applies half of
// the velocity change it computed (the owner learns which of its border bodies are affected from a
// flag message sent right after the narrow phase).
// Ghost/ghost and ghost/static pairs are never emitted (ghost rows carry no "has mass" bit in the
// cell entries: they belong to the owner), ghost rows are not position-integrated and vanish at the
// end of the step.  Messages have a 16-byte header {count} and a fixed capacity: nothing about them
// needs the host.
#include "fx_device.h"
#include "fx_kernels.h"
#include "fx_scan.cuh"
extern long long g_fx_launches;

#define ST_BLOCK 256

// ordered compaction of the border bodies of one side into `msg`; remembers their indices
// Also flags (overflow bit 128) an owned body whose position lies outside [stray_lo, stray_hi] = the strip widened by
// the margin: ownership is static, such a body has lost its cross-cut contacts.
__global__ void __launch_bounds__(ST_BLOCK) k_strip_pack(DevWorld w, int side, float cut, float margin, unsigned char *msg,
                                                         unsigned int *border_idx, unsigned int capacity, float stray_lo, float stray_hi,
                                                         unsigned long long *status, unsigned int *ticket) {
    GhostRow *rows = reinterpret_cast<GhostRow *>(msg + FX_MSG_HEADER);
    const int ntiles = (w.n_owned + ST_BLOCK - 1) / ST_BLOCK;
    if (ntiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<unsigned int *>(msg) = 0u;
    for (;;) {
        unsigned int tile = claim_tile(ticket);
        if ((int) tile >= ntiles) return;
        const int i = (int) tile * ST_BLOCK + (int) threadIdx.x;
        bool take = false;
        if (i < w.n_owned) {
            const int type = __float_as_int(w.mprop[i].w) & 0xff;
            const float4 a = w.aabb[i];
            // static bodies are replicated on every rank by construction, never exchanged
            if (type != FX_BODY_STATIC) {
                take = (side == 0) ? (a.x < cut + margin) : (a.x + a.z > cut - margin);
                const float px = w.posrot[i].x;          // the position, not the AABB (a state of its own, App. A.9)
                if (px < stray_lo || px > stray_hi) atomicOr(&w.ctr->overflow, 128u);
            }
        }
        unsigned int total;
        const unsigned int excl = block_exclusive_scan<ST_BLOCK>(take ? 1u : 0u, &total);
        const unsigned int base = tile_exclusive_prefix(status, tile, total);
        if (take) {
            const unsigned int r = base + excl;
            if (r < capacity) {
                GhostRow g;
                g.posrot = w.posrot[i]; g.vel = w.vel[i]; g.mprop = w.mprop[i]; g.force = w.force[i]; g.aabb = w.aabb[i];
                g.angle = w.angle[i];
                const int2 sh = w.shape[i];
                g.shape_idx = sh.x; g.shape_type = sh.y;
                g.uid = w.uid[i];
                rows[r] = g;
                border_idx[r] = (unsigned int) i;
            }
        }
        if ((int) tile == ntiles - 1 && threadIdx.x == 0) {
            unsigned int c = base + total;
            if (c > capacity) { atomicOr(&w.ctr->overflow, 32u); c = capacity; }
            *reinterpret_cast<unsigned int *>(msg) = c;
        }
    }
}

// forget where last step's ghosts were (uid -> row map), before the new ones arrive
__global__ void k_strip_clear_ghost_map(DevWorld w, int first_row, int rows) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const int i = first_row + r;
        if (w.shape[i].y != -7) w.idx_of_uid[w.uid[i]] = -1;     // -7 marks an inert padding row
    }
}

// append the neighbour's border bodies as ghost rows [first_row, first_row + capacity); rows beyond
// the message's count become inert (no cells, no mass, no shape)
__global__ void k_strip_unpack(DevWorld w, const unsigned char *msg, int first_row, unsigned int capacity, unsigned int *ghost_cnt) {
    const unsigned int count = min(*reinterpret_cast<const unsigned int *>(msg), capacity);
    const GhostRow *rows = reinterpret_cast<const GhostRow *>(msg + FX_MSG_HEADER);
    if (blockIdx.x == 0 && threadIdx.x == 0) *ghost_cnt = count;
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < capacity; r += gridDim.x * blockDim.x) {
        const int i = first_row + (int) r;
        if (r < count) {
            const GhostRow g = rows[r];
            w.posrot[i] = g.posrot; w.vel[i] = g.vel; w.mprop[i] = g.mprop; w.force[i] = g.force; w.aabb[i] = g.aabb;
            w.angle[i] = g.angle;
            w.shape[i] = make_int2(g.shape_idx, g.shape_type);
            w.uid[i] = g.uid;
            w.idx_of_uid[g.uid] = i;
        } else {
            w.posrot[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            w.vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            w.mprop[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(FX_BODY_STATIC));
            w.force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            w.aabb[i] = make_float4(0.f, 0.f, -1.0e30f, -1.0e30f);   // x1 < x0: covers no cell
            w.angle[i] = 0.f;
            w.shape[i] = make_int2(-1, -7);
            w.uid[i] = 0u;
        }
    }
}

// ---- which ghosts take part in a manifold this step (their owner must mass-split them too) ----
__global__ void k_strip_mark_used(DevWorld w, float4 *flags_msg, unsigned int capacity, int phase) {
    if (phase == 0) {
        for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < capacity; r += gridDim.x * blockDim.x)
            flags_msg[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const ManifoldSet man = w.man[*w.cur_flip];
    const unsigned int M = w.ctr->n_manifolds;
    for (unsigned int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const int2 b = man.body[m];
        if (b.x >= w.n_owned) flags_msg[b.x - w.n_owned].w = 1.0f;
        if (b.y >= w.n_owned) flags_msg[b.y - w.n_owned].w = 1.0f;
    }
}
// scale inverse mass and inverse inertia of the flagged rows by `k` (2 to split, 0.5 to restore: exact)
__global__ void k_strip_scale_mass(DevWorld w, const float4 *flags, unsigned int *keep, const unsigned int *idx, const unsigned char *count_msg,
                                   const unsigned int *ghost_cnt, int first_row, float k, int latch) {
    const unsigned int count = count_msg ? *reinterpret_cast<const unsigned int *>(count_msg) : *ghost_cnt;
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x) {
        if (latch) keep[r] = (flags[r].w != 0.0f) ? 1u : 0u;
        if (!keep[r]) continue;
        const int i = idx ? (int) idx[r] : first_row + (int) r;
        float4 v = w.vel[i], mp = w.mprop[i];
        v.w = v.w * k;
        mp.y = mp.y * k;
        w.vel[i] = v;
        w.mprop[i] = mp;
    }
}
// Per solver stage a strip runs TWO small kernels around the exchange (one before the first stage):
//   stage_begin  remembers the velocities of my ghosts and of my left-border bodies
//   stage_pack   ghosts: what the stage did to each (sent right, to the owner; halved for mass-split rows);
//                border bodies: their velocity after the stage (sent left); a mass-split body keeps only half of the
//                change its (twice too light) copy received
//   stage_apply  owner side: add the left neighbour's changes to my border bodies; ghost side: the owner's new
//                velocity plus my own change = the same bits the owner ends up with; both results are also the
//                "before" values of the next stage
__global__ void k_strip_stage_begin(DevWorld w, float4 *ghost_before, const unsigned int *ghost_cnt, const unsigned char *full_msg,
                                    const unsigned int *border_idx, float4 *border_before) {
    const unsigned int ghosts = *ghost_cnt, border = *reinterpret_cast<const unsigned int *>(full_msg);
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < ghosts; r += gridDim.x * blockDim.x)
        ghost_before[r] = w.vel[w.n_owned + (int) r];
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < border; r += gridDim.x * blockDim.x)
        border_before[r] = w.vel[border_idx[r]];
}
__global__ void k_strip_stage_pack(DevWorld w, const float4 *ghost_before, const unsigned int *ghost_split, float4 *delta,
                                   const unsigned int *ghost_cnt, const unsigned char *full_msg, const unsigned int *border_idx,
                                   const float4 *border_before, const unsigned int *border_split, float4 *border_vel) {
    const unsigned int ghosts = *ghost_cnt, border = *reinterpret_cast<const unsigned int *>(full_msg);
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < ghosts; r += gridDim.x * blockDim.x) {
        const float4 a = w.vel[w.n_owned + (int) r], b = ghost_before[r];
        const float k = ghost_split[r] ? 0.5f : 1.0f;
        delta[r] = make_float4(k * (a.x - b.x), k * (a.y - b.y), k * (a.z - b.z), 0.0f);
    }
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < border; r += gridDim.x * blockDim.x) {
        float4 v = w.vel[border_idx[r]];
        if (border_split[r]) {
            const float4 b = border_before[r];
            v.x = b.x + 0.5f * (v.x - b.x); v.y = b.y + 0.5f * (v.y - b.y); v.z = b.z + 0.5f * (v.z - b.z);
            w.vel[border_idx[r]] = v;
        }
        border_vel[r] = v;
    }
}
__global__ void k_strip_stage_apply(DevWorld w, const unsigned char *full_msg, const unsigned int *border_idx, const float4 *recv_delta,
                                    const float4 *recv_vel, const float4 *my_delta, const unsigned int *ghost_cnt,
                                    float4 *ghost_before, float4 *border_before) {
    const unsigned int ghosts = *ghost_cnt, border = *reinterpret_cast<const unsigned int *>(full_msg);
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < border; r += gridDim.x * blockDim.x) {
        float4 v = w.vel[border_idx[r]];
        const float4 d = recv_delta[r];
        v.x = v.x + d.x; v.y = v.y + d.y; v.z = v.z + d.z;
        w.vel[border_idx[r]] = v;
        border_before[r] = v;
    }
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < ghosts; r += gridDim.x * blockDim.x) {
        float4 v = w.vel[w.n_owned + (int) r];          // keeps the (possibly split) inverse mass in .w
        const float4 o = recv_vel[r], d = my_delta[r];
        v.x = o.x + d.x; v.y = o.y + d.y; v.z = o.z + d.z;
        w.vel[w.n_owned + (int) r] = v;
        ghost_before[r] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Peer-direct transport (one process per GPU, NVLink): instead of handing a message to ncclSend / ncclRecv, the
// sender's kernel STORES it straight into the neighbour's receive buffer (the neighbour's "inbox" allocation, mapped
// into this process through a CUDA IPC handle) and then raises a flag word in the same inbox; the receiver's stream
// runs a one-thread kernel that waits for the flag before the kernels that read the buffer.  Messages are count-sized
// (the count lives on the device), one launch per direction, no collective call on the step's critical path.
//   ordering   every block fences its stores system-wide, the LAST block to finish raises the flag (release);
//              the waiter polls with acquire loads at system scope
//   reuse      the per-stage buffers alternate by round parity: round e + 2 overwrites what round e used only after
//              the neighbour's round e + 1 message has arrived, which it sent after consuming round e
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_BLOCK) k_strip_push(const uint4 *src, uint4 *dst, const unsigned int *count_ptr, unsigned int header_q,
                                                         unsigned int q_per_item, unsigned int max_items, unsigned int *done,
                                                         unsigned int *peer_flag, unsigned int epoch) {
    const unsigned int items = min(*count_ptr, max_items);
    const unsigned int total = header_q + items * q_per_item;
    for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) dst[q] = src[q];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1u) {
        *done = 0u;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(epoch) : "memory");
    }
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// waits until the flag(s) have reached their epochs.  A lost neighbour must not hang the GPU: after 10 s the waiter gives
// up, reports it (overflow bit 64) and latches `*dead`, which makes every later wait of this strip return at once.
__global__ void k_strip_wait(DevWorld w, const unsigned int *flag_a, unsigned int epoch_a, const unsigned int *flag_b, unsigned int epoch_b,
                             unsigned int *dead) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (*dead) { atomicOr(&w.ctr->overflow, 64u); return; }
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const bool a = flag_a == nullptr || (int) (ld_acquire_sys(flag_a) - epoch_a) >= 0;
        const bool b = flag_b == nullptr || (int) (ld_acquire_sys(flag_b) - epoch_b) >= 0;
        if (a && b) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { *dead = 1u; atomicOr(&w.ctr->overflow, 64u); return; }
    }
}

static inline int grid_for(long long n, int block, int max_blocks) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int) g;
}

void fx_launch_strip_pack(const DevWorld &w, const FxLaunchCtx &cx, int side, float cut, float margin, float stray_lo, float stray_hi, unsigned char *msg,
                          unsigned int *border_idx, unsigned int capacity) {
    // the body-prepass status words / ticket are free here (begin_step clears them again afterwards)
    cudaMemsetAsync(cx.status[0], 0, cx.status_words[0] * sizeof(unsigned long long), cx.stream);
    cudaMemsetAsync(cx.tickets, 0, sizeof(unsigned int), cx.stream);
    k_strip_pack<<<grid_for(w.n_owned, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, side, cut, margin, msg, border_idx, capacity,
                                                                                         stray_lo, stray_hi, cx.status[0], cx.tickets);
    ++g_fx_launches;
}
void fx_launch_strip_clear_ghost_map(const DevWorld &w, const FxLaunchCtx &cx, int first_row, int rows) {
    if (rows <= 0) return;
    k_strip_clear_ghost_map<<<grid_for(rows, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, first_row, rows); ++g_fx_launches;
}
void fx_launch_strip_unpack(const DevWorld &w, const FxLaunchCtx &cx, const unsigned char *msg, int first_row, unsigned int capacity,
                            unsigned int *ghost_cnt) {
    k_strip_unpack<<<grid_for(capacity, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, msg, first_row, capacity, ghost_cnt); ++g_fx_launches;
}
void fx_launch_strip_mark_used(const DevWorld &w, const FxLaunchCtx &cx, float4 *flags_msg, unsigned int capacity) {
    k_strip_mark_used<<<grid_for(capacity, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, flags_msg, capacity, 0); ++g_fx_launches;
    k_strip_mark_used<<<grid_for((long long) w.cap_manifolds, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, flags_msg, capacity, 1); ++g_fx_launches;
}
// split (k = 2, latch the flags) or restore (k = 0.5) the inverse masses of the flagged ghosts / border bodies
void fx_launch_strip_scale_mass(const DevWorld &w, const FxLaunchCtx &cx, const float4 *ghost_flags, unsigned int *ghost_keep, const unsigned int *ghost_cnt,
                                const float4 *border_flags, unsigned int *border_keep, const unsigned int *border_idx, const unsigned char *full_msg,
                                unsigned int capacity, float k, int latch) {
    const int g = grid_for(capacity, ST_BLOCK, cx.max_blocks);
    k_strip_scale_mass<<<g, ST_BLOCK, 0, cx.stream>>>(w, ghost_flags, ghost_keep, nullptr, nullptr, ghost_cnt, w.n_owned, k, latch); ++g_fx_launches;
    k_strip_scale_mass<<<g, ST_BLOCK, 0, cx.stream>>>(w, border_flags, border_keep, border_idx, full_msg, nullptr, 0, k, latch); ++g_fx_launches;
}
void fx_launch_strip_stage_begin(const DevWorld &w, const FxLaunchCtx &cx, float4 *ghost_before, const unsigned int *ghost_cnt,
                                 const unsigned char *full_msg, const unsigned int *border_idx, float4 *border_before, unsigned int capacity) {
    k_strip_stage_begin<<<grid_for(capacity, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, ghost_before, ghost_cnt, full_msg, border_idx, border_before);
    ++g_fx_launches;
}
void fx_launch_strip_stage_pack(const DevWorld &w, const FxLaunchCtx &cx, const float4 *ghost_before, const unsigned int *ghost_split, float4 *delta,
                                const unsigned int *ghost_cnt, const unsigned char *full_msg, const unsigned int *border_idx,
                                const float4 *border_before, const unsigned int *border_split, float4 *border_vel, unsigned int capacity) {
    k_strip_stage_pack<<<grid_for(capacity, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, ghost_before, ghost_split, delta, ghost_cnt, full_msg,
                                                                                           border_idx, border_before, border_split, border_vel);
    ++g_fx_launches;
}
void fx_launch_strip_stage_apply(const DevWorld &w, const FxLaunchCtx &cx, const unsigned char *full_msg, const unsigned int *border_idx,
                                 const float4 *recv_delta, const float4 *recv_vel, const float4 *my_delta, const unsigned int *ghost_cnt,
                                 float4 *ghost_before, float4 *border_before, unsigned int capacity) {
    k_strip_stage_apply<<<grid_for(capacity, ST_BLOCK, cx.max_blocks), ST_BLOCK, 0, cx.stream>>>(w, full_msg, border_idx, recv_delta, recv_vel, my_delta, ghost_cnt,
                                                                                            ghost_before, border_before);
    ++g_fx_launches;
}

void fx_launch_strip_push(const FxLaunchCtx &cx, const void *src, void *dst, const unsigned int *count_ptr, unsigned int header_q, unsigned int q_per_item,
                          unsigned int max_items, unsigned int *done, unsigned int *peer_flag, unsigned int epoch) {
    const long long total = (long long) header_q + (long long) max_items * q_per_item;
    int g = grid_for(total, ST_BLOCK, 32);           // small messages: a few blocks keep the "last block" hand-off short
    k_strip_push<<<g, ST_BLOCK, 0, cx.stream>>>((const uint4 *) src, (uint4 *) dst, count_ptr, header_q, q_per_item, max_items, done, peer_flag, epoch);
    ++g_fx_launches;
}
void fx_launch_strip_wait(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *flag_a, unsigned int epoch_a, const unsigned int *flag_b,
                          unsigned int epoch_b, unsigned int *dead) {
    k_strip_wait<<<1, 32, 0, cx.stream>>>(w, flag_a, epoch_a, flag_b, epoch_b, dead);
    ++g_fx_launches;
}
