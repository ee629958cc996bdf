// fx_kernels.h -- host-callable launch surface of the step kernels (implemented in the .cu files).
#pragma once
#include "fx_device.h"

// Per-world launch scratch: stream, grid cap, and the zero-initialised words the ordered (look-back)
// sweeps need.  `fx_launch_begin_step` clears them; each sweep owns one status region + one ticket.
struct FxLaunchCtx {
    cudaStream_t stream;
    int max_blocks;                 // grid cap for grid-stride kernels (multiple of the SM count)
    int coop_blocks;                // co-resident grid of the cooperative solver kernel
    int num_sms;
    int solve_blocks_cap;           // > 0: the coloured solver takes at most this many blocks (FEROX_SOLVE_BLOCKS, read per world: tests)
    unsigned long long *status[4];  // 0 body prepass, 1 pair emission, 2 narrow phase, 3 carry-over
    size_t status_words[4];
    unsigned int *tickets;          // [16]
    unsigned int *sort_hist;        // [FX_SORT_MAX_PASSES][256]
    unsigned int *digit_status;     // [FX_SORT_MAX_PASSES][digit_status_stride]
    size_t digit_status_stride;
    void *zero_base;                // one allocation backs everything above; cleared per step
    size_t zero_bytes;
    size_t zero_fixed_bytes;        // everything ahead of digit_status (a step clears only the digit planes it sorts with)
};

// one LSD radix sort of (key, value) u32 pairs: ping-pong buffers, element count on the device
#define FX_SORT_MAX_PASSES 8         // 8-bit digits of a key of up to 64 bits
struct SortJob {
    unsigned int *keys[2], *vals[2];
    unsigned int *keys_hi[2];         // high key words (nullptr: 32-bit keys)
    const unsigned int *bits_dev;     // device-side key width; > 32 makes the high words live (nullptr: 32-bit keys)
    const unsigned int *n;
    const unsigned int *passes_dev;   // device-side pass count (nullptr: use `passes`)
    unsigned int passes;
    unsigned int *ghist;              // [passes][256], zeroed
    unsigned int *dstatus;            // [passes][dstatus_stride], zeroed
    size_t dstatus_stride;
    unsigned int *tickets;            // [passes], zeroed
};

// scratch of the serial-mode candidate ordering (allocated only when that mode is used)
struct FxLexScratch {
    unsigned int *keys[2], *vals[2];
    int2 *tmp;
    unsigned int *ghist;     // [2][4][256]
    unsigned int *dstatus;   // [2][4][dstatus_stride]
    size_t dstatus_stride;
    unsigned int *tickets;   // [8]
    void *zero_base;
    size_t zero_bytes;
};

// reference-order bookkeeping of the serial solver (fx_solve.cu)
struct SerialOrder {
    unsigned long long *r_key;   // keys in reference array order
    unsigned int *r_len;
    unsigned long long *t_key;   // open-addressing index key -> position in r_key
    unsigned int *t_val;
    unsigned int t_cap;
};

void fx_launch_begin_step(const DevWorld &w, const FxLaunchCtx &cx);
void fx_launch_integrate_velocity(const DevWorld &w, const FxLaunchCtx &cx, bool store_force = false);
void fx_launch_set_scalars(StepScalars *dst, const StepScalars &v, cudaStream_t s);
void fx_init_kernel_attributes();   // per DEVICE: opt-in shared memory sizes (call with the device current)
void fx_launch_broadphase(const DevWorld &w, const FxLaunchCtx &cx, bool fuse_velocity);
void fx_launch_lexsort_candidates(const DevWorld &w, const FxLaunchCtx &cx, const FxLexScratch &lex, int body_bits);
void fx_launch_narrowphase(const DevWorld &w, const FxLaunchCtx &cx);
void fx_launch_cache_update(const DevWorld &w, const FxLaunchCtx &cx);
#define FX_LOCAL_MAX_BODIES 4096
void fx_launch_solve_local(const DevWorld &w, const FxLaunchCtx &cx, int iterations, int max_bodies);
int fx_solve_colored_blocks(const FxLaunchCtx &cx, long long work_hint);
void fx_launch_solve_colored(const DevWorld &w, const FxLaunchCtx &cx, int iterations, long long work_hint, int phases = 7);
#define FX_CLUSTER_MAX_SMEM (220 * 1024)
int fx_cluster_size_for(int bodies);       // 16 / 8: the world's velocities fit one cluster's distributed shared memory; 0: they do not
void fx_launch_solve_cluster(const DevWorld &w, const FxLaunchCtx &cx, int iterations, long long work_hint, int cluster_size);
void fx_launch_serial_order(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o);
void fx_launch_solve_serial_with(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o, int iterations);
void fx_launch_integrate_position(const DevWorld &w, const FxLaunchCtx &cx);
void fx_launch_finish_step(const DevWorld &w, const FxLaunchCtx &cx);
int fx_query_coop_blocks(int device);
void fx_launch_rehash_prev(const DevWorld &w, const FxLaunchCtx &cx);
void fx_launch_serial_reindex(const FxLaunchCtx &cx, const SerialOrder &o);
// collision-handler bridge: live manifolds <-> packed public records, in visiting order (idx: reference-order mode)
void fx_launch_visit_index(const DevWorld &w, const FxLaunchCtx &cx, const SerialOrder &o, unsigned int *idx, unsigned int cap);
void fx_launch_pack_collisions(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *idx, unsigned int n, ManifoldRec *out, int4 *head);
void fx_launch_unpack_collisions(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *idx, unsigned int n, const ManifoldRec *in);

// world ray casts on the device state (fx_bodies.cu): rays = frRay[nrays], hits = frRaycastHit[nrays][max_hits] with the world
// index in the body slot, counts[r] = hits found for ray r (may exceed max_hits)
void fx_launch_raycast_batch(const DevWorld &w, const FxLaunchCtx &cx, const void *rays, int nrays, int max_hits, void *hits, int *counts);

void fx_launch_point_query_batch(const DevWorld &w, const FxLaunchCtx &cx, const void *points, int npoints, int max_hits, int *bodies, int *counts);

// single-pair / utility entry points (used by the public per-body functions and the tests)
void fx_launch_collide_pairs(const DevWorld &w, const FxLaunchCtx &cx, const int2 *pairs, int n, Manifold *out);

// strip decomposition (fx_strip.cu)
void fx_launch_strip_pack(const DevWorld &w, const FxLaunchCtx &cx, int side, float cut, float margin, float stray_lo, float stray_hi, unsigned char *msg,
                          unsigned int *border_idx, unsigned int capacity);
void fx_launch_strip_clear_ghost_map(const DevWorld &w, const FxLaunchCtx &cx, int first_row, int rows);
void fx_launch_strip_unpack(const DevWorld &w, const FxLaunchCtx &cx, const unsigned char *msg, int first_row, unsigned int capacity,
                            unsigned int *ghost_cnt);
void fx_launch_strip_mark_used(const DevWorld &w, const FxLaunchCtx &cx, float4 *flags_msg, unsigned int capacity);
void fx_launch_strip_scale_mass(const DevWorld &w, const FxLaunchCtx &cx, const float4 *ghost_flags, unsigned int *ghost_keep, const unsigned int *ghost_cnt,
                                const float4 *border_flags, unsigned int *border_keep, const unsigned int *border_idx, const unsigned char *full_msg,
                                unsigned int capacity, float k, int latch);
void fx_launch_strip_stage_begin(const DevWorld &w, const FxLaunchCtx &cx, float4 *ghost_before, const unsigned int *ghost_cnt,
                                 const unsigned char *full_msg, const unsigned int *border_idx, float4 *border_before, unsigned int capacity);
void fx_launch_strip_stage_pack(const DevWorld &w, const FxLaunchCtx &cx, const float4 *ghost_before, const unsigned int *ghost_split, float4 *delta,
                                const unsigned int *ghost_cnt, const unsigned char *full_msg, const unsigned int *border_idx,
                                const float4 *border_before, const unsigned int *border_split, float4 *border_vel, unsigned int capacity);
// peer-direct transport: count-sized copy into the neighbour's mapped inbox + flag; wait for the flag(s) on the receiving stream
void fx_launch_strip_push(const FxLaunchCtx &cx, const void *src, void *dst, const unsigned int *count_ptr, unsigned int header_q, unsigned int q_per_item,
                          unsigned int max_items, unsigned int *done, unsigned int *peer_flag, unsigned int epoch);
void fx_launch_strip_wait(const DevWorld &w, const FxLaunchCtx &cx, const unsigned int *flag_a, unsigned int epoch_a, const unsigned int *flag_b,
                          unsigned int epoch_b, unsigned int *dead);
void fx_launch_strip_stage_apply(const DevWorld &w, const FxLaunchCtx &cx, const unsigned char *full_msg, const unsigned int *border_idx,
                                 const float4 *recv_delta, const float4 *recv_vel, const float4 *my_delta, const unsigned int *ghost_cnt,
                                 float4 *ghost_before, float4 *border_before, unsigned int capacity);
